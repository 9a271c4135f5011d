/*
 * pemapper_gpu.c - C host for the B200 hot path: same command line, input formats and output files as the
 * reference pemapper (wingolab-org/pecaller src/pemapper.c main(), 178-904), with the per-batch worker
 * (pthread_create(map_everything), 684/759) replaced by the C-ABI of include/pemap.h.
 *
 *   pemapper_gpu out sdx s|sa file1 is_bisulfite min_match max_threads max_reads
 *   pemapper_gpu out sdx p|pa file1 file2 max_dist min_dist is_bisulfite min_match max_threads max_reads
 * and, with two more arguments, those of pemapper_tsw (src/pemapper_tsw.c main(), tsw:213-335):
 *   pemapper_gpu out sdx s|sa file1 ... max_reads trim_from_start trim_from_end
 *   pemapper_gpu out sdx p|pa file1 file2 ... max_reads trim_from_start trim_from_end
 * In that form every read is trimmed (tsw:693-704, 792-802), a second column of the array file names the sample
 * of each fastq, and whenever the sample changes the outputs are written and the counters zeroed
 * (dump_output, tsw:636-675, 849-965) - pemap_finish + pemap_reset_counts.
 *
 * Threads: one decoder thread per input file of a pair inflates and cuts its FASTQ straight into pinned batch rows
 * (pemap_host_alloc, DMA'd in place by the library); one submitting thread per GPU maps the batches, three batch slots
 * per GPU keep decoding, copying and mapping overlapped; the pileup is compacted on the GPU window by window
 * (pemap_finish_stream) and deflated by a pool of threads into concatenated gzip members, whose inflated stream is
 * the reference's byte for byte.  max_threads (argv) sizes that pool.
 * Environment:
 *   PEMAP_GPUS          number of GPUs of this box to use (default 1): batch b goes to GPU b mod N, every GPU keeps
 *                       its own counters, and pemap_reduce_counts_peer sums them onto GPU 0 over NVLink at the end
 *   PEMAP_DEVICE        first GPU ordinal (default 0)
 *   PEMAP_DEVICE_INDEX  1 = rebuild pos_index/mers on the GPU from .seq/.sdx instead of loading .idx/.mdx
 *   PEMAP_BATCH         reads per pemap_map_batch_rows call (default 1,000,000; results do not depend on it)
 *   PEMAP_GZ_LEVEL      deflate level of .pileup.gz (default 6 = gzopen's, as the reference; 1 is ~3x faster)
 *   PEMAP_TIMING        1 = print wall-clock rates of the stages (decode, map, write) at the end
 * Written from scratch; what must be byte-compatible (file formats, summary text) cites the reference line.
 */
#include <ctype.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#include "pemap.h"

#define ROW 304 /* bytes per read row handed to the library */
#define NAME_MAX_LEN 512
#define MAX_FILES 2000

static void die(const char *msg) { /* dump_error, pemapper.c:2808-2816 */
  printf("\n%s\n", msg);
  exit(1);
}

/* ---- line reader with my_gzgets semantics (2447-2483): '\n'-terminated lines only; a trailing
   unterminated line is dropped; the returned pointer is valid until the next call. */
typedef struct {
  gzFile f;
  char *buf;
  size_t cap, beg, end;
  int eof;
} reader;

static void reader_open(reader *r, const char *path) {
  r->f = gzopen(path, "r");
  if (!r->f) {
    printf("\n Can not open file %s for reading\n", path);
    exit(1);
  }
  gzbuffer(r->f, 1 << 25);
  r->cap = (size_t)64 << 20;
  r->buf = malloc(r->cap + 1);
  r->beg = r->end = 0;
  r->eof = 0;
}

static char *reader_line(reader *r) {
  for (;;) {
    char *nl = r->end > r->beg ? memchr(r->buf + r->beg, '\n', r->end - r->beg) : NULL;
    if (nl) {
      char *line = r->buf + r->beg;
      *nl = '\0';
      r->beg = (size_t)(nl - r->buf) + 1;
      return line;
    }
    if (r->eof) return NULL;
    memmove(r->buf, r->buf + r->beg, r->end - r->beg);
    r->end -= r->beg;
    r->beg = 0;
    if (r->end == r->cap) die(" A fastq line is longer than 64 MB ");
    int got = gzread(r->f, r->buf + r->end, (unsigned)(r->cap - r->end));
    if (got <= 0) r->eof = 1; else r->end += (size_t)got;
  }
}

static void reader_close(reader *r) {
  gzclose(r->f);
  free(r->buf);
}

/* skip '+', quality and look for the next '@' header, then return the sequence line (713-724) */
static char *next_sequence(reader *r) {
  char *s = reader_line(r);
  s = reader_line(r);
  s = reader_line(r);
  int found = 0;
  while (s && !found) {
    if (s[0] == '@') found = 1;
    s = reader_line(r);
  }
  return found ? s : NULL;
}

/* 250-267; outs (may be NULL): second column = sample / output name of the file (tsw:266-280) */
static int read_name_list(const char *path, char (*names)[NAME_MAX_LEN], char (*outs)[NAME_MAX_LEN]) {
  FILE *f = fopen(path, "r");
  if (!f) {
    printf("\n Can not open file %s for reading\n", path);
    exit(1);
  }
  int n = 0;
  char line[NAME_MAX_LEN];
  while (n < MAX_FILES && fgets(line, NAME_MAX_LEN - 1, f)) {
    char *tok = strtok(line, "\t \n");
    if (!tok || strlen(tok) <= 2) break;
    strcpy(names[n], tok);
    tok = strtok(NULL, "\t \n");
    if (outs) strcpy(outs[n], tok ? tok : "");
    n++;
  }
  fclose(f);
  return n;
}

typedef struct {
  uint32_t *starts; /* n+1 unpadded prefix sums */
  char (*names)[260];
  int n, idepth;
} sdx_t;

static void load_sdx(const char *path, sdx_t *s) { /* 411-448 */
  FILE *f = fopen(path, "r");
  char line[1100];
  if (!f) {
    printf("\n Can not open file %s\n", path);
    exit(1);
  }
  fgets(line, 256, f);
  s->n = atoi(line);
  s->starts = calloc((size_t)s->n + 2, 4);
  s->names = calloc((size_t)s->n + 1, 260);
  for (int i = 0; i < s->n; i++) {
    fgets(line, 1024, f);
    char *tok = strtok(line, "\t \n");
    s->starts[i + 1] = s->starts[i] + (uint32_t)atoi(tok);
    tok = strtok(NULL, "\t \n");
    strncpy(s->names[i], tok ? tok : "", 259);
  }
  fgets(line, 1024, f);
  s->idepth = atoi(line);
  fclose(f);
}

static int find_contig(const uint32_t *pos, int n, uint32_t v) { /* find_chrom 2168-2186 */
  int first = 0, last = n - 1, probe = 7;
  for (;;) {
    if (first == last) return first;
    if (pos[probe] <= v && pos[probe + 1] >= v) return probe;
    if (pos[probe] > v) last = probe - 1; else first = probe + 1;
    probe = (last + first) / 2;
  }
}

static void gz_read_all(gzFile f, void *dst, size_t n) {
  size_t done = 0;
  while (done < n) {
    unsigned want = (unsigned)((n - done) > (1u << 30) ? (1u << 30) : (n - done));
    int got = gzread(f, (char *)dst + done, want);
    if (got <= 0) break;
    done += (size_t)got;
  }
  if (done != n) die(" Short read on a compressed index file ");
}

/* ---- batch slots: pinned rows filled by the decoder threads, mapped by the GPU's submitting thread ---- */
typedef struct slot {
  char *rows1, *rows2;
  int *len1, *len2;
  uint32_t *bm1, *bm2;
  int *btype;
  long filled, first; /* reads in the batch, index of its first read in the file */
  struct slot *next;
} slot_t;

static void *pinned(size_t bytes) {
  void *p = pemap_host_alloc(bytes ? bytes : 1);
  if (!p) die(" Could not allocate the (page-locked) batch buffers ");
  return p;
}

static slot_t *slot_new(long cap, int paired) {
  slot_t *s = calloc(1, sizeof *s);
  s->rows1 = pinned((size_t)cap * ROW);
  s->rows2 = paired ? pinned((size_t)cap * ROW) : NULL;
  s->len1 = pinned(sizeof(int) * (size_t)cap);
  s->len2 = pinned(sizeof(int) * (size_t)cap);
  s->bm1 = pinned(4 * (size_t)cap);
  s->bm2 = pinned(4 * (size_t)cap);
  s->btype = pinned(sizeof(int) * (size_t)cap);
  return s;
}

/* pemap_fill_cb over a gzFile: the next `bytes` bytes of the inflated .idx */
static int idx_chunk(void *ctx, void *dst, size_t bytes) {
  size_t done = 0;
  while (done < bytes) {
    const unsigned want = (unsigned)((bytes - done) > (1u << 30) ? (1u << 30) : (bytes - done));
    const int got = gzread((gzFile)ctx, (char *)dst + done, want);
    if (got <= 0) return 1;
    done += (size_t)got;
  }
  return 0;
}

/* ---- one submitting thread per GPU (replaces the reference's pool of map_everything threads, 677-702) ---- */
typedef struct {
  pemap_t *h;
  int paired;
  uint32_t *maps1, *maps2;       /* shared per-file arrays; batches write disjoint ranges */
  long max_dist;
  long mate_counts[9], total_reads, total_bases, total_dist, no_dists; /* per worker, summed at the end */
  pthread_t thread;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  slot_t *q_head, *q_tail;       /* batches waiting to be mapped, FIFO */
  slot_t *free_list;             /* this GPU's idle slots */
  int busy, quit;
  double map_seconds;
} worker_t;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void worker_epilogue(worker_t *w, const slot_t *b) { /* batch epilogue, 1238-1265 */
  for (long j = 0; j < b->filled; j++) {
    w->mate_counts[b->btype[j]]++;
    w->maps1[b->first + j] = b->bm1[j];
    if (b->bm1[j]) {
      w->total_reads++;
      w->total_bases += b->len1[j];
      if (b->bm2[j]) {
        w->total_reads++;
        w->total_bases += b->len2[j];
        long test = (long)(uint32_t)(b->bm1[j] - b->bm2[j]); /* labs() of an unsigned difference (1250) */
        w->maps2[b->first + j] = b->bm2[j];
        if (test < w->max_dist * 4) {
          w->total_dist += test;
          w->no_dists++;
        }
      }
    } else if (b->bm2[j]) {
      w->total_reads++;
      w->total_bases += b->len2[j];
      w->maps2[b->first + j] = b->bm2[j];
    }
  }
}

static void *worker_main(void *arg) {
  worker_t *w = arg;
  for (;;) {
    pthread_mutex_lock(&w->mu);
    while (!w->q_head && !w->quit) pthread_cond_wait(&w->cv, &w->mu);
    slot_t *b = w->q_head;
    if (!b) {
      pthread_mutex_unlock(&w->mu);
      return NULL;
    }
    w->q_head = b->next;
    if (!w->q_head) w->q_tail = NULL;
    w->busy = 1;
    pthread_mutex_unlock(&w->mu);
    const double t0 = now_s();
    int rc = pemap_map_batch_rows(w->h, (int)b->filled, b->rows1, b->len1, b->rows2, w->paired ? b->len2 : NULL, ROW, b->bm1,
                                  b->bm2, b->btype);
    if (rc) {
      printf("\n pemap_map_batch failed: %s \n", pemap_last_error(w->h));
      exit(1);
    }
    worker_epilogue(w, b);
    w->map_seconds += now_s() - t0;
    pthread_mutex_lock(&w->mu);
    b->next = w->free_list;
    w->free_list = b;
    w->busy = 0;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
  }
}

static void worker_wait_idle(worker_t *w) { /* everything submitted so far has been mapped and booked */
  pthread_mutex_lock(&w->mu);
  while (w->q_head || w->busy) pthread_cond_wait(&w->cv, &w->mu);
  pthread_mutex_unlock(&w->mu);
}

static slot_t *worker_take_free(worker_t *w) {
  pthread_mutex_lock(&w->mu);
  while (!w->free_list) pthread_cond_wait(&w->cv, &w->mu);
  slot_t *b = w->free_list;
  w->free_list = b->next;
  pthread_mutex_unlock(&w->mu);
  b->next = NULL;
  return b;
}

static void worker_submit(worker_t *w, slot_t *b) {
  pthread_mutex_lock(&w->mu);
  b->next = NULL;
  if (w->q_tail) w->q_tail->next = b; else w->q_head = b;
  w->q_tail = b;
  pthread_cond_broadcast(&w->cv);
  pthread_mutex_unlock(&w->mu);
}

static void worker_quit(worker_t *w) {
  pthread_mutex_lock(&w->mu);
  w->quit = 1;
  pthread_cond_broadcast(&w->cv);
  pthread_mutex_unlock(&w->mu);
}

static int cmp_ins(const void *a, const void *b) {
  const pemap_insertion *x = a, *y = b;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return strcmp(x->seq, y->seq);
}

/* tsw:693-704, 792-802: skip trim_from_start characters, drop trim_from_end from the end (never below length 0) */
static char *trim_read(char *s, int trim_start, int trim_end) {
  if (!s || (trim_start == 0 && trim_end == 0)) return s;
  size_t n = strlen(s);
  s += (size_t)trim_start <= n ? (size_t)trim_start : n; /* the reference would run past a short line; we stop at its end */
  long len = (long)strlen(s) - trim_end;
  if (len < 0) len = 0;
  s[len] = '\0';
  return s;
}

/* ---- parallel deflate of the pileup: blocks of records become gzip members, written in order.  gzread (pecaller's
   reader, pecaller.c:839-845) and gunzip inflate concatenated members into one stream: the records, byte for byte. ---- */
#define GZ_BLOCK_RECORDS 65536 /* 1 MB of records per member */
typedef struct {
  int n_threads, level;
  pthread_t *threads;
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_done;
  const pemap_record *rec; /* the window being written */
  uint64_t n_rec;
  long n_blocks, next_block, done_blocks, generation;
  unsigned char **out; /* per block: compressed member */
  size_t *out_len, *out_cap;
  long out_slots;
  int quit;
  double seconds;
  uint64_t bytes_in, bytes_out;
} gzpool_t;

static void gz_one_block(gzpool_t *g, long blk) {
  const uint64_t first = (uint64_t)blk * GZ_BLOCK_RECORDS;
  const uint64_t n = g->n_rec - first < GZ_BLOCK_RECORDS ? g->n_rec - first : GZ_BLOCK_RECORDS;
  z_stream z;
  memset(&z, 0, sizeof z);
  if (deflateInit2(&z, g->level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) die(" deflateInit2 failed ");
  const size_t bound = deflateBound(&z, (uLong)(n * 16)) + 64;
  if (g->out_cap[blk] < bound) {
    free(g->out[blk]);
    g->out[blk] = malloc(bound);
    g->out_cap[blk] = bound;
  }
  z.next_in = (Bytef *)(g->rec + first);
  z.avail_in = (uInt)(n * 16);
  z.next_out = g->out[blk];
  z.avail_out = (uInt)bound;
  if (deflate(&z, Z_FINISH) != Z_STREAM_END) die(" deflate failed ");
  g->out_len[blk] = bound - z.avail_out;
  deflateEnd(&z);
}

static void *gz_thread(void *arg) {
  gzpool_t *g = arg;
  for (;;) {
    pthread_mutex_lock(&g->mu);
    while (!g->quit && g->next_block >= g->n_blocks) pthread_cond_wait(&g->cv_work, &g->mu);
    if (g->quit) {
      pthread_mutex_unlock(&g->mu);
      return NULL;
    }
    const long blk = g->next_block++;
    pthread_mutex_unlock(&g->mu);
    gz_one_block(g, blk);
    pthread_mutex_lock(&g->mu);
    if (++g->done_blocks == g->n_blocks) pthread_cond_broadcast(&g->cv_done);
    pthread_mutex_unlock(&g->mu);
  }
}

static void gzpool_start(gzpool_t *g, int n_threads, int level) {
  memset(g, 0, sizeof *g);
  g->n_threads = n_threads < 1 ? 1 : n_threads;
  g->level = level;
  pthread_mutex_init(&g->mu, NULL);
  pthread_cond_init(&g->cv_work, NULL);
  pthread_cond_init(&g->cv_done, NULL);
  g->threads = calloc((size_t)g->n_threads, sizeof(pthread_t));
  for (int i = 0; i < g->n_threads; i++)
    if (pthread_create(&g->threads[i], NULL, gz_thread, g)) die(" Could not start a compression thread ");
}

/* compress `n` records with every thread of the pool and append the members to `f` in order */
static void gzpool_write(gzpool_t *g, FILE *f, const pemap_record *rec, uint64_t n) {
  if (!n) return;
  const double t0 = now_s();
  const long nb = (long)((n + GZ_BLOCK_RECORDS - 1) / GZ_BLOCK_RECORDS);
  if (nb > g->out_slots) {
    g->out = realloc(g->out, sizeof(*g->out) * (size_t)nb);
    g->out_len = realloc(g->out_len, sizeof(size_t) * (size_t)nb);
    g->out_cap = realloc(g->out_cap, sizeof(size_t) * (size_t)nb);
    for (long i = g->out_slots; i < nb; i++) {
      g->out[i] = NULL;
      g->out_len[i] = g->out_cap[i] = 0;
    }
    g->out_slots = nb;
  }
  pthread_mutex_lock(&g->mu);
  g->rec = rec;
  g->n_rec = n;
  g->n_blocks = nb;
  g->next_block = 0;
  g->done_blocks = 0;
  g->generation++;
  pthread_cond_broadcast(&g->cv_work);
  while (g->done_blocks < nb) pthread_cond_wait(&g->cv_done, &g->mu);
  g->n_blocks = 0; /* nothing left to hand out until the next window */
  pthread_mutex_unlock(&g->mu);
  for (long i = 0; i < nb; i++) {
    if (fwrite(g->out[i], 1, g->out_len[i], f) != g->out_len[i]) die(" Short write on the pileup file ");
    g->bytes_out += g->out_len[i];
  }
  g->bytes_in += n * 16;
  g->seconds += now_s() - t0;
}

static void gzpool_stop(gzpool_t *g) {
  pthread_mutex_lock(&g->mu);
  g->quit = 1;
  pthread_cond_broadcast(&g->cv_work);
  pthread_mutex_unlock(&g->mu);
  for (int i = 0; i < g->n_threads; i++) pthread_join(g->threads[i], NULL);
  for (long i = 0; i < g->out_slots; i++) free(g->out[i]);
  free(g->out);
  free(g->out_len);
  free(g->out_cap);
  free(g->threads);
}

/* ---- outputs: the writer loop of main() (819-900) and, for the tsw form, dump_output (tsw:849-965) ---- */
typedef struct {
  FILE *pile; /* raw file: the gzip members come from the pool */
  gzFile indel;
  FILE *summary;
  pemap_t **hs;
  int n_gpus, device, paired, tsw;
  worker_t *ws;
  const sdx_t *sdx;
  const char *genome;
  uint64_t genome_size;
  long mate_counts[9], total_reads, total_bases, total_dist, no_dists, tot_pairs;
  gzpool_t *gz;
  /* state of the streaming writer */
  const pemap_insertion *ins;
  uint64_t n_ins, ins_at, n_records;
  uint32_t *padded;
} out_t;

static void open_outputs(out_t *o, const char *base) {
  char path[4300];
  snprintf(path, sizeof path, "%s.pileup.gz", base);
  o->pile = fopen(path, "wb");
  if (!o->pile) die(" Can not open the pileup file for writing ");
  setvbuf(o->pile, NULL, _IOFBF, 1 << 24);
  snprintf(path, sizeof path, "%s.indel.txt.gz", base);
  o->indel = gzopen(path, "w");
  if (!o->indel) die(" Can not open the indel file for writing ");
  gzbuffer(o->indel, 33554432);
  snprintf(path, sizeof path, "%s.summary.txt", base);
  o->summary = fopen(path, "w");
  if (!o->summary) die(" Can not open the summary file for writing ");
}

/* pemap_site_cb: one window of records in ascending coordinate (828-864) */
static int write_window(void *ctx, const pemap_record *rec, uint64_t n) {
  out_t *o = ctx;
  gzpool_write(o->gz, o->pile, rec, n);
  const sdx_t *sdx = o->sdx;
  for (uint64_t k = 0; k < n; k++) {
    if (rec[k].c[5] == 0) continue;
    const uint32_t pos = rec[k].pos;
    const char ref = o->genome[pos];
    const int tot = rec[k].c[0] + rec[k].c[1] + rec[k].c[2] + rec[k].c[3] + rec[k].c[4] + rec[k].c[5];
    const int ref_reads = ref == 'A' ? rec[k].c[0] : ref == 'C' ? rec[k].c[1] : ref == 'G' ? rec[k].c[2] : rec[k].c[3];
    const int which = find_contig(o->padded, sdx->n, pos);
    gzprintf(o->indel, "\n%s\t%d\t%c\t%d\t%d\t%d\t%d", sdx->names[which], (int)(1 + pos - o->padded[which]), ref, tot, ref_reads,
             rec[k].c[4], rec[k].c[5]);
    while (o->ins_at < o->n_ins && o->ins[o->ins_at].pos < pos) o->ins_at++;
    for (; o->ins_at < o->n_ins && o->ins[o->ins_at].pos == pos; o->ins_at++) gzprintf(o->indel, "\t%s", o->ins[o->ins_at].seq);
  }
  o->n_records += n;
  return 0;
}

typedef struct {
  pemap_t **hs;
  int n, which, rc;
  uint64_t s0, s1;
} rs_job;
static void *rs_thread(void *arg) {
  rs_job *j = arg;
  j->rc = pemap_reduce_scatter_local(j->hs, j->n, j->which, &j->s0, &j->s1);
  return NULL;
}

static void write_summary(FILE *f, const out_t *o, const char **names, double avg_len, double avg_depth, double avg_dist) {
  const char *bars = "\n================================================================";
  fprintf(f, "%s\n================= Summary ======================================%s%s", bars, bars, bars);
  if (o->total_bases <= 0)
    fprintf(f, "\n\nTotal Number of Mapping reads of Any Kind\t0\tWith average Length\t0\tAverage Depth\t0\tAverage Insert Size\t0");
  else
    fprintf(f, "\n\nTotal Number of Mapping reads of Any Kind\t%ld\tWith average Length\t%g\tAverage Depth\t%g\tAverage Insert Size\t%g",
            o->total_reads, avg_len, avg_depth, avg_dist);
  fprintf(f, "\n\nMapping Type\tCount\tFraction");
  fprintf(f, "\nAll\t%ld\t1", o->tot_pairs);
  for (int i = 0; i < 9; i++)
    if (names[i]) fprintf(f, "\n%s\t%ld\t%g", names[i], o->mate_counts[i], (double)o->mate_counts[i] / (double)o->tot_pairs);
  fprintf(f, "\n");
}

/* Collect the workers' statistics, write pileup / indel / summary of what has been mapped since the last call and
   (tsw form) zero the counters for the next sample.  Returns 1 when nothing mapped (790-808 / tsw:855-873). */
static int dump_output(out_t *o) {
  const char *pn[9] = {"Unique Mate-Paired", "Unique Mate-Paired with slip", "Unique Single End", "Unique Mis-size",
                       "Non-Unique Mate-Paired", "Non-Unique Mis-size", "Fragment Mismatch", "Non-unique with no map",
                       "Neither Map"}; /* 567-590 */
  const char *sn[9] = {NULL, NULL, "Unique Mapping", NULL, NULL, NULL, NULL, "Non-Unique Mapping, discarded",
                       "No mapping reaches threshold"};
  const char **names = o->paired ? pn : sn;
  for (int g = 0; g < o->n_gpus; g++) {
    worker_t *w = &o->ws[g];
    worker_wait_idle(w);
    for (int i = 0; i < 9; i++) {
      o->mate_counts[i] += w->mate_counts[i];
      w->mate_counts[i] = 0;
    }
    o->total_reads += w->total_reads;
    o->total_bases += w->total_bases;
    o->total_dist += w->total_dist;
    o->no_dists += w->no_dists;
    w->total_reads = w->total_bases = w->total_dist = w->no_dists = 0;
  }
  if (o->total_bases <= 0) { /* nothing mapped: summary only, and (tsw quirk) nothing is reset */
    write_summary(o->summary, o, names, 0, 0, 0);
    fclose(o->summary);
    return 1;
  }

  /* insertion strings of every GPU, sorted by site */
  pemap_insertion *all_ins = NULL;
  uint64_t n_all = 0;
  int rc;
  for (int g = 0; g < o->n_gpus; g++) {
    const pemap_insertion *ig;
    uint64_t ng;
    rc = pemap_get_insertions(o->hs[g], &ig, &ng);
    if (rc) {
      printf("\n pemap_get_insertions failed on GPU %d: %s \n", o->device + g, pemap_last_error(o->hs[g]));
      exit(1);
    }
    all_ins = realloc(all_ins, (size_t)(n_all + ng + 1) * sizeof(pemap_insertion));
    memcpy(all_ins + n_all, ig, (size_t)ng * sizeof(pemap_insertion));
    n_all += ng;
  }
  if (o->n_gpus > 1) qsort(all_ins, (size_t)n_all, sizeof(pemap_insertion), cmp_ins);
  o->ins = all_ins;
  o->n_ins = n_all;
  o->ins_at = 0;
  o->n_records = 0;
  const sdx_t *sdx = o->sdx;
  gzprintf(o->indel, "Fragment\tPositions\tReference Base\tTotal Coverage\tReference Reads\tNo Deletions\tNo Insertions\tInsertion Sequence"); /* 819-820 */
  o->padded = calloc((size_t)sdx->n + 16, 4);
  for (int i = 0; i <= sdx->n; i++) o->padded[i] = sdx->starts[i] + 15u * (uint32_t)i; /* 821-822 */
  /* the counters of all GPUs are summed slice-wise over NVLink (every GPU pulls its share of the genome from the
     others, all at once), then every GPU's slice is compacted and written in coordinate order */
  rs_job jobs[16];
  pthread_t rt[16];
  for (int g = 0; g < o->n_gpus; g++) {
    jobs[g].hs = o->hs;
    jobs[g].n = o->n_gpus;
    jobs[g].which = g;
    jobs[g].rc = 0;
    jobs[g].s0 = 0;
    jobs[g].s1 = o->genome_size;
    if (o->n_gpus > 1 && pthread_create(&rt[g], NULL, rs_thread, &jobs[g])) die(" Could not start a reduce thread ");
  }
  for (int g = 0; g < o->n_gpus && o->n_gpus > 1; g++) {
    pthread_join(rt[g], NULL);
    if (jobs[g].rc) {
      printf("\n summing the pileup counters failed on GPU %d: %s \n", o->device + g, pemap_last_error(o->hs[g]));
      exit(1);
    }
  }
  for (int g = 0; g < o->n_gpus; g++) {
    rc = pemap_finish_stream_range(o->hs[g], jobs[g].s0, jobs[g].s1, write_window, o, NULL);
    if (rc) {
      printf("\n pemap_finish failed on GPU %d: %s \n", o->device + g, pemap_last_error(o->hs[g]));
      exit(1);
    }
  }
  free(o->padded);
  free(all_ins);
  fclose(o->pile);
  gzclose(o->indel);

  double avg_len = (double)o->total_bases, avg_dist = (double)o->total_dist; /* 811-817, 868 */
  if (o->total_reads > 0) avg_len /= (double)o->total_reads;
  if (o->no_dists > 0) avg_dist /= (double)o->no_dists;
  const double avg_depth = (double)o->total_bases / (double)o->genome_size;
  if (!o->tsw) write_summary(stdout, o, names, avg_len, avg_depth, avg_dist); /* 870-883: pemapper also prints it */
  write_summary(o->summary, o, names, avg_len, avg_depth, avg_dist);
  fclose(o->summary);
  if (o->tsw) { /* tsw:932, 957-962: ready for the next sample */
    for (int g = 0; g < o->n_gpus; g++) {
      rc = pemap_reset_counts(o->hs[g]);
      if (rc) {
        printf("\n pemap_reset_counts failed: %s \n", pemap_last_error(o->hs[g]));
        exit(1);
      }
    }
    o->total_reads = o->total_bases = o->total_dist = o->no_dists = 0;
    for (int i = 0; i < 9; i++) o->mate_counts[i] = 0;
  }
  return 0;
}

/* ---- FASTQ decoder: one per input file, fills the rows of a batch slot.  Reads are taken exactly as main() takes them
   (663-739): the second line of the file, then after every read three lines are skipped and the next line that starts
   with '@' is a header whose following line is the read. ---- */
typedef struct {
  reader r;
  char *pending; /* the next read's line (inside the reader's buffer), or NULL at the end of the file */
  int trim_start, trim_end, check_short;
  char *rows;
  int *len;
  long want, got;
} decoder_t;

static void decoder_prime(decoder_t *d) {
  char *s = reader_line(&d->r);
  (void)s;
  d->pending = trim_read(reader_line(&d->r), d->trim_start, d->trim_end);
}

static void *decoder_fill(void *arg) {
  decoder_t *d = arg;
  long n = 0;
  while (n < d->want && d->pending) {
    const int l = (int)strlen(d->pending);
    if (d->check_short && l <= 12) { /* 663: a first-file read of 12 bases or fewer ends the file */
      d->pending = NULL;
      break;
    }
    if (l > PEMAP_MAX_READ + 20) die(" Read longer than the reference's DP buffers allow ");
    memcpy(d->rows + (size_t)n * ROW, d->pending, (size_t)l);
    d->len[n] = l;
    n++;
    d->pending = trim_read(next_sequence(&d->r), d->trim_start, d->trim_end);
  }
  d->got = n;
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 4) die("Usage: pemapper_gpu out_file sdx_file [s,sa,p,pa] ... (same arguments as pemapper)");
  const char mode = (char)toupper(argv[3][0]), arr = (char)toupper(argv[3][1]);
  int paired, max_dist = 0, min_dist = 0, bisulfite;
  double min_align;
  long max_reads;
  int max_threads = 0;
  const char *in1, *in2 = NULL;
  int tsw = 0, trim_start = 0, trim_end = 0;
  if (mode == 'S') { /* 233-280 */
    if (argc != 9 && argc != 11)
      die("Usage: pemapper_gpu out_file sdx_file [s,sa] file1 is_bisulfite[y,n] min_match_percentage max_threads max_reads [trim_from_start trim_from_end]");
    if (argc == 11) {
      tsw = 1;
      trim_start = atoi(argv[9]);
      trim_end = atoi(argv[10]);
    }
    paired = 0;
    in1 = argv[4];
    bisulfite = strchr(argv[5], 'Y') || strchr(argv[5], 'y');
    min_align = atof(argv[6]);
    max_threads = atoi(argv[7]);
    max_reads = atoi(argv[8]);
  } else if (mode == 'P') { /* 281-358 */
    if (argc != 12 && argc != 14)
      die("Usage: pemapper_gpu out_file sdx_file [p,pa] file1 file2 max_dist min_dist is_bisulfite[y,n] min_match_percentage max_threads max_reads [trim_from_start trim_from_end]");
    if (argc == 14) {
      tsw = 1;
      trim_start = atoi(argv[12]);
      trim_end = atoi(argv[13]);
    }
    paired = 1;
    in1 = argv[4];
    in2 = argv[5];
    max_dist = atoi(argv[6]);
    min_dist = atoi(argv[7]);
    bisulfite = strchr(argv[8], 'Y') || strchr(argv[8], 'y');
    min_align = atof(argv[9]);
    max_threads = atoi(argv[10]);
    max_reads = atol(argv[11]);
  } else {
    die("Usage: pemapper_gpu out_file sdx_file paired_or_single_or_array[p,s,pa,ps] ...");
    return 1;
  }
  static char files1[MAX_FILES][NAME_MAX_LEN], files2[MAX_FILES][NAME_MAX_LEN], out_names[MAX_FILES][NAME_MAX_LEN];
  int n_files = 1;
  if (arr == 'A') {
    n_files = read_name_list(in1, files1, tsw ? out_names : NULL);
    if (paired && read_name_list(in2, files2, NULL) != n_files) die(" Mismatch in number of files in the two arrays ");
  } else {
    strcpy(files1[0], in1);
    if (paired) strcpy(files2[0], in2);
  }

  char path[4300], base[1024], sdxbase[1024];
  strcpy(base, argv[1]);
  out_t out;
  memset(&out, 0, sizeof out);
  if (!tsw) open_outputs(&out, base); /* 374-393; the tsw form opens them per sample (tsw:636-675) */

  sdx_t sdx;
  load_sdx(argv[2], &sdx);
  strcpy(sdxbase, argv[2]);
  if (strstr(sdxbase, ".sdx")) *strrchr(sdxbase, '.') = '\0'; /* 400-408 */
  const uint64_t genome_size = (uint64_t)sdx.starts[sdx.n] + 15ull * (uint64_t)sdx.n; /* 453 */
  printf("\n Genome size is %llu \n\n", (unsigned long long)genome_size);
  char *genome = malloc(genome_size + 1);
  snprintf(path, sizeof path, "%s.seq", sdxbase);
  gzFile gf = gzopen(path, "r");
  if (!gf) die(" Can not open the .seq file ");
  gzbuffer(gf, 33554432);
  gz_read_all(gf, genome, genome_size);
  gzclose(gf);

  pemap_params prm;
  pemap_default_params(&prm);
  prm.idepth = sdx.idepth;
  prm.min_align = min_align;
  prm.is_bisulfite = bisulfite;
  prm.pair_flag = paired;
  prm.min_dist = min_dist;
  prm.max_dist = max_dist;
  const int device = getenv("PEMAP_DEVICE") ? atoi(getenv("PEMAP_DEVICE")) : 0;
  int n_gpus = getenv("PEMAP_GPUS") ? atoi(getenv("PEMAP_GPUS")) : 1;
  if (n_gpus < 1) n_gpus = 1;
  if (n_gpus > 16) n_gpus = 16;
  pemap_t *hs[16] = {NULL};
  int rc = 0;
  if (getenv("PEMAP_DEVICE_INDEX") && atoi(getenv("PEMAP_DEVICE_INDEX"))) {
    int64_t *lens = malloc(sizeof(int64_t) * (size_t)sdx.n);
    for (int i = 0; i < sdx.n; i++) lens[i] = (int64_t)(sdx.starts[i + 1] - sdx.starts[i]) + 15;
    for (int g = 0; g < n_gpus && !rc; g++) {
      rc = pemap_init_from_genome(&hs[g], genome, lens, sdx.n, &prm, device + g);
      if (rc) printf("\n pemap_init failed on GPU %d: %s \n", device + g, pemap_last_error(hs[g]));
    }
    free(lens);
  } else if (n_gpus == 1) { /* init_index_buffer 2129-2155, streamed: gzread inflates straight into the library's pinned staging */
    snprintf(path, sizeof path, "%s.mdx", sdxbase);
    FILE *mf = fopen(path, "rb");
    if (!mf) die(" Could not read the .mdx file ");
    fseek(mf, 0, SEEK_END);
    const uint64_t n_mers = (uint64_t)ftell(mf) / 4;
    fseek(mf, 0, SEEK_SET);
    uint32_t *mers = malloc((n_mers + 1) * 4);
    if (!mers || fread(mers, 4, n_mers, mf) != n_mers) die(" Could not read the .mdx file ");
    fclose(mf);
    snprintf(path, sizeof path, "%s.idx", sdxbase);
    gf = gzopen(path, "r");
    if (!gf) die(" Could Not Open the .idx file ");
    gzbuffer(gf, 33554432);
    printf("\n About to read kmers index \n\n");
    pemap_index ix = {NULL, mers, n_mers, genome, genome_size, sdx.starts, sdx.n};
    rc = pemap_init_streamed(&hs[0], &ix, idx_chunk, gf, &prm, device);
    if (rc) printf("\n pemap_init failed on GPU %d: %s \n", device, pemap_last_error(hs[0]));
    gzclose(gf);
    free(mers);
  } else { /* several GPUs: inflate once into the host table, every GPU gets a copy */
    const size_t words = ((size_t)1 << 32) + 1;
    uint32_t *pos_index = malloc(words * 4);
    if (!pos_index) die(" Can not allocate space for the position index ");
    snprintf(path, sizeof path, "%s.idx", sdxbase);
    gf = gzopen(path, "r");
    if (!gf) die(" Could Not Open the .idx file ");
    gzbuffer(gf, 33554432);
    printf("\n About to read kmers index \n\n");
    gz_read_all(gf, pos_index, words * 4);
    gzclose(gf);
    const uint64_t n_mers = pos_index[words - 1];
    uint32_t *mers = malloc((n_mers + 1) * 4);
    snprintf(path, sizeof path, "%s.mdx", sdxbase);
    FILE *mf = fopen(path, "rb");
    if (!mf || fread(mers, 4, n_mers, mf) != n_mers) die(" Could not read the .mdx file ");
    fclose(mf);
    pemap_index ix = {pos_index, mers, n_mers, genome, genome_size, sdx.starts, sdx.n};
    for (int g = 0; g < n_gpus && !rc; g++) {
      rc = pemap_init(&hs[g], &ix, &prm, device + g);
      if (rc) printf("\n pemap_init failed on GPU %d: %s \n", device + g, pemap_last_error(hs[g]));
    }
    free(pos_index);
    free(mers);
  }
  if (rc) exit(1);

  const long batch_cap = getenv("PEMAP_BATCH") ? atol(getenv("PEMAP_BATCH")) : 1000000;
  uint32_t *maps1 = calloc((size_t)max_reads + 1, 4), *maps2 = calloc((size_t)max_reads + 1, 4);
  if (!maps1 || !maps2) die(" Could not allocate space for mapping position of reads ");
  long tot_pairs = 0;
  worker_t *ws = calloc((size_t)n_gpus, sizeof(worker_t));
  for (int g = 0; g < n_gpus; g++) {
    worker_t *w = &ws[g];
    w->h = hs[g];
    w->paired = paired;
    w->maps1 = maps1;
    w->maps2 = maps2;
    w->max_dist = max_dist;
    pthread_mutex_init(&w->mu, NULL);
    pthread_cond_init(&w->cv, NULL);
    for (int k = 0; k < 3; k++) { /* three slots per GPU: one being decoded into, one queued / copied, one being mapped */
      slot_t *b = slot_new(batch_cap, paired);
      b->next = w->free_list;
      w->free_list = b;
    }
    if (pthread_create(&w->thread, NULL, worker_main, w)) die(" Could not start a submitting thread ");
  }
  long batch_no = 0;
  gzpool_t gz;
  {
    int nt = max_threads > 0 ? max_threads : 8;
    if (nt > 64) nt = 64;
    gzpool_start(&gz, nt, getenv("PEMAP_GZ_LEVEL") ? atoi(getenv("PEMAP_GZ_LEVEL")) : Z_DEFAULT_COMPRESSION);
  }
  out.hs = hs;
  out.n_gpus = n_gpus;
  out.device = device;
  out.paired = paired;
  out.tsw = tsw;
  out.ws = ws;
  out.sdx = &sdx;
  out.genome = genome;
  out.genome_size = genome_size;
  out.gz = &gz;
  int open_flag = 1;
  double decode_seconds = 0, t_map0 = now_s();
  long decoded_reads = 0;

  printf("\n About to start mapping everything \n\n");
  for (int fi = 0; fi < n_files; fi++) {
    decoder_t d1, d2;
    memset(&d1, 0, sizeof d1);
    memset(&d2, 0, sizeof d2);
    reader_open(&d1.r, files1[fi]);
    if (paired) reader_open(&d2.r, files2[fi]);
    if (tsw) { /* tsw:636-675: a new sample name closes the previous sample's outputs and opens its own */
      if (out_names[fi][0] != '\0') {
        if (strcmp(base, out_names[fi]) != 0) {
          open_flag = 1;
          if (fi > 0) {
            for (int g = 0; g < n_gpus; g++) worker_wait_idle(&ws[g]);
            out.tot_pairs = tot_pairs;
            dump_output(&out);
            tot_pairs = 0;
          }
        } else if (fi > 0)
          open_flag = 0;
      }
      if (open_flag) {
        open_flag = 0;
        if (out_names[fi][0] != '\0') strcpy(base, out_names[fi]);
        open_outputs(&out, base);
      }
    }
    d1.trim_start = d2.trim_start = trim_start;
    d1.trim_end = d2.trim_end = trim_end;
    d1.check_short = 1; /* 663: only the first file's read length ends the input */
    decoder_prime(&d1);
    if (paired) decoder_prime(&d2);
    long current = 0;
    int go = 1;
    while (go) {
      worker_t *w = &ws[batch_no % n_gpus];
      slot_t *b = worker_take_free(w);
      const long want = max_reads - current < batch_cap ? max_reads - current : batch_cap;
      const double t0 = now_s();
      d1.rows = b->rows1;
      d1.len = b->len1;
      d1.want = want;
      d2.rows = b->rows2;
      d2.len = b->len2;
      d2.want = want;
      pthread_t t2;
      if (paired && pthread_create(&t2, NULL, decoder_fill, &d2)) die(" Could not start a decoder thread ");
      decoder_fill(&d1);
      if (paired) pthread_join(t2, NULL);
      decode_seconds += now_s() - t0;
      long filled = d1.got;
      if (paired && d2.got < filled) filled = d2.got; /* 663-700: a pair exists while both files have a read */
      if (filled < want) go = 0;                        /* one of the files ended, or a read of <= 12 bases (663) */
      current += filled;
      decoded_reads += filled * (paired ? 2 : 1);
      if (current >= max_reads) go = 0;
      if (filled > 0) {
        b->filled = filled;
        b->first = current - filled;
        worker_submit(w, b);
        printf("\n We have read %ld reads \n\n", current);
        batch_no++;
      } else { /* nothing in this slot: hand it back */
        pthread_mutex_lock(&w->mu);
        b->next = w->free_list;
        w->free_list = b;
        pthread_mutex_unlock(&w->mu);
      }
    }
    for (int g = 0; g < n_gpus; g++) worker_wait_idle(&ws[g]); /* the .mfile arrays are complete (768-774) */
    printf("\n Made it out alive, and have started cleanup \n\n");
    snprintf(path, sizeof path, "%s.mfile", files1[fi]); /* 775-781 */
    FILE *mf = fopen(path, "wb");
    if (!mf) die(" Can not open the .mfile for writing ");
    fwrite(maps1, 4, (size_t)current, mf);
    fclose(mf);
    if (paired) {
      snprintf(path, sizeof path, "%s.mfile", files2[fi]);
      mf = fopen(path, "wb");
      if (!mf) die(" Can not open the .mfile for writing ");
      fwrite(maps2, 4, (size_t)current, mf);
      fclose(mf);
      reader_close(&d2.r);
    }
    reader_close(&d1.r);
    tot_pairs += current;
  }
  const double map_wall = now_s() - t_map0;

  for (int g = 0; g < n_gpus; g++) worker_wait_idle(&ws[g]);
  out.tot_pairs = tot_pairs;
  const double t_dump0 = now_s();
  const int nothing = dump_output(&out); /* 786-900 / tsw:847 */
  const double out_wall = now_s() - t_dump0;
  const double t_out0 = now_s();
  (void)t_out0;
  for (int g = 0; g < n_gpus; g++) {
    worker_quit(&ws[g]);
    pthread_join(ws[g].thread, NULL);
  }
  if (getenv("PEMAP_TIMING") && atoi(getenv("PEMAP_TIMING"))) {
    double map_s = 0;
    for (int g = 0; g < n_gpus; g++) map_s += ws[g].map_seconds;
    fprintf(stderr,
            "{\"read_mates\": %ld, \"decode_and_map_wall_s\": %.3f, \"decode_s\": %.3f, \"decode_reads_per_s\": %.0f, "
            "\"gpu_map_s\": %.3f, \"gpu_map_reads_per_s\": %.0f, \"fastq_to_mapped_reads_per_s\": %.0f, "
            "\"pileup_records\": %llu, \"gz_threads\": %d, \"gz_level\": %d, \"gz_s\": %.3f, \"gz_in_GBs\": %.3f, \"gz_ratio\": %.2f, "
            "\"output_wall_s\": %.3f}\n",
            decoded_reads, map_wall, decode_seconds, decoded_reads / (decode_seconds > 0 ? decode_seconds : 1e-9), map_s,
            decoded_reads / (map_s > 0 ? map_s : 1e-9), decoded_reads / (map_wall > 0 ? map_wall : 1e-9),
            (unsigned long long)out.n_records, gz.n_threads, gz.level, gz.seconds,
            gz.bytes_in / 1e9 / (gz.seconds > 0 ? gz.seconds : 1e-9), gz.bytes_out ? (double)gz.bytes_in / (double)gz.bytes_out : 0.0,
            out_wall);
  }
  gzpool_stop(&gz);
  for (int g = 0; g < n_gpus; g++) pemap_destroy(hs[g]);
  return (nothing && !tsw) ? 1 : 0;
}
