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
 * max_threads is accepted and ignored: the reader thread fills batches, one submitting thread per GPU maps them.
 * Environment:
 *   PEMAP_GPUS          number of GPUs of this box to use (default 1): batch b goes to GPU b mod N, every GPU keeps
 *                       its own counters, and pemap_reduce_counts_peer sums them onto GPU 0 over NVLink at the end
 *   PEMAP_DEVICE        first GPU ordinal (default 0)
 *   PEMAP_DEVICE_INDEX  1 = rebuild pos_index/mers on the GPU from .seq/.sdx instead of loading .idx/.mdx
 *   PEMAP_BATCH         reads per pemap_map_batch_rows call (default 1,000,000; results do not depend on it)
 * Written from scratch; what must be byte-compatible (file formats, summary text) cites the reference line.
 */
#include <ctype.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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

/* ---- one submitting thread per GPU (replaces the reference's pool of map_everything threads, 677-702) ---- */
typedef struct {
  pemap_t *h;
  int paired;
  long cap, filled, first; /* batch capacity, reads in the batch, index of its first read in the file */
  char *rows1, *rows2;
  int *len1, *len2;
  uint32_t *bm1, *bm2;
  int *btype;
  uint32_t *maps1, *maps2;       /* shared per-file arrays; batches write disjoint ranges */
  long max_dist;
  long mate_counts[9], total_reads, total_bases, total_dist, no_dists; /* per worker, summed at the end */
  pthread_t thread;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int state; /* 0 idle (buffers free), 1 batch ready, 2 quit */
} worker_t;

static void worker_epilogue(worker_t *w) { /* batch epilogue, 1238-1265 */
  for (long j = 0; j < w->filled; j++) {
    w->mate_counts[w->btype[j]]++;
    w->maps1[w->first + j] = w->bm1[j];
    if (w->bm1[j]) {
      w->total_reads++;
      w->total_bases += w->len1[j];
      if (w->bm2[j]) {
        w->total_reads++;
        w->total_bases += w->len2[j];
        long test = (long)(uint32_t)(w->bm1[j] - w->bm2[j]); /* labs() of an unsigned difference (1250) */
        w->maps2[w->first + j] = w->bm2[j];
        if (test < w->max_dist * 4) {
          w->total_dist += test;
          w->no_dists++;
        }
      }
    } else if (w->bm2[j]) {
      w->total_reads++;
      w->total_bases += w->len2[j];
      w->maps2[w->first + j] = w->bm2[j];
    }
  }
}

static void *worker_main(void *arg) {
  worker_t *w = arg;
  for (;;) {
    pthread_mutex_lock(&w->mu);
    while (w->state == 0) pthread_cond_wait(&w->cv, &w->mu);
    const int st = w->state;
    pthread_mutex_unlock(&w->mu);
    if (st == 2) return NULL;
    int rc = pemap_map_batch_rows(w->h, (int)w->filled, w->rows1, w->len1, w->rows2, w->paired ? w->len2 : NULL, ROW, w->bm1,
                                  w->bm2, w->btype);
    if (rc) {
      printf("\n pemap_map_batch failed: %s \n", pemap_last_error(w->h));
      exit(1);
    }
    worker_epilogue(w);
    pthread_mutex_lock(&w->mu);
    w->state = 0;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
  }
}

static void worker_wait_idle(worker_t *w) {
  pthread_mutex_lock(&w->mu);
  while (w->state != 0) pthread_cond_wait(&w->cv, &w->mu);
  pthread_mutex_unlock(&w->mu);
}

static void worker_submit(worker_t *w, int state) {
  pthread_mutex_lock(&w->mu);
  w->state = state;
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

/* ---- outputs: the writer loop of main() (819-900) and, for the tsw form, dump_output (tsw:849-965) ---- */
typedef struct {
  gzFile pile, indel;
  FILE *summary;
  pemap_t **hs;
  int n_gpus, device, paired, tsw;
  worker_t *ws;
  const sdx_t *sdx;
  const char *genome;
  uint64_t genome_size;
  long mate_counts[9], total_reads, total_bases, total_dist, no_dists, tot_pairs;
} out_t;

static void open_outputs(out_t *o, const char *base) {
  char path[4300];
  snprintf(path, sizeof path, "%s.pileup.gz", base);
  o->pile = gzopen(path, "wb");
  if (!o->pile) die(" Can not open the pileup file for writing ");
  gzbuffer(o->pile, 33554432);
  snprintf(path, sizeof path, "%s.indel.txt.gz", base);
  o->indel = gzopen(path, "w");
  if (!o->indel) die(" Can not open the indel file for writing ");
  gzbuffer(o->indel, 33554432);
  snprintf(path, sizeof path, "%s.summary.txt", base);
  o->summary = fopen(path, "w");
  if (!o->summary) die(" Can not open the summary file for writing ");
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

  pemap_t *h = o->hs[0];
  const pemap_record *rec;
  const pemap_insertion *ins;
  uint64_t n_rec, n_ins;
  pemap_insertion *all_ins = NULL; /* insertion strings of every GPU, sorted by site */
  uint64_t n_all = 0;
  int rc;
  for (int g = 1; g < o->n_gpus; g++) { /* every GPU's counters onto GPU 0 over NVLink; its insertion strings to the host */
    const pemap_record *r2;
    const pemap_insertion *i2;
    uint64_t nr2, ni2;
    rc = pemap_reduce_counts_peer(o->hs[0], o->hs[g]);
    if (rc) {
      printf("\n reducing GPU %d failed: %s \n", o->device + g, pemap_last_error(o->hs[0]));
      exit(1);
    }
    rc = pemap_finish(o->hs[g], &r2, &nr2, &i2, &ni2);
    if (rc) {
      printf("\n pemap_finish failed on GPU %d: %s \n", o->device + g, pemap_last_error(o->hs[g]));
      exit(1);
    }
    all_ins = realloc(all_ins, (size_t)(n_all + ni2 + 1) * sizeof(pemap_insertion));
    memcpy(all_ins + n_all, i2, (size_t)ni2 * sizeof(pemap_insertion));
    n_all += ni2;
  }
  rc = pemap_finish(h, &rec, &n_rec, &ins, &n_ins);
  if (rc) {
    printf("\n pemap_finish failed: %s \n", pemap_last_error(h));
    exit(1);
  }
  if (o->n_gpus > 1) {
    all_ins = realloc(all_ins, (size_t)(n_all + n_ins + 1) * sizeof(pemap_insertion));
    memcpy(all_ins + n_all, ins, (size_t)n_ins * sizeof(pemap_insertion));
    n_all += n_ins;
    qsort(all_ins, (size_t)n_all, sizeof(pemap_insertion), cmp_ins);
    ins = all_ins;
    n_ins = n_all;
  }
  const sdx_t *sdx = o->sdx;
  gzprintf(o->indel, "Fragment\tPositions\tReference Base\tTotal Coverage\tReference Reads\tNo Deletions\tNo Insertions\tInsertion Sequence"); /* 819-820 */
  uint32_t *padded = calloc((size_t)sdx->n + 16, 4);
  for (int i = 0; i <= sdx->n; i++) padded[i] = sdx->starts[i] + 15u * (uint32_t)i; /* 821-822 */
  uint64_t q = 0;
  for (uint64_t k = 0; k < n_rec; k++) { /* 828-864 */
    gzwrite(o->pile, &rec[k].pos, 4);
    gzwrite(o->pile, rec[k].c, 12);
    if (rec[k].c[5] > 0) {
      const uint32_t pos = rec[k].pos;
      const char ref = o->genome[pos];
      const int tot = rec[k].c[0] + rec[k].c[1] + rec[k].c[2] + rec[k].c[3] + rec[k].c[4] + rec[k].c[5];
      const int ref_reads = ref == 'A' ? rec[k].c[0] : ref == 'C' ? rec[k].c[1] : ref == 'G' ? rec[k].c[2] : rec[k].c[3];
      const int which = find_contig(padded, sdx->n, pos);
      gzprintf(o->indel, "\n%s\t%d\t%c\t%d\t%d\t%d\t%d", sdx->names[which], (int)(1 + pos - padded[which]), ref, tot,
               ref_reads, rec[k].c[4], rec[k].c[5]);
      while (q < n_ins && ins[q].pos < pos) q++;
      for (; q < n_ins && ins[q].pos == pos; q++) gzprintf(o->indel, "\t%s", ins[q].seq);
    }
  }
  free(padded);
  free(all_ins);
  gzclose(o->pile);
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

int main(int argc, char **argv) {
  if (argc < 4) die("Usage: pemapper_gpu out_file sdx_file [s,sa,p,pa] ... (same arguments as pemapper)");
  const char mode = (char)toupper(argv[3][0]), arr = (char)toupper(argv[3][1]);
  int paired, max_dist = 0, min_dist = 0, bisulfite;
  double min_align;
  long max_reads;
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
  } else { /* init_index_buffer 2129-2155 */
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
    w->cap = batch_cap;
    w->rows1 = malloc((size_t)batch_cap * ROW);
    w->rows2 = paired ? malloc((size_t)batch_cap * ROW) : NULL;
    w->len1 = malloc(sizeof(int) * (size_t)batch_cap);
    w->len2 = malloc(sizeof(int) * (size_t)batch_cap);
    w->bm1 = malloc(4 * (size_t)batch_cap);
    w->bm2 = malloc(4 * (size_t)batch_cap);
    w->btype = malloc(sizeof(int) * (size_t)batch_cap);
    if (!w->rows1 || (paired && !w->rows2) || !w->len1 || !w->len2 || !w->bm1 || !w->bm2 || !w->btype)
      die(" Could not allocate the batch buffers ");
    w->maps1 = maps1;
    w->maps2 = maps2;
    w->max_dist = max_dist;
    pthread_mutex_init(&w->mu, NULL);
    pthread_cond_init(&w->cv, NULL);
    w->state = 0;
    if (pthread_create(&w->thread, NULL, worker_main, w)) die(" Could not start a submitting thread ");
  }
  long batch_no = 0;
  out.hs = hs;
  out.n_gpus = n_gpus;
  out.device = device;
  out.paired = paired;
  out.tsw = tsw;
  out.ws = ws;
  out.sdx = &sdx;
  out.genome = genome;
  out.genome_size = genome_size;
  int open_flag = 1;

  printf("\n About to start mapping everything \n\n");
  for (int fi = 0; fi < n_files; fi++) {
    reader r1, r2;
    reader_open(&r1, files1[fi]);
    if (paired) reader_open(&r2, files2[fi]);
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
    char *s1 = reader_line(&r1), *s2 = NULL;
    s1 = trim_read(reader_line(&r1), trim_start, trim_end);
    if (paired) {
      s2 = reader_line(&r2);
      s2 = trim_read(reader_line(&r2), trim_start, trim_end);
    }
    long current = 0, filled = 0, batch_first = 0;
    int go = s1 != NULL;
    worker_t *w = &ws[batch_no % n_gpus];
    worker_wait_idle(w);
    while (go || filled) {
      char *rows1 = w->rows1, *rows2 = w->rows2;
      int *len1 = w->len1, *len2 = w->len2;
      const int have = go && s1 && (int)strlen(s1) > 12 && (!paired || s2); /* 663 */
      if (have) {
        const int l1 = (int)strlen(s1), l2 = paired ? (int)strlen(s2) : 0;
        if (l1 > PEMAP_MAX_READ + 20 || l2 > PEMAP_MAX_READ + 20) die(" Read longer than the reference's DP buffers allow ");
        memcpy(rows1 + (size_t)filled * ROW, s1, (size_t)l1);
        len1[filled] = l1;
        if (paired) {
          memcpy(rows2 + (size_t)filled * ROW, s2, (size_t)l2);
          len2[filled] = l2;
        }
        filled++;
        current++;
        if (current >= max_reads) go = 0;
        else {
          s1 = trim_read(next_sequence(&r1), trim_start, trim_end);
          if (!s1) go = 0;
          if (paired && go) {
            s2 = trim_read(next_sequence(&r2), trim_start, trim_end);
            if (!s2) go = 0;
          }
        }
      } else
        go = 0;
      if (filled == batch_cap || (!go && filled)) {
        w->filled = filled;
        w->first = batch_first;
        worker_submit(w, 1);
        batch_first += filled;
        filled = 0;
        printf("\n We have read %ld reads \n\n", current);
        batch_no++;
        w = &ws[batch_no % n_gpus];
        worker_wait_idle(w); /* its previous batch has been mapped and booked: the buffers are free */
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
      reader_close(&r2);
    }
    reader_close(&r1);
    tot_pairs += current;
  }

  for (int g = 0; g < n_gpus; g++) worker_wait_idle(&ws[g]);
  out.tot_pairs = tot_pairs;
  const int nothing = dump_output(&out); /* 786-900 / tsw:847 */
  for (int g = 0; g < n_gpus; g++) {
    worker_submit(&ws[g], 2);
    pthread_join(ws[g].thread, NULL);
  }
  for (int g = 0; g < n_gpus; g++) pemap_destroy(hs[g]);
  return (nothing && !tsw) ? 1 : 0;
}
