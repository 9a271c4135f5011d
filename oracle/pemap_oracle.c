/*
 * pemap_oracle.c - TEST INFRASTRUCTURE ONLY (see pemap_oracle.h for the rules and parity status).
 *
 * A from-scratch CPU restatement of the PEMapper hot path.  Each function cites the lines of
 * /root/reference/src/pemapper.c (or index_genome_whole.c) whose behaviour it restates, including
 * the quirks that are part of the bit-exact contract.  All score arithmetic is
 * IEEE double evaluated with the reference's own expressions; build with -ffp-contract=off.
 */
#include "pemap_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

enum { T_UNIQUE_MATE = 0, T_UNIQUE_SLIP, T_UNIQUE_SINGLE, T_UNIQUE_MIS, T_NON_MATE, T_NON_MIS, T_FRAG_MIS, T_NON_NO,
       T_NEITHER_MAP }; /* pemapper.c:37-45 */

#define DP_DIM 301          /* reference buffers are 300x300 (pemapper.c:916, 932-956); one spare row/col */
#define MAX_SEG 20          /* total_cuts+1 <= 299/16+1 */
#define SEG_CAP 5002        /* max_mers = too_many_spots*50 (pemapper.c:1544), +count slot */
#define KV 49               /* exact 16-mer + 48 one-substitution neighbours */

typedef struct ins_rec {
  uint32_t pos;
  uint32_t len;
  uint64_t off;
} ins_rec;

struct orc_ctx {
  orc_params p;
  int n_contigs;
  int64_t genome_size;
  char *genome;
  uint32_t *cstart; /* unpadded prefix sums of (len-15), n+1 entries */
  uint64_t n_mers;
  uint32_t *mers;
  uint32_t n_dk;
  uint32_t *dk, *dk_start, *bucket;
  uint16_t (*cnt)[6];
  pthread_mutex_t ins_mu;
  ins_rec *ins;
  uint64_t n_ins, cap_ins;
  char *ins_pool;
  uint64_t pool_len, pool_cap;
  double border[DP_DIM]; /* S*[0][j], pemapper.c:2073-2081 */
  double go, ge, match, mism;
  double bonus[128][128]; /* pemapper.c:2006-2035 */
  int mismatch[256][12];  /* pemapper.c:546-565 */
  uint64_t cells;
};

typedef struct scratch {
  double (*S)[DP_DIM][DP_DIM]; /* S[3][i][j] */
  uint32_t (*seg[2])[SEG_CAP]; /* [strand][segment][0]=count, then positions */
} scratch;

void orc_default_params(orc_params *p) {
  p->idepth = 16;
  p->max_hits = 200;
  p->too_many_spots = 100;
  p->min_align = 0.9;
  p->match_bonus = 1.0;
  p->is_bisulfite = 0;
  p->pair_flag = 0;
  p->min_dist = 0;
  p->max_dist = 500;
  p->misalign_slop = 10;
}

/* ---------------------------------------------------------------- tables */

static void build_tables(orc_ctx *c) {
  const double mb = c->p.match_bonus;
  c->match = mb;
  c->mism = -1.0 / ((double)3.0 * mb); /* pemapper.c:2011 */
  c->go = 2.0 * mb;                    /* pemapper.c:2039 */
  c->ge = mb / 36.0;                   /* pemapper.c:2040 */
  for (int a = 0; a < 128; a++)
    for (int b = 0; b < 128; b++) c->bonus[a][b] = (a == b) ? c->match : c->mism;
  for (int a = 0; a < 128; a++) c->bonus[a]['N'] = c->bonus['N'][a] = c->bonus[a]['n'] = c->bonus['n'][a] = c->match;
  if (c->p.is_bisulfite) /* pemapper.c:2024-2034 */
    c->bonus['C']['T'] = c->bonus['C']['t'] = c->bonus['c']['T'] = c->bonus['c']['t'] = c->match;
  c->border[0] = 0.0;
  for (int j = 1; j < DP_DIM; j++) c->border[j] = -(c->go + (double)(j - 1) * c->ge); /* pemapper.c:2077-2078 */
  /* one-substitution table of a packed byte (4 bases): 3 alternatives for each 2-bit field, low field first */
  for (int v = 0; v < 256; v++) {
    int w = 0;
    for (int f = 0; f < 4; f++) {
      int cur = (v >> (2 * f)) & 3, rest = v & ~(3 << (2 * f));
      for (int k = 0; k < 4; k++)
        if (k != cur) c->mismatch[v][w++] = rest | (k << (2 * f));
    }
  }
}

/* ---------------------------------------------------------------- index (index_genome_whole.c) */

static inline uint32_t P_of(const orc_ctx *c, uint64_t x) { /* number of indexed positions with k-mer < x */
  if (x >= ((uint64_t)1 << 32)) return (uint32_t)c->n_mers;
  uint32_t key = (uint32_t)x, b = key >> 8;
  uint32_t lo = c->bucket[b], hi = c->bucket[b + 1];
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (c->dk[mid] < key) lo = mid + 1; else hi = mid;
  }
  return c->dk_start[lo];
}

static void radix_sort_pairs(uint32_t *key, uint32_t *val, uint64_t n) {
  uint32_t *k2 = malloc(n * 4 + 4), *v2 = malloc(n * 4 + 4);
  for (int pass = 0; pass < 4; pass++) {
    uint64_t hist[257] = {0};
    int sh = 8 * pass;
    for (uint64_t i = 0; i < n; i++) hist[((key[i] >> sh) & 255) + 1]++;
    for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
    for (uint64_t i = 0; i < n; i++) {
      uint64_t d = hist[(key[i] >> sh) & 255]++;
      k2[d] = key[i];
      v2[d] = val[i];
    }
    uint32_t *t = key; key = k2; k2 = t;
    t = val; val = v2; v2 = t;
  } /* 4 passes: data is back in the caller's arrays */
  free(k2);
  free(v2);
}

static void build_index(orc_ctx *c, const int64_t *contig_len) {
  /* index_genome_whole.c:169-177 (2-bit code; C==T when bisulfite), 248-299 (rolling 32-bit code,
     N resets, position = gpos + contig offset of the k-mer start), 213-216 (gpos += len-15). */
  unsigned bit[256] = {0};
  bit['G'] = 2;
  bit['T'] = 3;
  bit['C'] = c->p.is_bisulfite ? 3 : 1;
  const int k = c->p.idepth;
  uint64_t cap = (uint64_t)c->genome_size + 1, n = 0;
  uint32_t *key = malloc(cap * 4), *val = malloc(cap * 4);
  uint32_t gpos = 0;
  int64_t base = 0;
  c->cstart[0] = 0;
  for (int ci = 0; ci < c->n_contigs; ci++) {
    uint32_t code = 0;
    int run = 0;
    for (int64_t q = 0; q < contig_len[ci]; q++) {
      unsigned char ch = (unsigned char)c->genome[base + q];
      if (ch == 'N') {
        run = 0;
        code = 0;
        continue;
      }
      code = (code << 2) + bit[ch];
      if (++run >= k) {
        key[n] = code;
        val[n] = gpos + (uint32_t)(q + 1 - k);
        n++;
      }
    }
    gpos += (uint32_t)(contig_len[ci] - (k - 1));
    c->cstart[ci + 1] = gpos;
    base += contig_len[ci];
  }
  radix_sort_pairs(key, val, n); /* stable: positions stay ascending inside a k-mer (291-292, 339-340) */
  c->n_mers = n;
  c->mers = val;
  uint32_t nd = 0;
  for (uint64_t i = 0; i < n; i++)
    if (i == 0 || key[i] != key[i - 1]) nd++;
  c->n_dk = nd;
  c->dk = malloc((uint64_t)(nd + 1) * 4);
  c->dk_start = malloc((uint64_t)(nd + 1) * 4);
  nd = 0;
  for (uint64_t i = 0; i < n; i++)
    if (i == 0 || key[i] != key[i - 1]) {
      c->dk[nd] = key[i];
      c->dk_start[nd] = (uint32_t)i;
      nd++;
    }
  c->dk_start[nd] = (uint32_t)n;
  free(key);
  const uint32_t nb = 1u << 24;
  c->bucket = malloc((uint64_t)(nb + 1) * 4);
  uint32_t j = 0;
  for (uint32_t b = 0; b <= nb; b++) {
    uint64_t lim = (uint64_t)b << 8;
    while (j < nd && c->dk[j] < lim) j++;
    c->bucket[b] = j;
  }
}

orc_ctx *orc_create(const char *genome, const int64_t *contig_len, int n_contigs, const orc_params *p) {
  orc_ctx *c = calloc(1, sizeof(*c));
  c->p = *p;
  c->n_contigs = n_contigs;
  for (int i = 0; i < n_contigs; i++) c->genome_size += contig_len[i];
  c->genome = malloc(c->genome_size + 1);
  memcpy(c->genome, genome, c->genome_size);
  c->genome[c->genome_size] = 0;
  c->cstart = calloc(n_contigs + 16, 4); /* slack: find_chrom reads [7],[8] (pemapper.c:2175) */
  build_tables(c);
  build_index(c, contig_len);
  c->cnt = calloc(c->genome_size + 1, sizeof(*c->cnt));
  pthread_mutex_init(&c->ins_mu, NULL);
  return c;
}

void orc_set_params(orc_ctx *c, const orc_params *p) {
  int bis = c->p.is_bisulfite;
  c->p = *p;
  c->p.is_bisulfite = bis | p->is_bisulfite; /* the index encoding cannot change after the build */
  build_tables(c);
}

void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->genome); free(c->cstart); free(c->mers); free(c->dk); free(c->dk_start); free(c->bucket);
  free(c->cnt); free(c->ins); free(c->ins_pool);
  free(c);
}

uint64_t orc_n_mers(const orc_ctx *c) { return c->n_mers; }
const uint32_t *orc_mers(const orc_ctx *c) { return c->mers; }
uint32_t orc_pos_index(const orc_ctx *c, uint64_t w) { return P_of(c, w); }
int64_t orc_genome_size(const orc_ctx *c) { return c->genome_size; }
const uint32_t *orc_contig_starts(const orc_ctx *c) { return c->cstart; }
uint64_t orc_cells(const orc_ctx *c) { return c->cells; }

void orc_fill_pos_index(const orc_ctx *c, uint32_t *table) { /* index_genome_whole.c:334-342 */
  uint64_t w = 0;
  for (uint32_t d = 0; d < c->n_dk; d++) {
    uint64_t upto = c->dk[d]; /* entries (prev kmer, this kmer] hold dk_start[d] */
    uint32_t v = c->dk_start[d];
    for (; w <= upto; w++) table[w] = v;
  }
  for (; w <= ((uint64_t)1 << 32); w++) table[w] = (uint32_t)c->n_mers;
}

int orc_write_index(const orc_ctx *c, const char *base, const char *const *names, int with_idx) {
  char path[4096];
  snprintf(path, sizeof path, "%s.sdx", base);
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "%d\n", c->n_contigs); /* index_genome_whole.c:347-350 */
  for (int i = 0; i < c->n_contigs; i++) fprintf(f, "%d\t%s\n", (int)(c->cstart[i + 1] - c->cstart[i]), names[i]);
  fprintf(f, "%d\n", c->p.idepth);
  fclose(f);
  snprintf(path, sizeof path, "%s.seq", base);
  gzFile g = gzopen(path, "w");
  if (!g) return -1;
  for (int64_t o = 0; o < c->genome_size; o += 1 << 30) {
    int64_t m = c->genome_size - o;
    gzwrite(g, c->genome + o, (unsigned)(m < (1 << 30) ? m : (1 << 30)));
  }
  gzclose(g);
  snprintf(path, sizeof path, "%s.mdx", base);
  f = fopen(path, "wb");
  if (!f) return -1;
  fwrite(c->mers, 4, c->n_mers, f);
  fclose(f);
  if (with_idx) {
    snprintf(path, sizeof path, "%s.idx", base);
    g = gzopen(path, "wb1");
    if (!g) return -1;
    const uint64_t chunk = 1u << 24;
    uint32_t *buf = malloc(chunk * 4);
    uint64_t w = 0, total = ((uint64_t)1 << 32) + 1;
    uint32_t d = 0;
    while (w < total) {
      uint64_t m = total - w < chunk ? total - w : chunk;
      for (uint64_t i = 0; i < m; i++) {
        uint64_t x = w + i;
        while (d < c->n_dk && c->dk[d] < x) d++;
        buf[i] = (d < c->n_dk) ? c->dk_start[d] : (uint32_t)c->n_mers;
      }
      gzwrite(g, buf, (unsigned)(m * 4));
      w += m;
    }
    free(buf);
    gzclose(g);
  }
  return 0;
}

/* ---------------------------------------------------------------- seeding (initial_map) */

static void reverse_transcribe(const char *s, char *out, int n) { /* pemapper.c:2303-2337 */
  for (int i = n - 1; i >= 0; i--) {
    char ch;
    switch (s[i]) {
      case 'A': ch = 'T'; break;
      case 'C': ch = 'G'; break;
      case 'G': ch = 'C'; break;
      case 'T': ch = 'A'; break;
      case 'W': ch = 'W'; break;
      case 'S': ch = 'S'; break;
      case 'K': ch = 'M'; break;
      case 'M': ch = 'K'; break;
      case 'Y': ch = 'R'; break;
      case 'R': ch = 'Y'; break;
      default: ch = 'N';
    }
    *out++ = ch;
  }
  *out = 0;
}

static inline unsigned base_code(char ch) { /* fill_cv_mat, pemapper.c:2379-2383: N and everything else = A */
  switch (ch) {
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default: return 0;
  }
}

static uint32_t pack16(const char *s) { /* convert_seq_int, pemapper.c:2408-2423 */
  uint32_t w = 0;
  for (int i = 0; i < 16; i++) w = (w << 2) | base_code(s[i]);
  return w;
}

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

/* one strand: k-mer variants (fill_mers 1969-2003), lookups (get_mers 2158-2165), the >=too_many veto and
   the sort (1594-1616) */
static void gather_strand(const orc_ctx *c, const char *seq, int total_cuts, const int *offsets,
                          uint32_t (*seg)[SEG_CAP]) {
  for (int s = 0; s <= total_cuts; s++) {
    uint32_t kv[KV], exact = pack16(seq + offsets[s]);
    int m = 0;
    kv[m++] = exact;
    for (int byte = 0; byte < 4; byte++) {
      uint32_t field = (exact >> (8 * byte)) & 255, rest = exact - (field << (8 * byte));
      for (int k = 0; k < 12; k++) kv[m++] = rest + ((uint32_t)c->mismatch[field][k] << (8 * byte));
    }
    uint32_t *lst = seg[s];
    lst[0] = 0;
    for (int j = 0; j < KV; j++) {
      uint32_t w = kv[j];
      uint32_t lo = P_of(c, w), hi = P_of(c, (uint32_t)(w + 1)); /* which+1 wraps in 32 bits (2163) */
      uint32_t cnt = hi - lo;
      if (cnt >= (uint32_t)c->p.too_many_spots) {
        lst[0] = 0;
        break;
      }
      memcpy(lst + 1 + lst[0], c->mers + lo, (size_t)cnt * 4);
      lst[0] += cnt;
    }
    if (lst[0] > 1) qsort(lst + 1, lst[0], 4, cmp_u32);
  }
}

typedef struct hitlist {
  uint32_t hits[256];
  int hits_off[256];
  char orient[256];
  int tot;
  int min_match;
} hitlist;

/* find_matches, pemapper.c:2189-2289, restated with its window/cursor mechanics */
static void chain_strand(const orc_ctx *c, uint32_t (*seg)[SEG_CAP], int max_depth, const int *offsets, hitlist *h,
                         char strand) {
  const uint32_t max_off = (uint32_t)(c->p.idepth - 4 > 2 ? c->p.idepth - 4 : 2);
  const int max_hits = c->p.max_hits;
  uint32_t cursor[MAX_SEG];
  uint32_t min_spots = 10000;
  for (int s = 0; s <= max_depth; s++)
    if (seg[s][0] < min_spots) min_spots = seg[s][0];
  if (min_spots > (uint32_t)max_hits) { /* 2203-2207: also wipes what the forward strand found */
    h->tot = 0;
    return;
  }
  for (uint32_t loop = 0; loop <= (uint32_t)(1 + max_depth - h->min_match); loop++) {
    long lo_d = -((long)offsets[loop] + (long)max_off);
    long hi_d = max_off;
    for (int j = (int)loop + 1; j <= max_depth; j++) {
      long e = (long)max_off + offsets[j] - offsets[loop];
      if (e > hi_d) hi_d = e;
    }
    for (int s = (int)loop; s <= max_depth; s++) cursor[s] = 1;
    const uint32_t *anchor = seg[loop];
    for (uint32_t i = 1; i <= anchor[0]; i++) {
      long w_lo = (long)anchor[i] + lo_d, w_hi = (long)anchor[i] + hi_d;
      if (w_lo < 0) w_lo = 0;
      if (w_hi < 0) w_hi = 0;
      for (int j = (int)loop + 1; j <= max_depth; j++)
        while (cursor[j] < seg[j][0] && (long)seg[j][cursor[j]] < w_lo) cursor[j]++;
      uint32_t found = 1;
      for (int j = (int)loop + 1; j <= max_depth; j++)
        for (uint32_t k = cursor[j]; k <= seg[j][0] && (long)seg[j][k] <= w_hi; k++) {
          int32_t d = (int32_t)((uint32_t)(anchor[i] - seg[j][k]) - (uint32_t)(offsets[loop] - offsets[j]));
          if ((uint32_t)abs(d) < max_off) {
            found++;
            break;
          }
        }
      if (found > (uint32_t)h->min_match) { /* 2251-2260: better chain resets the list */
        h->min_match = (int)found;
        h->tot = 0;
        h->hits[0] = anchor[i];
        h->hits_off[0] = offsets[loop];
        h->orient[0] = strand;
        h->tot = 1;
      } else if (found == (uint32_t)h->min_match) {
        if (h->tot < max_hits) { /* 2264-2282: dedup on pos-offset across both strands */
          int is_new = 1;
          for (int k = 0; k < h->tot; k++)
            if (h->hits[k] - (uint32_t)h->hits_off[k] == anchor[i] - (uint32_t)offsets[loop]) {
              is_new = 0;
              break;
            }
          if (is_new) {
            h->hits[h->tot] = anchor[i];
            h->hits_off[h->tot] = offsets[loop];
            h->orient[h->tot] = strand;
            h->tot++;
          }
        } else
          return; /* 2283-2284 */
      }
    }
  }
}

static int seed_read(const orc_ctx *c, scratch *sc, const char *fwd_in, const char *rev_in, int len, uint32_t *spots,
                     char *orients) {
  /* initial_map, pemapper.c:1539-1690 */
  int n_count = 0;
  for (int i = 0; i < len; i++) n_count += (fwd_in[i] == 'N');
  if (n_count >= 1 + len / 10) return 0; /* 1552-1559 */
  char fwd[DP_DIM + 16], rev[DP_DIM + 16];
  memset(fwd, 0, sizeof fwd);
  memset(rev, 0, sizeof rev);
  memcpy(fwd, fwd_in, len);
  memcpy(rev, rev_in, len);
  if (c->p.is_bisulfite) /* convert_ct 2292-2300 */
    for (int i = 0; i < len; i++) {
      if (fwd[i] == 'C') fwd[i] = 'T';
      if (rev[i] == 'C') rev[i] = 'T';
    }
  const int k = c->p.idepth;
  int total_cuts = len / k; /* 1573-1587 */
  if (len % k == 0) total_cuts--;
  int offsets[MAX_SEG];
  offsets[0] = 0;
  int i = 1;
  for (; i < total_cuts; i++) offsets[i] = offsets[i - 1] + k;
  if (i == total_cuts) offsets[i] = len - k;
  gather_strand(c, fwd, total_cuts, offsets, sc->seg[0]);
  gather_strand(c, rev, total_cuts, offsets, sc->seg[1]);
  hitlist h;
  h.tot = 0;
  h.min_match = total_cuts > 1 ? total_cuts : 1; /* 1642-1645 */
  if (total_cuts > 4) h.min_match = (4 * total_cuts) / 5;
  if (h.min_match > 4) h.min_match = 4;
  chain_strand(c, sc->seg[0], total_cuts, offsets, &h, 0);
  if (h.tot < c->p.max_hits) chain_strand(c, sc->seg[1], total_cuts, offsets, &h, 1); /* 1657-1660 */
  for (int q = 0; q < h.tot; q++) {
    long t = (long)h.hits[q] - (long)h.hits_off[q]; /* 1664-1669 */
    spots[q] = (uint32_t)(t > 0 ? t : 0);
    orients[q] = h.orient[q];
  }
  return h.tot;
}

static scratch *scratch_new(void) {
  scratch *sc = malloc(sizeof *sc);
  sc->S = malloc(sizeof(double) * 3 * DP_DIM * DP_DIM);
  sc->seg[0] = malloc(sizeof(uint32_t) * MAX_SEG * SEG_CAP);
  sc->seg[1] = malloc(sizeof(uint32_t) * MAX_SEG * SEG_CAP);
  return sc;
}
static void scratch_free(scratch *sc) {
  free(sc->S); free(sc->seg[0]); free(sc->seg[1]); free(sc);
}

int orc_initial_map(const orc_ctx *c, const char *read, int len, uint32_t *spots, char *orients) {
  scratch *sc = scratch_new();
  char rev[DP_DIM + 16];
  reverse_transcribe(read, rev, len);
  int n = seed_read(c, sc, read, rev, len, spots, orients);
  scratch_free(sc);
  return n;
}

/* ---------------------------------------------------------------- windows */

static int find_chrom(const uint32_t *pos, int first, int last, int probe, uint32_t v) { /* pemapper.c:2168-2186 */
  for (;;) {
    if (first == last) return first;
    if (pos[probe] <= v && pos[probe + 1] >= v) return probe;
    if (pos[probe] > v) last = probe - 1; else first = probe + 1;
    probe = (last + first) / 2;
  }
}

int orc_window(const orc_ctx *c, uint32_t spot, int len, uint32_t *start, int *blen) { /* pemapper.c:1052-1060 */
  int ch = find_chrom(c->cstart, 0, c->n_contigs - 1, 7, spot);
  uint32_t extra = 15u * (uint32_t)ch;
  long t = (long)extra + (long)spot - (long)c->p.misalign_slop;
  if (t < 0) t = 0;
  long lo = (long)(uint32_t)(c->cstart[ch] + extra);
  uint32_t s = (uint32_t)(lo > t ? lo : t);
  uint32_t e1 = c->cstart[ch + 1] + extra, e2 = extra + spot + (uint32_t)len + (uint32_t)c->p.misalign_slop;
  uint32_t e = e1 < e2 ? e1 : e2;
  *start = s;
  *blen = (int)(1u + e - s);
  return ch;
}

/* ---------------------------------------------------------------- Smith-Waterman */

static void dp_borders(const orc_ctx *c, double (*S)[DP_DIM][DP_DIM]) { /* init_penalty_matrices 2051-2095 */
  for (int i = 0; i < DP_DIM; i++) {
    S[0][i][0] = 0.0;
    S[1][i][0] = 0.0;
    S[2][i][0] = -1.0 * c->go;
  }
  for (int j = 1; j < DP_DIM; j++) S[0][0][j] = S[1][0][j] = S[2][0][j] = c->border[j];
}

#define MAXD(a, b) (((a) > (b)) ? (a) : (b))

/* smith_waterman_align, pemapper.c:1694-1748 */
static double sw_fill(const orc_ctx *c, double (*S)[DP_DIM][DP_DIM], const char *ref, int nn, const char *seq, int mm,
                      int *start3) {
  const double go = c->go, ge = c->ge;
  for (int i = 1; i <= nn; i++) {
    const double *brow = c->bonus[(unsigned char)ref[i - 1] & 127];
    for (int j = 1; j <= mm; j++) {
      S[2][i][j] = MAXD(S[0][i][j - 1] - go, S[2][i][j - 1] - ge);
      S[1][i][j] = MAXD(S[0][i - 1][j] - go, S[1][i - 1][j] - ge);
      double bump = brow[(unsigned char)seq[j - 1] & 127];
      S[0][i][j] = MAXD(MAXD(S[0][i - 1][j - 1] + bump, S[1][i - 1][j - 1] + bump), S[2][i - 1][j - 1] + bump);
    }
  }
  int bk = 0, bi = 0; /* scan of the last column, 1717-1742 */
  for (int i = 1; i <= nn; i++)
    for (int k = 0; k < 3; k++)
      if (S[k][i][mm] > S[bk][bi][mm]) {
        bk = k;
        bi = i;
      }
  start3[0] = bk;
  start3[1] = bi;
  start3[2] = mm;
  return S[bk][bi][mm];
}

double orc_sw_align(const orc_ctx *c, uint32_t win_start, int blen, const char *seq, int mm, int *start3) {
  double(*S)[DP_DIM][DP_DIM] = malloc(sizeof(double) * 3 * DP_DIM * DP_DIM);
  dp_borders(c, S);
  double r = sw_fill(c, S, c->genome + win_start, blen, seq, mm, start3);
  free(S);
  return r;
}

/* smith_waterman_align (pemapper.c:1694-1748) for windows beyond the reference's 300 x 300 buffers: the same
   recurrence (1710-1713), borders (2062-2081: column 0 = 0 / 0 / -go, row 0 = the c->border values, which the
   reference only defines for j < 300, i.e. reads up to 299 bases) and last-column scan (1717-1742), with the three
   matrices allocated for the window at hand.  BASELINE configs[3] (1000-bp windows) has no reference result to be
   pinned to - the reference overruns its buffers there - so this restatement IS the checker for that shape; for
   windows that fit it returns what orc_sw_align returns (tests/test_oracle_golden.py). */
double orc_sw_align_long(const orc_ctx *c, uint32_t win_start, int blen, const char *seq, int mm, int *start3) {
  const double go = c->go, ge = c->ge;
  const size_t W = (size_t)mm + 1;
  double *S0 = malloc(sizeof(double) * 3 * W * ((size_t)blen + 1));
  double *S1 = S0 + W * ((size_t)blen + 1), *S2 = S1 + W * ((size_t)blen + 1);
  const char *ref = c->genome + win_start;
  for (int i = 0; i <= blen; i++) {
    S0[i * W] = 0.0;
    S1[i * W] = 0.0;
    S2[i * W] = -1.0 * go;
  }
  for (int j = 1; j <= mm; j++) S0[j] = S1[j] = S2[j] = c->border[j];
  for (int i = 1; i <= blen; i++) {
    const double *brow = c->bonus[(unsigned char)ref[i - 1] & 127];
    for (int j = 1; j <= mm; j++) {
      S2[i * W + j] = MAXD(S0[i * W + j - 1] - go, S2[i * W + j - 1] - ge);
      S1[i * W + j] = MAXD(S0[(i - 1) * W + j] - go, S1[(i - 1) * W + j] - ge);
      const double bump = brow[(unsigned char)seq[j - 1] & 127];
      S0[i * W + j] = MAXD(MAXD(S0[(i - 1) * W + j - 1] + bump, S1[(i - 1) * W + j - 1] + bump), S2[(i - 1) * W + j - 1] + bump);
    }
  }
  const double *M[3] = {S0, S1, S2};
  int bk = 0, bi = 0;
  for (int i = 1; i <= blen; i++)
    for (int k = 0; k < 3; k++)
      if (M[k][i * W + mm] > M[bk][bi * W + mm]) {
        bk = k;
        bi = i;
      }
  start3[0] = bk;
  start3[1] = bi;
  start3[2] = mm;
  const double r = M[bk][bi * W + mm];
  free(S0);
  return r;
}

/* ---------------------------------------------------------------- traceback + pileup */

static void add_insertion(orc_ctx *c, uint32_t pos, const char *rev_buf, int n) {
  pthread_mutex_lock(&c->ins_mu);
  if (c->n_ins == c->cap_ins) {
    c->cap_ins = c->cap_ins ? c->cap_ins * 2 : 1024;
    c->ins = realloc(c->ins, c->cap_ins * sizeof(ins_rec));
  }
  if (c->pool_len + n + 1 > c->pool_cap) {
    c->pool_cap = (c->pool_cap ? c->pool_cap * 2 : 65536) + n + 1;
    c->ins_pool = realloc(c->ins_pool, c->pool_cap);
  }
  ins_rec *r = &c->ins[c->n_ins++];
  r->pos = pos;
  r->len = (uint32_t)n;
  r->off = c->pool_len;
  for (int m = 0; m < n; m++) c->ins_pool[c->pool_len + m] = rev_buf[n - (m + 1)]; /* 1892-1893 */
  c->ins_pool[c->pool_len + n] = 0;
  c->pool_len += n + 1;
  pthread_mutex_unlock(&c->ins_mu);
}

/* smith_waterman_backtrack, pemapper.c:1752-1965; win_start = real coordinate of base[0] */
static void sw_backtrack(orc_ctx *c, double (*S)[DP_DIM][DP_DIM], uint32_t win_start, const char *seq, const int *start3) {
  const double go = c->go, ge = c->ge;
  char pend[DP_DIM + 8];
  int n_pend = 0;
  int k = start3[0], i = start3[1], j = start3[2], i1 = 0, j1 = 0;
  while (i > 0 && j > 0) {
    i1 = i - 1;
    j1 = j - 1;
    int pk, pi, pj;
    if (k == 0) { /* 1799-1813 */
      pi = i1; pj = j1; pk = 0;
      double best = S[0][pi][pj];
      if (S[1][pi][pj] > best) { pk = 1; best = S[1][pi][pj]; }
      if (S[2][pi][pj] > best) pk = 2;
    } else if (k == 2) { /* 1814-1822 */
      pi = i; pj = j1; pk = 0;
      if (S[2][pi][pj] - ge > S[0][pi][pj] - go) pk = 2;
    } else { /* 1823-1831 */
      pi = i1; pj = j; pk = 0;
      if (S[1][pi][pj] - ge > S[0][pi][pj] - go) pk = 1;
    }
    uint16_t *ctr = c->cnt[win_start + (uint32_t)i1];
    if (pi != i) {
      if (pj != j) { /* 1846-1858 */
        char r = seq[j1];
        int col = r == 'A' ? 0 : r == 'C' ? 1 : r == 'G' ? 2 : r == 'T' ? 3 : -1;
        if (col >= 0) __atomic_fetch_add(&ctr[col], 1, __ATOMIC_RELAXED);
      } else
        __atomic_fetch_add(&ctr[4], 1, __ATOMIC_RELAXED); /* 1868 */
      if (n_pend > 0) { /* 1871-1904 */
        add_insertion(c, win_start + (uint32_t)i1, pend, n_pend);
        __atomic_fetch_add(&ctr[5], 1, __ATOMIC_RELAXED);
      }
      n_pend = 0;
    } else
      pend[n_pend++] = seq[j1]; /* 1910-1911 */
    i = pi; j = pj; k = pk;
  }
  if (n_pend > 0 && i >= 1) { /* 1918-1958 */
    add_insertion(c, win_start + (uint32_t)i1, pend, n_pend);
    __atomic_fetch_add(&c->cnt[win_start + (uint32_t)i1][5], 1, __ATOMIC_RELAXED);
  }
}

/* ---------------------------------------------------------------- per read / pair */

typedef struct cand_set {
  int n;
  uint32_t spot[256];
  char orient[256];
  uint32_t wstart[256];
  int blen[256];
  double score[256];
  int start[256][3];
} cand_set;

static void score_all(orc_ctx *c, scratch *sc, cand_set *cs, const char *const seq[2], int len, uint64_t *cells) {
  for (int q = 0; q < cs->n; q++) {
    cs->score[q] = sw_fill(c, sc->S, c->genome + cs->wstart[q], cs->blen[q], seq[(int)cs->orient[q]], len, cs->start[q]);
    if (cs->blen[q] > 0) *cells += (uint64_t)cs->blen[q] * (uint64_t)len;
  }
}

/* the "one mate has candidates" rule, pemapper.c:1084-1128 / 1130-1174 */
static int single_rule(const orc_ctx *c, const cand_set *cs, int len, int *best) {
  double good = len * c->p.min_align * c->p.match_bonus;
  double top = -c->go * len;
  int count = 0;
  *best = -1;
  for (int q = 0; q < cs->n; q++) {
    double s = cs->score[q];
    if (s > top && s >= good) {
      top = s;
      count = 1;
      *best = q;
    } else if (fabs(s - top) < 0.0001 && count > 0)
      count++;
  }
  if (count == 0) { *best = -1; return T_NEITHER_MAP; }
  if (count == 1) return T_UNIQUE_SINGLE;
  *best = -1;
  return T_NON_NO;
}

/* find_mate_pairs, pemapper.c:1313-1536 (scores already computed, 1361-1379) */
static int pair_rule(const orc_ctx *c, const cand_set *a, int l1, const cand_set *b, int l3, int *keep1, int *keep2) {
  const int n1 = a->n, n2 = b->n;
  double smax1[257], smax2[257];
  for (int i = 0; i <= 256; i++) smax1[i] = smax2[i] = -1.0; /* 1346-1351 */
  for (int i = 0; i < n1; i++) smax1[i] = a->score[i];
  for (int i = 0; i < n2; i++) smax2[i] = b->score[i];
  double good1 = l1 * c->p.min_align * c->p.match_bonus, good2 = l3 * c->p.min_align * c->p.match_bonus;
  double tot_best = -1e5;
  int perfect = 0, slip = 0, sm1 = -1, sm2 = -1;
  *keep1 = *keep2 = -1;
  for (int w1 = 0; w1 < n1; w1++) {
    if (!(smax1[w1] >= good1)) continue;
    for (int w2 = 0; w2 < n2; w2++) {
      if (!(smax2[w2] >= good2)) continue;
      long dist = labs((long)a->spot[w1] - (long)b->spot[w2]); /* 1388-1394 */
      int ok = dist >= c->p.min_dist && dist <= c->p.max_dist && a->orient[w1] != b->orient[w2];
      if (!ok) continue;
      double t1 = smax1[w1], t2 = smax2[w2];
      double inc = smax1[w1] + smax2[w2] - tot_best;
      if (inc > 0.001) { /* 1401-1410 */
        perfect = 1;
        sm1 = w1;
        sm2 = w2;
        tot_best = t1 + t2;
        slip = 1;
      } else if (inc > -0.001) { /* 1411-1416 */
        if (sm1 == w1 || sm2 == w2) slip++;
        perfect++;
      }
    }
  }
  if (perfect > 0) { /* 1424-1447 */
    if (perfect == 1) { *keep1 = sm1; *keep2 = sm2; return T_UNIQUE_MATE; }
    if (slip == perfect) { *keep1 = sm1; *keep2 = sm2; return T_UNIQUE_SLIP; }
    return T_NON_MATE;
  }
  int best1 = 0, best2 = 0, c1 = 0, c2 = 0; /* 1450-1469, quirks included */
  for (int i = 1; i < n1; i++) {
    if (smax1[i] > smax1[best1]) { best1 = i; c1 = 1; }
    else if (smax1[i] - smax1[best1] > -0.0001) c1++;
  }
  for (int i = 1; i < n2; i++) {
    if (smax2[i] > smax2[best2]) { best2 = i; c2 = 1; }
    else if (smax2[i] - smax2[best1] > -0.0001) c2++; /* sic: best1 (1468) */
  }
  int ok2 = (smax2[best2] >= good2) && (c2 < 2);
  if (smax1[best1] >= good1 && c1 < 2) { /* 1483-1527 */
    *keep1 = best1;
    if (ok2) { *keep2 = best2; return T_UNIQUE_MIS; }
    return T_UNIQUE_SINGLE;
  }
  if (ok2) { *keep2 = best2; return T_UNIQUE_SINGLE; }
  return T_NON_MIS;
}

static void windows_for(const orc_ctx *c, cand_set *cs, int len) {
  for (int q = 0; q < cs->n; q++) orc_window(c, cs->spot[q], len, &cs->wstart[q], &cs->blen[q]);
}

static void map_one(orc_ctx *c, scratch *sc, const char *r1, int l1, const char *r2, int l2, uint32_t *m1, uint32_t *m2,
                    int *type, orc_detail *det, uint64_t *cells) {
  /* body of the read loop of map_everything, pemapper.c:1010-1235 */
  char f1[DP_DIM + 16], v1[DP_DIM + 16], f2[DP_DIM + 16], v2[DP_DIM + 16];
  cand_set *A = (cand_set *)malloc(2 * sizeof(cand_set)), *B = A + 1;
  memcpy(f1, r1, l1);
  f1[l1] = 0;
  reverse_transcribe(f1, v1, l1);
  const char *seq1[2] = {f1, v1}, *seq2[2] = {f2, v2};
  A->n = seed_read(c, sc, f1, v1, l1, A->spot, A->orient);
  B->n = 0;
  if (c->p.pair_flag && r2) {
    memcpy(f2, r2, l2);
    f2[l2] = 0;
    reverse_transcribe(f2, v2, l2);
    B->n = seed_read(c, sc, f2, v2, l2, B->spot, B->orient);
  }
  windows_for(c, A, l1);
  windows_for(c, B, l2);
  int keep1 = -1, keep2 = -1, call;
  if (A->n > 0 && B->n == 0) {
    score_all(c, sc, A, seq1, l1, cells);
    call = single_rule(c, A, l1, &keep1);
  } else if (B->n > 0 && A->n == 0) {
    score_all(c, sc, B, seq2, l2, cells);
    call = single_rule(c, B, l2, &keep2);
  } else if (A->n > 0 && B->n > 0) {
    score_all(c, sc, A, seq1, l1, cells);
    score_all(c, sc, B, seq2, l2, cells);
    call = pair_rule(c, A, l1, B, l2, &keep1, &keep2);
  } else
    call = T_NEITHER_MAP; /* 1186-1192 */
  *m1 = *m2 = 0;
  if (keep1 >= 0) { /* 1195-1212: recompute the winner's matrices, then walk them */
    int st[3];
    sw_fill(c, sc->S, c->genome + A->wstart[keep1], A->blen[keep1], seq1[(int)A->orient[keep1]], l1, st);
    sw_backtrack(c, sc->S, A->wstart[keep1], seq1[(int)A->orient[keep1]], st);
    *m1 = A->wstart[keep1] + (uint32_t)st[1] + 1;
  }
  if (keep2 >= 0) { /* 1216-1231 */
    int st[3];
    sw_fill(c, sc->S, c->genome + B->wstart[keep2], B->blen[keep2], seq2[(int)B->orient[keep2]], l2, st);
    sw_backtrack(c, sc->S, B->wstart[keep2], seq2[(int)B->orient[keep2]], st);
    *m2 = B->wstart[keep2] + (uint32_t)st[1] + 1;
  }
  *type = call;
  if (det) {
    det->hits1 = A->n;
    det->hits2 = B->n;
    det->best1 = keep1;
    det->best2 = keep2;
    det->orient1 = keep1 >= 0 ? A->orient[keep1] : -1;
    det->orient2 = keep2 >= 0 ? B->orient[keep2] : -1;
    det->score1 = keep1 >= 0 ? A->score[keep1] : 0.0;
    det->score2 = keep2 >= 0 ? B->score[keep2] : 0.0;
  }
  free(A);
}

typedef struct job {
  orc_ctx *c;
  int lo, hi, stride;
  const char *reads1, *reads2;
  const int *len1, *len2;
  uint32_t *m1, *m2;
  int *type;
  orc_detail *det;
  uint64_t cells;
} job;

static void *job_run(void *arg) {
  job *jb = arg;
  scratch *sc = scratch_new();
  dp_borders(jb->c, sc->S);
  for (int r = jb->lo; r < jb->hi; r++) {
    const char *r2 = jb->reads2 ? jb->reads2 + (size_t)r * jb->stride : NULL;
    map_one(jb->c, sc, jb->reads1 + (size_t)r * jb->stride, jb->len1[r], r2, jb->len2 ? jb->len2[r] : 0, &jb->m1[r],
            &jb->m2[r], &jb->type[r], jb->det ? &jb->det[r] : NULL, &jb->cells);
  }
  scratch_free(sc);
  return NULL;
}

void orc_map_batch(orc_ctx *c, int n, const char *reads1, const int *len1, const char *reads2, const int *len2,
                   int stride, uint32_t *m1, uint32_t *m2, int *mapping_type, orc_detail *det, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  job *jobs = calloc(nthreads, sizeof(job));
  pthread_t *th = calloc(nthreads, sizeof(pthread_t));
  /* interleaved small blocks would balance better; contiguous ranges keep it simple and deterministic */
  for (int t = 0; t < nthreads; t++) {
    job *jb = &jobs[t];
    jb->c = c;
    jb->lo = (int)((int64_t)n * t / nthreads);
    jb->hi = (int)((int64_t)n * (t + 1) / nthreads);
    jb->stride = stride;
    jb->reads1 = reads1; jb->reads2 = reads2; jb->len1 = len1; jb->len2 = len2;
    jb->m1 = m1; jb->m2 = m2; jb->type = mapping_type; jb->det = det;
    if (nthreads == 1) job_run(jb); else pthread_create(&th[t], NULL, job_run, jb);
  }
  for (int t = 0; t < nthreads; t++) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    c->cells += jobs[t].cells;
  }
  free(jobs);
  free(th);
}

/* ---------------------------------------------------------------- results */

static inline int site_total(const uint16_t *v) { return v[0] + v[1] + v[2] + v[3] + v[4] + v[5]; }

uint64_t orc_count_sites(const orc_ctx *c) {
  uint64_t n = 0;
  for (int64_t p = 0; p < c->genome_size; p++) n += site_total(c->cnt[p]) > 0;
  return n;
}

uint64_t orc_get_records(const orc_ctx *c, orc_record *out, uint64_t cap) { /* pemapper.c:828-842 */
  uint64_t n = 0;
  for (int64_t p = 0; p < c->genome_size; p++)
    if (site_total(c->cnt[p]) > 0) {
      if (n < cap) {
        out[n].pos = (uint32_t)p;
        memcpy(out[n].c, c->cnt[p], 12);
      }
      n++;
    }
  return n;
}

uint64_t orc_n_insertions(const orc_ctx *c) { return c->n_ins; }

uint32_t orc_get_insertion(const orc_ctx *c, uint64_t i, char *buf, int cap) {
  const ins_rec *r = &c->ins[i];
  int n = (int)r->len < cap - 1 ? (int)r->len : cap - 1;
  memcpy(buf, c->ins_pool + r->off, n);
  buf[n] = 0;
  return r->pos;
}

void orc_reset_counts(orc_ctx *c) {
  memset(c->cnt, 0, (size_t)(c->genome_size + 1) * sizeof(*c->cnt));
  c->n_ins = 0;
  c->pool_len = 0;
  c->cells = 0;
}

static const orc_ctx *g_sort_ctx;
static int cmp_ins(const void *a, const void *b) {
  const ins_rec *x = a, *y = b;
  if (x->pos != y->pos) return (x->pos > y->pos) - (x->pos < y->pos);
  return strcmp(g_sort_ctx->ins_pool + x->off, g_sort_ctx->ins_pool + y->off);
}

int orc_write_indel_txt(const orc_ctx *c, const char *path, const char *const *names) { /* pemapper.c:819-862 */
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  ins_rec *srt = malloc((c->n_ins + 1) * sizeof(ins_rec));
  memcpy(srt, c->ins, c->n_ins * sizeof(ins_rec));
  g_sort_ctx = c;
  qsort(srt, c->n_ins, sizeof(ins_rec), cmp_ins);
  uint32_t *padded = calloc(c->n_contigs + 16, 4);
  for (int i = 0; i <= c->n_contigs; i++) padded[i] = c->cstart[i] + 15u * (uint32_t)i; /* 821-822 */
  fprintf(f, "Fragment\tPositions\tReference Base\tTotal Coverage\tReference Reads\tNo Deletions\tNo Insertions\tInsertion Sequence");
  uint64_t q = 0;
  for (int64_t p = 0; p < c->genome_size; p++) {
    const uint16_t *v = c->cnt[p];
    int tot = site_total(v);
    if (tot > 0 && v[5] > 0) {
      char ref = c->genome[p];
      int ref_reads = ref == 'A' ? v[0] : ref == 'C' ? v[1] : ref == 'G' ? v[2] : v[3];
      int which = find_chrom(padded, 0, c->n_contigs - 1, 7, (uint32_t)p);
      fprintf(f, "\n%s\t%d\t%c\t%d\t%d\t%d\t%d", names[which], (int)(1 + (uint32_t)p - padded[which]), ref, tot, ref_reads,
              v[4], v[5]);
      while (q < c->n_ins && srt[q].pos < (uint32_t)p) q++;
      for (; q < c->n_ins && srt[q].pos == (uint32_t)p; q++) fprintf(f, "\t%s", c->ins_pool + srt[q].off);
    }
  }
  fclose(f);
  free(srt);
  free(padded);
  return 0;
}
