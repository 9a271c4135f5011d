/*
 * pemap_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, IEEE double exactly like the reference) of the PEMapper read-mapping
 * hot path of wingolab-org/pecaller, src/pemapper.c.  It exists to CHECK the CUDA path and to be
 * timed as the CPU baseline ("port").  Nothing in the product (pecaller_b200/, include/) may link,
 * load or call it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * Parity status: PINNED.  Every stage is checked against the unmodified reference compiled from
 * /root/reference/src (oracle/_ref, see oracle/Makefile and oracle/ref_wrap.c) and against the
 * committed golden outputs of the reference binaries (tests/golden, tools/make_golden.py).
 * The reference repository itself ships no tests or golden vectors.
 */
#ifndef PEMAP_ORACLE_H
#define PEMAP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  int idepth;         /* k-mer length; index_genome_whole.c:149 fixes it at 16 */
  int max_hits;       /* pemapper.c:162 (200) */
  int too_many_spots; /* pemapper.c:163 (100) */
  double min_align;   /* MIN_ALIGN, pemapper.c:151/244/292 */
  double match_bonus; /* pemapper.c:170 (1.0) */
  int is_bisulfite;   /* IS_BISULFITE, pemapper.c:152 */
  int pair_flag;      /* pemapper.c:143 */
  int min_dist;       /* argv, pemapper.c:295 */
  int max_dist;       /* argv, pemapper.c:294 */
  int misalign_slop;  /* MISALIGN_SLOP, pemapper.c:47 (10) */
} orc_params;

typedef struct orc_ctx orc_ctx;

/* per-read detail beyond what the reference emits (strand / score / candidate counts) */
typedef struct orc_detail {
  int hits1, hits2;     /* candidates from initial_map */
  int best1, best2;     /* winning candidate index or -1 */
  int orient1, orient2; /* strand of the winner (0 fwd, 1 rev) or -1 */
  double score1, score2;/* SW score of the winner (0 if none) */
} orc_detail;

typedef struct orc_record { /* one .pileup.gz record, pemapper.c:834-842 */
  uint32_t pos;
  uint16_t c[6]; /* A C G T Del Ins */
} orc_record;

void orc_default_params(orc_params *p);

/* genome: concatenated upper-case contigs (the inflated .seq), contig_len: REAL lengths.
   Builds the k-mer index exactly as index_genome_whole.c:169-177, 209-299, 334-351 would. */
orc_ctx *orc_create(const char *genome, const int64_t *contig_len, int n_contigs, const orc_params *p);
void orc_destroy(orc_ctx *c);
void orc_set_params(orc_ctx *c, const orc_params *p); /* change mapping params, keep the index */

/* index inspection */
uint64_t orc_n_mers(const orc_ctx *c);
const uint32_t *orc_mers(const orc_ctx *c);
uint32_t orc_pos_index(const orc_ctx *c, uint64_t w); /* value of .idx entry w, w in [0, 2^32] */
int64_t orc_genome_size(const orc_ctx *c);
const uint32_t *orc_contig_starts(const orc_ctx *c); /* n+1 prefix sums of (len-15), as pemapper.c:447-448 */
/* fill a caller-provided dense table of 2^32+1 entries (16 GiB) = the inflated .idx */
void orc_fill_pos_index(const orc_ctx *c, uint32_t *table);
/* write <base>.sdx/.seq/.mdx (and .idx when with_idx) in the reference's formats */
int orc_write_index(const orc_ctx *c, const char *base, const char *const *names, int with_idx);

/* stages */
int orc_initial_map(const orc_ctx *c, const char *read, int len, uint32_t *spots, char *orients);
int orc_window(const orc_ctx *c, uint32_t spot, int len, uint32_t *start, int *blen);
double orc_sw_align(const orc_ctx *c, uint32_t win_start, int blen, const char *seq, int mm, int *start3);
/* the same for windows longer than the reference's 300 x 300 buffers (BASELINE configs[3]) */
double orc_sw_align_long(const orc_ctx *c, uint32_t win_start, int blen, const char *seq, int mm, int *start3);

/* whole batch: reads are rows of a (n x stride) char matrix, NUL-terminated at len[i].
   reads2/len2 NULL for single-end.  Accumulates pileup counters + insertions in the context.
   nthreads<=1: serial.  det may be NULL. */
void orc_map_batch(orc_ctx *c, int n, const char *reads1, const int *len1, const char *reads2, const int *len2,
                   int stride, uint32_t *m1, uint32_t *m2, int *mapping_type, orc_detail *det, int nthreads);

/* results */
uint64_t orc_count_sites(const orc_ctx *c);
uint64_t orc_get_records(const orc_ctx *c, orc_record *out, uint64_t cap);
uint64_t orc_n_insertions(const orc_ctx *c);
/* insertion i: returns site position, copies the string (NUL-terminated) into buf */
uint32_t orc_get_insertion(const orc_ctx *c, uint64_t i, char *buf, int cap);
void orc_reset_counts(orc_ctx *c);
/* text of .indel.txt (without gz) as pemapper.c:819-862 writes it; insertion strings of a site sorted */
int orc_write_indel_txt(const orc_ctx *c, const char *path, const char *const *names);
uint64_t orc_cells(const orc_ctx *c); /* SW cells (nn*mm) evaluated by orc_map_batch so far */

#ifdef __cplusplus
}
#endif
#endif
