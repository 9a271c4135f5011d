/*
 * pemap.h - C-ABI of the B200-native PEMapper read-mapping hot path.
 *
 * The reference (wingolab-org/pecaller, src/pemapper.c) has no library interface: main() hands batches of
 * 20,000 reads to worker threads through pthread_create(map_everything, PTHREAD_DATA_NODE*) (pemapper.c:684,
 * 759) and later reads the workers' side effects out of global arrays (maps1/maps2 775-781, all_base_list
 * 828-864, mate_counts 868-900).  This header is that seam as a C-ABI: plain pointers and sizes, no CUDA or
 * torch types.  A C host keeps the reference's argv handling, FASTQ reader, index loader and gz writers and
 * calls these entry points instead of pthread_create (see INTEGRATION.md for the patch).
 *
 * Everything runs on one B200 per pemap_t; there is NO CPU fallback: every call fails with
 * PEMAP_ERR_CUDA when no sm_100 device is usable.  All functions return 0 on success or a negative
 * PEMAP_ERR_* code and never exit(); pemap_last_error() gives the text the caller can pass to the
 * reference's dump_error() (pemapper.c:2808-2816).
 */
#ifndef PEMAP_H
#define PEMAP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEMAP_OK 0
#define PEMAP_ERR_ARG (-1)         /* bad argument (NULL, sizes, read longer than PEMAP_MAX_READ) */
#define PEMAP_ERR_CUDA (-2)        /* no device / CUDA runtime error */
#define PEMAP_ERR_NOMEM (-3)       /* host or device allocation failed */
#define PEMAP_ERR_UNSUPPORTED (-4) /* outside the reference's defined behaviour (e.g. 2..7 contigs) */

/* Reads longer than PEMAP_MAX_READ overflow the reference's 300x300 DP buffers (window = len+21 <= 299,
   pemapper.c:155, 916, 2073-2081); shorter than PEMAP_MIN_READ read past the read in convert_seq_int (2408-2423).
   Parity with the reference is defined for PEMAP_MIN_READ..PEMAP_MAX_READ only.  The library is lenient at both
   ends: lengths 0..15 are accepted and reported unmapped (m = 0, the single-end / pair rules then see a mate
   without candidates), lengths 279..298 are mapped by the same kernels as an extension WITHOUT a reference
   result to compare with, and only lengths < 0 or > 298 return PEMAP_ERR_ARG. */
#define PEMAP_MAX_READ 278
#define PEMAP_MIN_READ 16

/* mapping_type codes, pemapper.c:37-45 */
enum {
  PEMAP_UNIQUE_MATE = 0, PEMAP_UNIQUE_SLIP = 1, PEMAP_UNIQUE_SINGLE = 2, PEMAP_UNIQUE_MIS = 3, PEMAP_NON_MATE = 4,
  PEMAP_NON_MIS = 5, PEMAP_FRAG_MIS = 6, PEMAP_NON_NO = 7, PEMAP_NEITHER_MAP = 8
};

/* The reference's tunables: compile-time statics and argv values. */
typedef struct pemap_params {
  int idepth;          /* last line of .sdx; always 16 (index_genome_whole.c:149) */
  int max_hits;        /* pemapper.c:162 = 200 (<= 200 supported) */
  int too_many_spots;  /* pemapper.c:163 = 100 */
  double min_align;    /* MIN_ALIGN, argv (pemapper.c:244/292) */
  double match_bonus;  /* pemapper.c:170 = 1.0 */
  int is_bisulfite;    /* IS_BISULFITE, argv (pemapper.c:278/296) */
  int pair_flag;       /* pemapper.c:235/283 */
  int min_dist;        /* argv (pemapper.c:295) */
  int max_dist;        /* argv (pemapper.c:294) */
  int misalign_slop;   /* MISALIGN_SLOP, pemapper.c:47 = 10 */
} pemap_params;

/* The index exactly as pemapper's main() holds it after loading G.sdx/.seq/.idx/.mdx (pemapper.c:411-538). */
typedef struct pemap_index {
  const uint32_t *pos_index;     /* 2^32+1 entries: inflated .idx (init_index_buffer, pemapper.c:2129-2149) */
  const uint32_t *mers;          /* n_mers entries: .mdx (pemapper.c:2151-2154) */
  uint64_t n_mers;               /* == pos_index[2^32] */
  const char *genome;            /* genome_size bytes: inflated .seq, upper-case, contigs concatenated (483-493) */
  uint64_t genome_size;          /* contig_starts[n] + 15*n (pemapper.c:453) */
  const uint32_t *contig_starts; /* no_contigs+1 prefix sums of the .sdx lengths, unpadded (pemapper.c:434-448) */
  int no_contigs;
} pemap_index;

/* one .pileup.gz record, byte for byte what pemapper.c:834-842 gzwrite()s */
typedef struct pemap_record {
  uint32_t pos;  /* 0-based coordinate in the concatenated genome */
  uint16_t c[6]; /* As Cs Gs Ts Dels no_ins */
} pemap_record;

/* one insertion string attached to a site (BASE_NODE.ins[], pemapper.c:1871-1904, 1918-1958) */
typedef struct pemap_insertion {
  uint32_t pos;
  uint32_t len;
  const char *seq; /* NUL-terminated, owned by the library until the next finish/reset/destroy */
} pemap_insertion;

/* optional per-read detail of the last batch (the reference does not emit strand or score) */
typedef struct pemap_detail {
  int32_t hits1, hits2;     /* candidates returned by initial_map for each mate */
  int32_t best1, best2;     /* index of the kept candidate in initial_map order, or -1 */
  int32_t orient1, orient2; /* strand of the kept candidate (0 forward, 1 reverse), or -1 */
  double score1, score2;    /* its smith_waterman_align score (bit pattern of the reference's double) */
} pemap_detail;

typedef struct pemap_stats {
  uint64_t reads;           /* read-mates mapped so far */
  uint64_t lookups;         /* pos_index lookups issued (2 words each) */
  uint64_t mer_positions;   /* positions copied out of mers */
  uint64_t candidates;      /* (read, locus) pairs scored */
  uint64_t sw_cells;        /* sum of nn*mm over scored candidates (pemapper.c:1705-1742 loop bounds) */
  uint64_t tb_cells;        /* cells recomputed in fp64 for the winners' tracebacks */
  uint64_t replayed;        /* read-mates whose integer result had a rational tie and was re-scored in fp64 */
  double ms_seed, ms_sw, ms_select, ms_traceback, ms_total; /* CUDA-event time per stage, accumulated */
  uint64_t launches;        /* kernels launched by this library */
  uint64_t diag_traced;     /* winners whose traceback was a pure diagonal (no gap, no rational tie): no DP recompute */
  uint64_t exact_traced;    /* winners whose integer traceback met a rational tie and was redone in fp64 */
  double ms_tb_diag, ms_tb_int, ms_tb_fp64; /* ms_traceback split: pure-diagonal pileup, integer traceback, fp64 traceback */
  uint64_t tb_cells_int;    /* cells recomputed by the integer traceback kernel */
  uint64_t sw_cells_certified; /* part of sw_cells whose candidates were decided by the ungapped-diagonal certificate, no DP */
} pemap_stats;

typedef struct pemap_ctx pemap_t;

void pemap_default_params(pemap_params *p);

/* Replaces the global set-up of main() (pemapper.c:411-601).  Host arrays are borrowed for the duration of the
   call and copied to `device`.  Called once per GPU. */
int pemap_init(pemap_t **h, const pemap_index *ix, const pemap_params *p, int device);

/* pemap_init without the 16 GiB host table: ix->pos_index may be NULL and `next_idx_bytes(ctx, dst, bytes)` is called
   (2^32+1)*4 / 64 MiB times, in order, to write the next `bytes` bytes of the inflated .idx stream into page-locked
   staging of the library (return non-zero to abort); each chunk is copied to the GPU while the caller inflates the
   next one.  The C host's gzread goes straight into `dst` (init_index_buffer, pemapper.c:2129-2149, chunked). */
typedef int (*pemap_fill_cb)(void *ctx, void *dst, size_t bytes);
int pemap_init_streamed(pemap_t **h, const pemap_index *ix, pemap_fill_cb next_idx_bytes, void *ctx,
                        const pemap_params *p, int device);

/* Same, but builds pos_index/mers on the device from the genome (what index_genome_whole.c:169-177, 248-299,
   334-351 computes), so no 16 GiB host table is needed.  contig_len are REAL lengths (.sdx value + 15). */
int pemap_init_from_genome(pemap_t **h, const char *genome, const int64_t *contig_len, int no_contigs,
                           const pemap_params *p, int device);

/* Change the mapping parameters (min_align, pair_flag, min/max_dist) between batches; the index stays. */
int pemap_set_params(pemap_t *h, const pemap_params *p);

/* Replaces one map_everything() call (pemapper.c:907-1309) for n reads / pairs.  Same fields as
   PTHREAD_DATA_NODE (62-78): NUL-terminated reads and their lengths in, m1/m2/mapping_type out
   (m = 0 for an unmapped mate).  read2/len2 NULL when single-end.  Side effect: pileup counters and insertion
   strings accumulate on the device.  Blocking; inputs are borrowed until return. */
int pemap_map_batch(pemap_t *h, int n, const char *const *read1, const int *len1, const char *const *read2,
                    const int *len2, uint32_t *m1, uint32_t *m2, int *mapping_type);

/* Same with reads as rows of an (n x stride) char matrix (the FASTQ reader can fill it directly).
   Rows need not be NUL-terminated.  Pinned (cudaHostAlloc'ed / registered) buffers are DMA'd in place. */
int pemap_map_batch_rows(pemap_t *h, int n, const char *reads1, const int *len1, const char *reads2,
                         const int *len2, int stride, uint32_t *m1, uint32_t *m2, int *mapping_type);

/* Same with 2-bit packed reads - what a FASTQ decoder thread can emit as it copies a read, and 64 bytes per 150-bp
   read on the host-to-device link instead of 160.  A packed row holds ceil(max_len / 16) code words (base i at bits
   31-2(i%16), 30-2(i%16) of word i/16, first base on top exactly as convert_seq_int builds a k-mer, pemapper.c:2408-2423;
   A 0, C 1, G 2, T 3, N stored as 0) followed by ceil(max_len / 32) N-mask words (bit i%32 of word i/32), padded to a
   multiple of 16 bytes = pemap_packed_stride(max_len); every row of a batch uses the same max_len.  The seed kernel cuts
   its k-mers out of the code words with funnel shifts (reverse strand: bit reversal + complement) and the DP kernels
   read ASCII rows restored on the device.  pemap_pack_read returns PEMAP_ERR_UNSUPPORTED for a read with any character
   other than upper-case A C G T N (IUPAC codes, lower case: the reference scores those by exact character): such a
   batch goes through pemap_map_batch_rows.  Lengths still travel as an int array. */
size_t pemap_packed_stride(int max_len);
int pemap_pack_read(const char *read, int len, int max_len, void *dst);
int pemap_map_batch_packed(pemap_t *h, int n, const void *packed1, const int *len1, const void *packed2,
                           const int *len2, int max_len, uint32_t *m1, uint32_t *m2, int *mapping_type);

/* Same with inputs and outputs already in device memory (device pointers); used to time the kernels alone. */
int pemap_map_batch_device(pemap_t *h, int n, const char *d_reads1, const int *d_len1, const char *d_reads2,
                           const int *d_len2, int stride, int max_len, uint32_t *d_m1, uint32_t *d_m2,
                           int *d_mapping_type);

/* What to retain from each batch for inspection: bit 0 = per-read detail, bit 1 = candidate lists.
   Off by default (costs extra device-to-host copies). */
#define PEMAP_KEEP_DETAIL 1
#define PEMAP_KEEP_CANDIDATES 2
int pemap_keep(pemap_t *h, int flags);

/* Per-read detail of the most recent batch (n entries); needs PEMAP_KEEP_DETAIL. */
int pemap_get_detail(pemap_t *h, pemap_detail *out, int n);

/* Candidate list (initial_map output, pemapper.c:1664-1669) of read-mate `mate` (0/1) of read i of the most
   recent batch: up to cap (spot, orient) pairs; returns the count or a negative error.
   Needs PEMAP_KEEP_CANDIDATES. */
int pemap_get_candidates(pemap_t *h, int i, int mate, uint32_t *spots, int8_t *orients, int cap);

/* Replaces the writer loop of main() (pemapper.c:828-864): every covered site in ascending coordinate as the
   exact 16-byte pileup record, plus the insertion strings.  Buffers are owned by the library and stay valid
   until the next pemap_finish / pemap_reset_counts / pemap_destroy.  Counters are NOT cleared. */
int pemap_finish(pemap_t *h, const pemap_record **records, uint64_t *n_records, const pemap_insertion **ins,
                 uint64_t *n_ins);

/* The same writer loop without materialising the records: the counters are compacted window by window (16 M sites)
   through two bounded device buffers and two pinned host buffers, and every window's records are handed to `cb` in
   ascending coordinate while the next window is compacted and copied.  A fully covered 3.1 Gb genome (50 GB of
   records) therefore needs 0.5 GB of device memory and no host copy of the whole file; the C host gzwrite()s from
   the callback exactly as pemapper.c:834-842 does site by site.  `records` is only valid during the call; a
   non-zero return from `cb` aborts with PEMAP_ERR_ARG.  pemap_finish() is this call collecting into one array. */
typedef int (*pemap_site_cb)(void *ctx, const pemap_record *records, uint64_t n);
int pemap_finish_stream(pemap_t *h, pemap_site_cb cb, void *ctx, uint64_t *n_records);

/* The insertion strings accumulated so far, grouped by site (the second half of pemap_finish; shards that only
   contribute their strings to the writer GPU call this and skip the record pass).  The device append buffer
   (PEMAP_INS_MB, default 256 MB) is drained to host memory whenever a chunk of reads leaves it more than half full,
   so its size bounds one chunk's insertions, not the run's; a chunk that overflows it fails pemap_map_batch* with
   PEMAP_ERR_NOMEM at once. */
int pemap_get_insertions(pemap_t *h, const pemap_insertion **ins, uint64_t *n_ins);

/* Zero the counters and drop the insertions (pemapper_tsw.c dump_output, tsw:849-965, between samples). */
int pemap_reset_counts(pemap_t *h);

/* Multi-GPU plumbing (one pemap_t per GPU, reads sharded by the caller): device pointer and length (in uint32
   words, genome_size*6) of this GPU's counter array so that the caller can sum shards with
   ncclReduce/ncclAllReduce(ncclUint32 or ncclInt32, sum) before pemap_finish on the root.
   Insertions stay per shard: each rank's pemap_finish returns its own. */
int pemap_counts_device(pemap_t *h, void **d_counts, uint64_t *n_words);

/* Single-process multi-GPU hosts (one pemap_t per GPU, e.g. one submitting thread each): add the pileup counters of
   `src` (another GPU of the same box) into `dst` with one kernel on dst's device that reads src's array through
   NVLink peer memory; src's counters are left unchanged.  Replaces what the reference gets for free from its shared
   all_base_list (pemapper.c:1752-1965: every worker thread increments the same array).  Call before
   pemap_finish(dst); insertion strings stay with the handle that produced them. */
int pemap_reduce_counts_peer(pemap_t *dst, pemap_t *src);

/* The reduce SURVEY 8e calls for, as a reduce-scatter over NVLink peer memory instead of a reduce onto one GPU: the
   genome is cut into n_ranks equal slices (at multiples of 2048 sites), and every GPU adds the other GPUs' counters of
   ITS slice to its own with one kernel that reads their arrays through NVLink (16-byte loads).  Afterwards rank r holds
   the final counters of slice r = [*site_first, *site_end) and compacts it itself with pemap_finish_stream_range, so the
   writer receives the slices' records in rank order; no GPU ever holds or moves more than 1/n of the sum.
   Callers must make sure that every rank has finished mapping before any rank starts (a barrier), and that no rank
   resets or maps again before every rank is done (another barrier).
     _ipc   : one process per GPU (torchrun / MPI): exchange pemap_counts_ipc_handle() blobs (64 bytes per rank, in
              rank order) and pass them all; the mappings are cached in the handle.
     _local : one process, one pemap_t per GPU (the C host): hs[which] pulls from the other handles. */
int pemap_counts_ipc_handle(pemap_t *h, void *handle64);
int pemap_reduce_scatter_ipc(pemap_t *h, const void *handles, int n_ranks, int rank, uint64_t *site_first,
                             uint64_t *site_end);
int pemap_reduce_scatter_local(pemap_t *const *hs, int n, int which, uint64_t *site_first, uint64_t *site_end);

/* pemap_finish_stream restricted to the sites [site_first, site_end). */
int pemap_finish_stream_range(pemap_t *h, uint64_t site_first, uint64_t site_end, pemap_site_cb cb, void *ctx,
                              uint64_t *n_records);

/* BASELINE configs[3], the Smith-Waterman kernel alone: read i (forward orientation) is scored against the window
   genome[win_start[i] .. win_start[i] + win_len[i]) with the mapper's own integer kernel (smith_waterman_align,
   pemapper.c:1694-1748, in units of 1/36), for windows of up to PEMAP_SW_MAX_WINDOW rows - the reference's 300 x 300
   buffers stop at len + 21.  All pointers are device pointers; outputs are the score * 36, start[1] (maxi), start[0]
   (maxk) and the tie flags (bit 0: the last-column maximum is not unique in exact arithmetic; may be NULL); *ms is the
   kernel's time from CUDA events.  Does not touch the pileup counters. */
#define PEMAP_SW_MAX_WINDOW 1056
int pemap_sw_score_device(pemap_t *h, int n, const char *d_reads, const int *d_len, int stride, int max_len,
                          const uint32_t *d_win_start, const int *d_win_len, int max_window, int32_t *d_score36,
                          int32_t *d_maxi, int32_t *d_maxk, int32_t *d_flags, float *ms);

/* Page-locked host memory for the batch buffers (read rows, lengths, m1/m2/mapping_type): buffers from here are DMA'd in
   place by pemap_map_batch_rows instead of being staged through the library's own pinned buffers.  The FASTQ reader
   of the C host fills such rows directly.  NULL when the allocation fails. */
void *pemap_host_alloc(size_t bytes);
void pemap_host_free(void *p);

/* The CUDA stream (cudaStream_t as void*) all of this handle's work is issued on, so that a caller can bracket
   calls with its own events. */
int pemap_stream(pemap_t *h, void **stream);

int pemap_get_stats(pemap_t *h, pemap_stats *out);
int pemap_reset_stats(pemap_t *h);

/* Device pointers of the index as it lives in HBM (for tests of the device index builder). */
int pemap_index_device(pemap_t *h, const uint32_t **d_pos_index, const uint32_t **d_mers, uint64_t *n_mers);
/* Copy n words of the device pos_index starting at word `first` to the host. */
int pemap_read_pos_index(pemap_t *h, uint64_t first, uint64_t n, uint32_t *out);
int pemap_read_mers(pemap_t *h, uint64_t first, uint64_t n, uint32_t *out);

const char *pemap_last_error(pemap_t *h);
void pemap_destroy(pemap_t *h);
const char *pemap_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PEMAP_H */
