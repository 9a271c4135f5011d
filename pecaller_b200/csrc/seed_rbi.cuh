// seed_rbi.cuh - seed lookup + co-linear chaining over a device-private ROTATED BUCKET INDEX (sm_100a).
//
// Same contract as seed_chain.cuh (initial_map, pemapper.c:1539-1690: fill_mers 1969-2003, get_mers 2158-2165, the
// gather / veto / sort loop 1594-1640, find_matches 2189-2289, then the windows of map_everything 1047-1081), other
// memory layout.  pos_index / mers are the reference's FILE format; looked up as the reference does they cost
// 49 isolated DRAM accesses per segment (980 per 150-bp read-mate plus ~500 list gathers on a human-sized genome), and
// a B200 serves 41.7 G isolated accesses/s whatever their size (profiles/gather_probe_r01.json) while it streams random
// 1 KiB chunks at 5.3-6.0 TB/s (tools/chunk_probe.cu, profiles/chunk_probe_r02.json).
//
// The 49 k-mers of a segment are the exact code and code ^ (d << 2f), f = 0..15, d = 1..3.  Cut the 32 code bits into
// four byte-wide groups g = 0..3 ("tag" = bits 8g..8g+7, "bucket" = the other 24 bits): the 12 neighbours that differ
// inside group g share the segment's bucket of rotation g.  The index keeps, for every rotation, all indexed positions
// grouped by bucket (a directory of 2^24+1 offsets) with their one-byte tags, so the 49 lists of a segment are FOUR
// contiguous reads of ~256 * (genome / 2^32) entries each (~0.9 KB on 3.1 Gb) filtered by tag, instead of 49 + ~25
// isolated accesses.  A k-mer with >= too_many_spots positions is stored as ONE marker entry (its positions are never
// used: 1602-1606 empties the segment).  Nothing is approximated: every position of every one of the 49 lists comes out.
//
// Bucket layout: 80-byte blocks of 16 entries - four quads of positions (64 bytes) followed by their 16 tags - so that
// a bucket is ONE contiguous run that the lanes read front to back (isolated pieces are what HBM serves slowly:
// chunk_probe_r02.json; with tags behind all positions the same loads ran at 3.2 TB/s, profiles/README_r02.md).  The
// directory counts blocks.  Unused slots of the last block hold PM_RBI_EMPTY.
//
// Chaining: the reference sorts every segment list and, per anchor, scans the later lists for a position whose
// diagonal (pos - segment offset) is within +-11 of the anchor's.  Here the entries of a strand are hashed by
// diagonal / 16 in shared memory, every entry gets its `found` count in one parallel pass over the three bins around
// it, and only the few anchors that reach min_match are ordered (segment, position) and fed to the reference's
// sequential rules (reset on better, dedup, the 200-cap return).
#pragma once
#include "pemap_common.cuh"

#define PM_RBI_EMPTY 0xFFFFFFFFu
#define PM_RBI_MARK 0xFFFFFFFEu
#define PM_RBI_CAP 512                  // entries of one strand kept in shared memory (a 150-bp read on 3.1 Gb has ~360)
#ifndef PM_RBI_PREFETCH
#define PM_RBI_PREFETCH 1               // segments whose buckets are prefetched into L2 ahead of the one being read
#endif
#ifndef PM_RBI_PREFETCH_MODE
#define PM_RBI_PREFETCH_MODE 0
#endif
#ifndef PM_RBI_PREFETCH_STRIDE
#define PM_RBI_PREFETCH_STRIDE 128
#endif
#ifndef PM_RBI_EXPERIMENT
#define PM_RBI_EXPERIMENT 0
#endif
#define PM_RBI_CAP2 2048                // second pass: strands of reads that sit in repeats
#define PM_RBI_MAXB (4 * PM_MAX_SEG)    // buckets per strand
#define PM_RBI_BIG_CAP 98304            // >= 19 segments * 49 * 99 positions: the slow path holds any strand
#define PM_RBI_BIG_TAB 131072
#ifndef PM_RBI_UNROLL
#define PM_RBI_UNROLL 4
#endif

namespace pm {

__host__ __device__ __forceinline__ uint32_t rbi_tag(uint32_t code, int g) { return (code >> (8 * g)) & 255u; }
__host__ __device__ __forceinline__ uint32_t rbi_bucket(uint32_t code, int g) {
  const uint32_t lo = g ? (code & ((1u << (8 * g)) - 1u)) : 0u;
  const uint32_t hi = g == 3 ? 0u : (code >> (8 * g + 8));
  return (hi << (8 * g)) | lo;
}
#define PM_RBI_BLOCK_BYTES 80            // 16 positions + 16 tags
__host__ __device__ __forceinline__ uint32_t rbi_units(uint32_t n) {  // blocks of a bucket with n entries
  return (n + 15u) >> 4;
}

struct RbiIndex {
  const uint4* data[4];     // bucket arrays of the four rotations
  const uint32_t* dir[4];   // 2^24+1 offsets each, in 80-byte blocks
};

// ------------------------------------------------------------------------------------------------ builder kernels

// code_of[i] = k-mer code owning mers[i]: one thread per code, runs of pos_index (only the pemap_init path needs it)
__global__ void __launch_bounds__(256) k_rbi_expand_codes(const uint32_t* pos_index, uint32_t* code_of, uint64_t n_mers) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w > 0xFFFFFFFFull) return;
  uint64_t lo = pos_index[w], hi = pos_index[w + 1];
  if (hi > n_mers) hi = n_mers;
  for (uint64_t i = lo; i < hi; i++) code_of[i] = (uint32_t)w;
}

// keep flag of entry i of the code-sorted (code, pos) list: positions of k-mers below the veto threshold are kept, a
// crowded k-mer keeps its first entry only (turned into the marker by k_rbi_mark after the compaction).  Counts are
// get_mers' (2158-2165): pos_index[(u32)(w+1)] - pos_index[w] in 32-bit arithmetic, so the count of 0xFFFFFFFF wraps
// (`last_cnt`, computed on the host).
__device__ __forceinline__ uint32_t rbi_kmer_count(const uint32_t* pos_index, uint32_t c, uint32_t last_cnt) {
  return c == 0xFFFFFFFFu ? last_cnt : pos_index[(uint64_t)c + 1] - pos_index[c];
}
__global__ void __launch_bounds__(256) k_rbi_flag(const uint32_t* code, const uint32_t* pos_index, uint64_t n, uint32_t too_many,
                                                  uint32_t last_cnt, unsigned char* flag) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = code[i];
  const uint32_t cnt = rbi_kmer_count(pos_index, c, last_cnt);
  const uint64_t rank = i - pos_index[c];
  // cnt < true count only for the wrapped 0xFFFFFFFF (then the reference sees the first cnt positions)
  flag[i] = cnt >= too_many ? (rank == 0 ? 1 : 0) : (rank < cnt ? 1 : 0);
}
__global__ void __launch_bounds__(256) k_rbi_mark(const uint32_t* code, const uint32_t* pos_index, uint64_t n, uint32_t too_many,
                                                  uint32_t last_cnt, uint32_t* val) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && rbi_kmer_count(pos_index, code[i], last_cnt) >= too_many) val[i] = PM_RBI_MARK;
}

__global__ void __launch_bounds__(256) k_rbi_keys(const uint32_t* code, uint64_t n, int g, uint32_t* key) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (rbi_bucket(code[i], g) << 8) | rbi_tag(code[i], g);
}

// bstart[b] = first entry of bucket b in the key-sorted list, b = 0..2^24
__global__ void __launch_bounds__(256) k_rbi_bucket_starts(const uint32_t* key, uint64_t n, uint32_t* bstart) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > (1u << 24)) return;
  uint64_t lo = 0, hi = n;
  if (b == (1u << 24)) lo = n;
  else {
    const uint32_t k = b << 8;
    while (lo < hi) {
      const uint64_t m = (lo + hi) >> 1;
      if (key[m] < k) lo = m + 1; else hi = m;
    }
  }
  bstart[b] = (uint32_t)lo;
}

__global__ void __launch_bounds__(256) k_rbi_bucket_units(const uint32_t* bstart, uint32_t* units) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > (1u << 24)) return;
  units[b] = b == (1u << 24) ? 0u : rbi_units(bstart[b + 1] - bstart[b]);
}

__global__ void __launch_bounds__(256) k_rbi_fill(const uint32_t* key, const uint32_t* val, uint64_t n, const uint32_t* bstart,
                                                  const uint32_t* dir, uint4* data) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = key[i], b = k >> 8;
  const uint32_t first = bstart[b], nb = bstart[b + 1] - first, r = (uint32_t)(i - first);
  (void)nb;
  unsigned char* blk = reinterpret_cast<unsigned char*>(data) + ((uint64_t)dir[b] + (r >> 4)) * PM_RBI_BLOCK_BYTES;
  reinterpret_cast<uint32_t*>(blk)[r & 15u] = val[i];
  blk[64 + (r & 15u)] = (unsigned char)(k & 255u);
}

// ------------------------------------------------------------------------------------------------ packed reads
// Row of a packed read (pemap.h: pemap_pack_read): code words (base i at bits 31-2(i%16), 30-2(i%16) of word i/16; A 0,
// C 1, G 2, T 3 as convert_seq_int codes them, N stored as 0), then N-mask words (bit i%32 of word i/32).
// k_unpack_reads restores the ASCII rows the DP kernels read; the seed kernel works on the packed words directly.
__global__ void __launch_bounds__(256) k_unpack_reads(const unsigned char* packed, int pstride, int code_words, const int* len, int n,
                                                      char* rows, int stride) {
  const int words = stride / 16;  // 16 bases per thread
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * words) return;
  const int r = (int)(t / words), w = (int)(t % words);
  uint4 out = make_uint4(0, 0, 0, 0);
  if (16 * w < len[r] && w < code_words) {
    const uint32_t* prow = reinterpret_cast<const uint32_t*>(packed + (size_t)r * pstride);
    const uint32_t c = prow[w];
    const uint32_t m = (prow[code_words + (w >> 1)] >> (16 * (w & 1))) & 0xFFFFu;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int i = 4 * k + j;
        const uint32_t code = (c >> (30 - 2 * i)) & 3u;
        const uint32_t ch = ((m >> i) & 1u) ? 'N' : (code == 0 ? 'A' : code == 1 ? 'C' : code == 2 ? 'G' : 'T');
        x |= ch << (8 * j);
      }
      o[k] = x;
    }
    out = make_uint4(o[0], o[1], o[2], o[3]);
  }
  *reinterpret_cast<uint4*>(rows + (size_t)r * stride + 16 * w) = out;
}

// ------------------------------------------------------------------------------------------------ seed kernel

struct SeedRbiArgs {
  RbiIndex ix;
  const uint32_t* cstart;      // n_contigs+1 (padded to >= 9 entries)
  const char* reads[2];        // [n][stride] each (ASCII)
  const int* len[2];
  int stride;
  const unsigned char* packed[2];  // the same reads 2-bit packed (pemap.h: pemap_pack_read), or nullptr: then `reads` is decoded
  int pstride, pcode_words, pmask_words;
  int n_reads;                 // reads (pairs) in this chunk
  int paired;
  Task* tasks;
  uint32_t* task_cursor;
  uint32_t task_cap;
  uint32_t* cand_base;         // [2*n_reads]
  uint32_t* cand_n;            // [2*n_reads]
  SeedCounters* counters;
  uint32_t* next_list;         // work items whose strand lists did not fit this pass's stores (nullptr in the last pass)
  uint32_t* next_cursor;
  const uint32_t* work_list;   // the items to process, *work_n of them; nullptr = every read-mate of the chunk
  const uint32_t* work_n;
  unsigned char* big_scratch;  // BIG: per warp PM_RBI_BIG_BYTES
  int fast_cap;                // entries of a strand the first pass accepts (<= PM_RBI_CAP; PEMAP_RBI_CAP lowers it in tests)
  DevParams p;
};

#define PM_RBI_BIG_BYTES ((size_t)PM_RBI_BIG_CAP * 10 + (size_t)PM_RBI_BIG_TAB * 4)

struct RbiWarpSmem {           // per warp, every path
  uint32_t b_off[2 * PM_RBI_MAXB];   // bucket start (16-byte units), both strands: [4 * (strand * nseg + segment) + rotation]
  uint16_t b_n4[2 * PM_RBI_MAXB];    // its position quads (<= 256 * 99 / 4)
  uint32_t kcode[2 * PM_MAX_SEG];
  uint32_t hit_pos[PM_MAX_HITS];
  uint16_t hit_off[PM_MAX_HITS];
  uint8_t hit_or[PM_MAX_HITS];
  union {
    uint32_t rd_words[36];           // packed input: code words [0, 19], N-mask words after them, one guard word at [32]
    char rd[2][PM_DP_MAX];           // forward read and its reverse_transcribe (C->T converted when bisulfite): until the k-mers are cut
    struct {
      uint32_t n_ent;                // entries of the strand gathered so far (appended to by the lanes that hold a match)
      uint32_t pend[64];             // anchors waiting for their exact found count
    } g;
  };
};

template <int CAP>
struct RbiSmemStore {          // per warp: the entries of one strand (CAP = 512 first pass, 2048 second pass)
  uint32_t pos[CAP];
  uint32_t head[CAP];                // chain heads + segment sets; afterwards scratch of the anchor sort
  uint16_t next[CAP];
  uint8_t seg[CAP];
  uint8_t found[CAP];
};

__device__ __forceinline__ uint4 rbi_ld16(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t rbi_ld4(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t rbi_hash(uint32_t bin, uint32_t mask) {
  uint32_t x = bin * 0x9E3779B1u;
  return (x ^ (x >> 15)) & mask;
}

// ascending sort of idx[0..n) by key pos[idx] (keys are distinct: positions of one segment); idx has room for the next
// power of two, the padding sorts last
__device__ __forceinline__ void rbi_sort_idx(uint32_t* idx, const uint32_t* pos, int n, int lane) {
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + lane; i < P; i += 32) idx[i] = 0xFFFFFFFFu;
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int x = i ^ j;
        if (x > i) {
          const uint32_t a = idx[i], b = idx[x];
          const uint32_t ka = a == 0xFFFFFFFFu ? 0xFFFFFFFFu : pos[a], kb = b == 0xFFFFFFFFu ? 0xFFFFFFFFu : pos[b];
          const bool up = (i & k) == 0;
          if ((ka > kb) == up) {
            idx[i] = b;
            idx[x] = a;
          }
        }
      }
      __syncwarp();
    }
}

// One read-mate.  BIG = false: entries in shared memory (st_*), a strand with more than `cap` entries makes the function
// return false and the caller queues the read-mate for the BIG pass, whose stores live in a per-warp global scratch.
template <int CAPT, class NextT>   // CAPT = capacity of the shared-memory stores, 0 = stores in global memory (BIG)
__device__ __forceinline__ bool rbi_map_read_mate(const SeedRbiArgs& a, RbiWarpSmem& sm, uint32_t* st_pos, uint32_t* st_head,
                                                  NextT* st_next, uint8_t* st_seg, uint8_t* st_found, const int cap,
                                                  const uint32_t tab_mask, const int w, const int lane,
                                                  unsigned long long& stat_pos, unsigned long long& stat_cand,
                                                  unsigned long long& stat_cells, unsigned long long& stat_lookups) {
  constexpr NextT NIL = (NextT)~(NextT)0;
  constexpr bool BIG = CAPT == 0;
  constexpr uint32_t IMASK = BIG ? 0xFFFFFFFFu : (uint32_t)(2 * CAPT - 1);  // index field of a slot word; all ones = none
  static_assert(CAPT <= 4096, "the segment set of a slot word starts at bit 13");
  const int r = a.paired ? (w >> 1) : w, mate = a.paired ? (w & 1) : 0;
  const uint32_t rm = 2u * (uint32_t)r + (uint32_t)mate;
  const int len = a.len[mate][r];
  const char* read = a.reads[mate] + (size_t)r * a.stride;
  int tot = 0;
  unsigned long long l_pos = 0, l_lookups = 0;  // booked when the read-mate is finished (not when it is handed to the BIG pass)
  bool ok = (len >= 16 && len < PM_DP_MAX - 21);
  const bool from_packed = a.packed[0] != nullptr;
  uint32_t* const pw = reinterpret_cast<uint32_t*>(&sm.rd[0][0]);  // packed input: code words, then N-mask words
  int n_masked = 0;
  if (ok && from_packed) {
    // 2-bit packed read: one coalesced 4-byte load per lane brings codes and N mask (<= 30 words); the N filter
    // (1552-1559) is a population count, the k-mers are funnel shifts of the code words
    const uint32_t* prow = reinterpret_cast<const uint32_t*>(a.packed[mate] + (size_t)r * a.pstride);
    const int nw = a.pcode_words + a.pmask_words;
    uint32_t v = lane < nw ? __ldg(prow + lane) : 0u;
    n_masked = __reduce_add_sync(0xFFFFFFFFu, (lane >= a.pcode_words && lane < nw) ? __popc(v) : 0);
    pw[lane] = v;
    if (lane == 0) pw[32] = 0u;
    if (n_masked >= 1 + len / 10) ok = false;
  } else if (ok) {
    // N filter (1552-1559) + forward / reverse-transcribed copies (1019-1021, 1561-1570)
    int n_count = 0;
    for (int i = lane; i < len; i += 32) {
      const char ch = read[i];
      n_count += (ch == 'N');
      char f = ch, v = rt_char(ch);
      if (a.p.is_bisulfite) {
        if (f == 'C') f = 'T';
        if (v == 'C') v = 'T';
      }
      sm.rd[0][i] = f;
      sm.rd[1][len - 1 - i] = v;
    }
    n_count = __reduce_add_sync(0xFFFFFFFFu, n_count);
    if (n_count >= 1 + len / 10) ok = false;
  }
  __syncwarp();

  if (ok) {
    int total_cuts = len / 16;  // 1573-1587 with idepth == 16
    if ((len & 15) == 0) total_cuts--;
    const int nseg = total_cuts + 1;
    // exact 16-mer of every (strand, segment): convert_seq_int 2408-2423
    for (int ss = lane; ss < 2 * nseg; ss += 32) {
      const int strand = ss >= nseg, s = strand ? ss - nseg : ss;
      const int off = (s < total_cuts) ? 16 * s : len - 16;
      uint32_t code = 0;
      if (from_packed) {
        // base i of the read sits at bits 31-2(i%16), 30-2(i%16) of code word i/16: 16 bases from any offset are one funnel
        // shift, first base on top as convert_seq_int has it; N is packed as A (cv[], 2379-2383)
        const int fo = strand ? len - off - 16 : off;  // the reverse strand's segment covers these forward bases
        code = __funnelshift_l(pw[(fo >> 4) + 1], pw[fo >> 4], 2 * (fo & 15));
        if (strand) {  // reverse_transcribe (2303-2337): field order reversed, every base complemented (3 - c = ~c), N stays N = 0
          uint32_t y = __brev(code);
          y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
          code = ~y;
          if (n_masked) {
            const uint32_t* mw = pw + a.pcode_words;
            uint32_t m = __funnelshift_r(mw[fo >> 5], (fo >> 5) + 1 < a.pmask_words ? mw[(fo >> 5) + 1] : 0u, fo & 31) & 0xFFFFu;
            m = (m | (m << 8)) & 0x00FF00FFu;  // forward base fo + i is field i from the bottom of the reversed word
            m = (m | (m << 4)) & 0x0F0F0F0Fu;
            m = (m | (m << 2)) & 0x33333333u;
            m = (m | (m << 1)) & 0x55555555u;
            code &= ~(m | (m << 1));
          }
        }
        // convert_ct (2292-2300) rewrites C as T in the forward read AND in its reverse transcript: C (01) -> T (11)
        if (a.p.is_bisulfite) code |= (code & 0x55555555u & ~(code >> 1)) << 1;
      } else {
        const char* q = sm.rd[strand] + off;
#pragma unroll
        for (int i = 0; i < 16; i++) code = (code << 2) | base_code(q[i]);
      }
      sm.kcode[ss] = code;
    }
    __syncwarp();  // every lane is done with the read's staging (it shares its shared memory with the gather's scratch)
    l_lookups = (unsigned long long)(2 * nseg * PM_KV);

    int min_match = total_cuts > 1 ? total_cuts : 1;  // 1642-1645
    if (total_cuts > 4) min_match = (4 * total_cuts) / 5;
    if (min_match > 4) min_match = 4;
    const int max_depth = total_cuts;
    const int mo = (a.p.idepth - 4 > 2) ? a.p.idepth - 4 : 2;  // max_off (2196)
    const int nb = 4 * nseg;
    // ---- bucket directory: four buckets per (strand, segment), all 8 * nseg look-ups in one go
    for (int b = lane; b < 2 * nb; b += 32) {
      const int g = b & 3;
      const uint32_t code = sm.kcode[b >> 2];
      const uint32_t bk = rbi_bucket(code, g);
      const uint32_t d0 = __ldg(a.ix.dir[g] + bk), d1 = __ldg(a.ix.dir[g] + bk + 1);
      sm.b_off[b] = d0;
      sm.b_n4[b] = (uint16_t)(4u * (d1 - d0));  // quads of positions: four per block
    }
    __syncwarp();
    // lanes 8g..8g+7 read the bucket of rotation g
    const int rot = lane >> 3, l8 = lane & 7;
    const uint4* const rdata = a.ix.data[rot];
    const uint32_t keep_exact = rot == 0 ? 0x80808080u : 0u;  // the exact k-mer sits in all four buckets: rotation 0 takes it

    for (int strand = 0; strand < 2; strand++) {
      if (strand == 1 && tot >= a.p.max_hits) break;  // 1658
      // ---- hash of the entries by diagonal / 16.  Shared-memory paths: a slot word holds the chain head (low bits, all
      // ones = none) and the set of segments hashed into the slot (bit 13 + segment); BIG: the head index alone.
      constexpr uint32_t HNIL = IMASK;
      auto build_hash = [&](const int n_entries) {
        for (uint32_t i = 4u * lane; i <= tab_mask; i += 128)
          *reinterpret_cast<uint4*>(st_head + i) = make_uint4(HNIL, HNIL, HNIL, HNIL);
        __syncwarp();
        for (int e = lane; e < n_entries; e += 32) {
          const int s = st_seg[e];
          const uint32_t off = (s < total_cuts) ? 16u * (uint32_t)s : (uint32_t)(len - 16);
          const uint32_t bin = (uint32_t)(((unsigned long long)st_pos[e] + 512ull - off) >> 4);
          uint32_t* hp = &st_head[rbi_hash(bin, tab_mask)];
          uint32_t old;
          if (BIG) {
            old = atomicExch(hp, (uint32_t)e);
          } else {
            old = *hp;
            uint32_t assumed;
            do {
              assumed = old;
              old = atomicCAS(hp, assumed, (assumed & 0xFFFFE000u) | (1u << (13 + s)) | (uint32_t)e);
            } while (old != assumed);
            old &= IMASK;
          }
          st_next[e] = old == HNIL ? NIL : (NextT)old;
        }
        __syncwarp();
      };
      // Is there a pair of entries of two different segments whose diagonals lie within 2 * (max_off - 1) of each other?
      // An anchor with `found` = F has F of the nseg segments agreeing within max_off - 1 of its diagonal, so at most
      // nseg - F segments are outside the chain and ANY nseg - F + 2 segments hold two of its members: without such a
      // pair among the first nseg - F + 2 segments no anchor of the strand reaches F.
      auto close_pair_exists = [&](const int n_entries) {
        bool yes = false;
        for (int e = lane; e < n_entries; e += 32) {
          const int s = st_seg[e];
          const uint32_t off = (s < total_cuts) ? 16u * (uint32_t)s : (uint32_t)(len - 16);
          const long long dg = (long long)st_pos[e] + 512ll - (long long)off;
          const uint32_t bin = (uint32_t)(dg >> 4);
#pragma unroll
          for (int db = -2; db <= 2; db++) {
            uint32_t q = st_head[rbi_hash(bin + (uint32_t)db, tab_mask)] & IMASK;
            while (q != HNIL) {
              const int sq = st_seg[q];
              if (sq != s) {
                const uint32_t offq = (sq < total_cuts) ? 16u * (uint32_t)sq : (uint32_t)(len - 16);
                const long long d = (long long)st_pos[q] + 512ll - (long long)offq - dg;
                if (d > -2ll * mo + 1 && d < 2ll * mo - 1) yes = true;
              }
              const NextT nx = st_next[q];
              q = nx == NIL ? HNIL : (uint32_t)nx;
            }
          }
        }
        return __any_sync(0xFFFFFFFFu, yes) != 0;
      };
      // ---- gather (get_mers 2158-2165, loop 1594-1612 / 1619-1637): per segment, every entry of its four buckets whose
      // tag is the segment's tag or one 2-bit field away from it
      // With the running min_match at F the first nseg - F + 2 segments decide whether the strand can matter at all (see
      // close_pair_exists): after a full-length hit on the forward strand the reverse strand is settled by two segments.
      // (tried for every strand, i.e. also 8 of 10 segments at the initial min_match of 4: 269 against 201 ms per 8 M
      // read-mates - the probe's hash build and the broken prefetch run cost more than two segments of reads)
      const int k_probe = nseg - min_match + 2 <= nseg / 2 ? nseg - min_match + 2 : 0;
      bool strand_dead = false;
      int cnt = 0;              // warp-uniform copy of sm.g.n_ent between segments
      uint32_t min_spots = 10000;
      bool overflow = false;
      if (lane == 0) sm.g.n_ent = 0;
      __syncwarp();
      // Bytes in flight, not bandwidth, bound the bucket reads (24 warps x 32 lanes x a few 16-byte registers keep
      // HBM's queues too short for it to schedule well: 3.2 TB/s; the same reads issued 250 KB deep per SM reach
      // 5.5-6 TB/s, tools/chunk_probe.cu).  So the buckets of the segments ahead are pulled into L2 by prefetches, one
      // 128-byte line per lane and instruction, which cost neither registers nor shared memory.
      auto prefetch_segment = [&](const int sp) {
        if (sp >= nseg) return;
        const int bp = 4 * (strand * nseg + sp) + rot;
        const char* p0 = reinterpret_cast<const char*>(rdata + 5ull * sm.b_off[bp]);
        const uint32_t bytes = 20u * sm.b_n4[bp];  // 80 bytes per block = 20 per quad
#if PM_RBI_PREFETCH_MODE == 1   /* one bulk prefetch per bucket (TMA unit): the whole byte range at once */
        if (l8 == 0 && bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"(bytes) : "memory");
#else                           /* one line-sized piece per lane and instruction */
        const char* line = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)(PM_RBI_PREFETCH_STRIDE - 1)) +
                           PM_RBI_PREFETCH_STRIDE * l8;
        for (; line < p0 + bytes; line += 8 * PM_RBI_PREFETCH_STRIDE) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
#endif
      };
#if PM_RBI_PREFETCH > 0
      for (int sp = 0; sp < PM_RBI_PREFETCH; sp++)
        if (k_probe == 0 || sp < k_probe) prefetch_segment(sp);
#endif
      for (int s = 0; s < nseg; s++) {
#if PM_RBI_PREFETCH > 0
        if (k_probe == 0 || s + PM_RBI_PREFETCH < k_probe || s >= k_probe) prefetch_segment(s + PM_RBI_PREFETCH);
#endif
        const int b = 4 * (strand * nseg + s) + rot;
        const uint32_t n4 = sm.b_n4[b];
        const uint4* base = rdata + 5ull * sm.b_off[b];  // 80-byte blocks
        const uint32_t etagx = ((sm.kcode[strand * nseg + s] >> (8 * rot)) & 255u) * 0x01010101u;
        const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, n4);
        const int cnt0 = cnt;
        for (uint32_t q0 = 0; q0 < nmax; q0 += 8 * PM_RBI_UNROLL) {
          uint4 P[PM_RBI_UNROLL];
          uint32_t T[PM_RBI_UNROLL];
#pragma unroll
          for (int u = 0; u < PM_RBI_UNROLL; u++) {
            const uint32_t q = q0 + (uint32_t)(8 * u + l8);
            T[u] = etagx ^ 0x0F0F0F0Fu;  // two fields away: never qualifies (P[u] is then never looked at)
            if (q < n4) {
#if PM_RBI_EXPERIMENT == 2   /* timing experiment: no bucket loads, pseudo-random tags and positions */
              uint32_t x = (q + sm.b_off[b]) * 0x9E3779B1u;
              x ^= x >> 15;
              x *= 0x85EBCA77u;
              T[u] = x ^ (x >> 13);
              P[u] = make_uint4(x & 0x7FFFFFFFu, (x * 3u) & 0x7FFFFFFFu, (x * 5u) & 0x7FFFFFFFu, (x * 7u) & 0x7FFFFFFFu);
#else
              const uint4* blk = base + 5u * (q >> 2);  // quad q: positions at 16 * (q & 3), its four tags at 64 + 4 * (q & 3)
              P[u] = rbi_ld16(blk + (q & 3u));
              T[u] = rbi_ld4(reinterpret_cast<const uint32_t*>(blk + 4) + (q & 3u));
#endif
            }
          }
#if PM_RBI_EXPERIMENT == 1   /* timing experiment: the bucket loads alone */
          {
            uint32_t acc = 0;
#pragma unroll
            for (int u = 0; u < PM_RBI_UNROLL; u++)
              if (q0 + (uint32_t)(8 * u + l8) < n4) acc ^= P[u].x ^ P[u].y ^ P[u].z ^ P[u].w ^ T[u];
            if (acc == 0x12345678u) overflow = true;
            continue;
          }
#endif
#pragma unroll
          for (int u = 0; u < PM_RBI_UNROLL; u++) {
            // per tag byte: fields that differ from the exact tag; a byte qualifies when at most one field differs
            const uint32_t X = T[u] ^ etagx;
            const uint32_t D = (X | (X >> 1)) & 0x55555555u;
            const uint32_t Z = D & ((D | 0x80808080u) - 0x01010101u);             // byte == 0 <=> <= 1 field differs
            uint32_t hit = ((Z + 0x7F7F7F7Fu) & 0x80808080u) ^ 0x80808080u;       // bit 7 of byte k: tag k qualifies
            hit &= ((D + 0x7F7F7F7Fu) & 0x80808080u) | keep_exact;                 // ... and it is not the exact tag (rot > 0)
            if (hit) {  // ~19 % of the quads: the lane reserves its slots in the strand's list and fills them, no loop
              const uint32_t at = atomicAdd(&sm.g.n_ent, (uint32_t)__popc(hit));
              if (at + 4u <= (uint32_t)cap) {
                if (hit & 0x80u) st_pos[at] = P[u].x;
                if (hit & 0x8000u) st_pos[at + ((hit >> 7) & 1u)] = P[u].y;
                if (hit & 0x800000u) st_pos[at + (uint32_t)__popc(hit & 0x8080u)] = P[u].z;
                if (hit & 0x80000000u) st_pos[at + (uint32_t)__popc(hit & 0x808080u)] = P[u].w;
              } else {
                overflow = true;
              }
            }
          }
        }
        __syncwarp();
        cnt = (int)sm.g.n_ent;
        if (cnt > cap) cnt = cap;  // (overflow is reported below; keep the reads in range)
        // markers of crowded k-mers and the padding of a bucket's last quad came along as positions: a marker empties
        // the segment's list (1602-1606), padding is dropped (only tags next to 0xFF ever pick it up)
        bool crowded = false, padded = false;
        for (int e = cnt0 + lane; e < cnt; e += 32) {
          const uint32_t v = st_pos[e];
          crowded |= v == PM_RBI_MARK;
          padded |= v == PM_RBI_EMPTY;
          st_seg[e] = (uint8_t)s;
        }
        if (__any_sync(0xFFFFFFFFu, crowded)) cnt = cnt0;
        else if (__any_sync(0xFFFFFFFFu, padded)) {  // stable in-place compaction, 32 entries at a time
          int wr = cnt0;
          for (int e0 = cnt0; e0 < cnt; e0 += 32) {
            const int e = e0 + lane;
            const uint32_t v = e < cnt ? st_pos[e] : PM_RBI_EMPTY;
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, v != PM_RBI_EMPTY);
            __syncwarp();
            if (v != PM_RBI_EMPTY) st_pos[wr + __popc(bal & ((1u << lane) - 1u))] = v;
            wr += __popc(bal);
            __syncwarp();
          }
          cnt = wr;
        }
        __syncwarp();
        if (lane == 0) sm.g.n_ent = (uint32_t)cnt;
        __syncwarp();
        min_spots = min(min_spots, (uint32_t)(cnt - cnt0));
        if (s + 1 == k_probe && min_spots <= (uint32_t)a.p.max_hits && !__any_sync(0xFFFFFFFFu, overflow)) {
          // (min_spots <= max_hits: the rule of 2200-2207 cannot fire whatever the other segments hold)
          build_hash(cnt);
          if (!close_pair_exists(cnt)) {
            strand_dead = true;
            break;
          }
        }
      }
      if (strand_dead) {  // no anchor of this strand reaches min_match: the strand leaves the hit list as it is
        l_pos += (unsigned long long)cnt;
        continue;
      }
      if (__any_sync(0xFFFFFFFFu, overflow)) {
        if (!BIG) return false;  // the BIG stores hold 19 * 49 * 99 entries
      }
      const int N = cnt;
      l_pos += (unsigned long long)N;
      __syncwarp();
      // ---- 2200-2207: every segment list longer than max_hits -> no hits at all (also wipes the other strand's)
      if (min_spots > (uint32_t)a.p.max_hits) {
        tot = 0;
        continue;
      }

      build_hash(N);
      // ---- found count of every entry as an anchor (2230-2249): later segments with a position whose diagonal is
      // within max_off - 1 of the anchor's.  The segments present in the three slots around the anchor bound it from
      // above; only anchors whose bound reaches min_match (the read's true locus, rarely a chance cluster) walk the chains.
      bool relevant = false;  // some anchor reaches the running min_match
      auto exact_found = [&](const uint32_t e) {
        const int s = st_seg[e];
        const uint32_t off = (s < total_cuts) ? 16u * (uint32_t)s : (uint32_t)(len - 16);
        const long long dg = (long long)st_pos[e] + 512ll - (long long)off;
        const uint32_t bin = (uint32_t)(dg >> 4);
        uint32_t segs = 0;
#pragma unroll
        for (int db = 0; db < 3; db++) {
          const uint32_t hw = st_head[rbi_hash(bin + (uint32_t)(db - 1), tab_mask)];
          uint32_t q = hw & IMASK;
          while (q != HNIL) {
            const int sq = st_seg[q];
            if (sq > s) {
              const uint32_t offq = (sq < total_cuts) ? 16u * (uint32_t)sq : (uint32_t)(len - 16);
              const long long d = (long long)st_pos[q] + 512ll - (long long)offq - dg;
              if (d > -(long long)mo && d < (long long)mo) segs |= 1u << sq;  // 2244
            }
            const NextT nx = st_next[q];
            q = nx == NIL ? HNIL : (uint32_t)nx;
          }
        }
        const int fnd = 1 + __popc(segs);
        st_found[e] = (uint8_t)fnd;
        relevant |= fnd >= min_match;
      };
      int np = 0;  // anchors waiting in sm.pend (warp-uniform): they are walked 32 at a time, all lanes busy
      for (int e0 = 0; e0 < N; e0 += 32) {
        const int e = e0 + lane;
        bool pass = false;
        if (e < N) {
          const int s = st_seg[e];
          if (1 + max_depth - s >= min_match) {  // an anchor of segment s reaches at most 1 + max_depth - s
            pass = true;
            if (!BIG) {
              const uint32_t off = (s < total_cuts) ? 16u * (uint32_t)s : (uint32_t)(len - 16);
              const uint32_t bin = (uint32_t)(((unsigned long long)st_pos[e] + 512ull - off) >> 4);
              const uint32_t hw = st_head[rbi_hash(bin - 1u, tab_mask)] | st_head[rbi_hash(bin, tab_mask)] |
                                  st_head[rbi_hash(bin + 1u, tab_mask)];
              pass = 1 + __popc((hw >> 13) & ~((2u << s) - 1u)) >= min_match;
            }
          }
          if (!pass) st_found[e] = 0;
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
        if (pass) sm.g.pend[np + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)e;
        np += __popc(bal);
        __syncwarp();
        if (np >= 32) {
          exact_found(sm.g.pend[lane]);
          const uint32_t keep = lane < np - 32 ? sm.g.pend[32 + lane] : 0u;
          __syncwarp();
          if (lane < np - 32) sm.g.pend[lane] = keep;
          np -= 32;
          __syncwarp();
        }
      }
      if (lane < np) exact_found(sm.g.pend[lane]);
      __syncwarp();
      if (!__any_sync(0xFFFFFFFFu, relevant)) continue;  // the usual fate of the strand the read does not come from

      // ---- the reference's sequential rules over the anchors that can still matter, in its order: segment, position
      bool done = false;
      uint32_t* order = st_head;  // the chains are no longer needed
      for (int loop = 0; !done && loop <= 1 + max_depth - min_match; loop++) {
        const int off_loop = (loop < total_cuts) ? 16 * loop : len - 16;
        int R = 0;
        for (int e0 = 0; e0 < N; e0 += 32) {
          const int e = e0 + lane;
          const bool c = e < N && st_seg[e] == loop && (int)st_found[e] >= min_match;
          const unsigned bal = __ballot_sync(0xFFFFFFFFu, c);
          if (c) order[R + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)e;
          R += __popc(bal);
        }
        __syncwarp();
        if (R > 1) rbi_sort_idx(order, st_pos, R, lane);
        for (int i = 0; i < R && !done; i++) {
          const uint32_t e = order[i];
          const int f = st_found[e];
          const uint32_t apos = st_pos[e];
          if (f > min_match) {  // 2251-2260
            min_match = f;
            if (lane == 0) {
              sm.hit_pos[0] = apos;
              sm.hit_off[0] = (uint16_t)off_loop;
              sm.hit_or[0] = (uint8_t)strand;
            }
            tot = 1;
            __syncwarp();
          } else if (f == min_match) {
            if (tot < a.p.max_hits) {  // 2264-2282
              const uint32_t key = apos - (uint32_t)off_loop;
              bool dup = false;
              for (int k = lane; k < tot; k += 32) dup |= (sm.hit_pos[k] - (uint32_t)sm.hit_off[k]) == key;
              dup = __any_sync(0xFFFFFFFFu, dup);
              if (!dup) {
                if (lane == 0) {
                  sm.hit_pos[tot] = apos;
                  sm.hit_off[tot] = (uint16_t)off_loop;
                  sm.hit_or[tot] = (uint8_t)strand;
                }
                tot++;
                __syncwarp();
              }
            } else {  // 2283-2284
              done = true;
            }
          }
        }
        __syncwarp();
      }
    }
  }

  // ---- candidate windows (1047-1081) -> alignment tasks
  uint32_t base = 0;
  if (lane == 0 && tot > 0) base = atomicAdd(a.task_cursor, (uint32_t)tot);
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (tot > 0 && base + (uint32_t)tot > a.task_cap) tot = 0;  // cannot happen: cap = 200 * work items
  for (int c = lane; c < tot; c += 32) {
    const long long t = (long long)sm.hit_pos[c] - (long long)sm.hit_off[c];  // 1664-1669
    const uint32_t spot = (uint32_t)(t > 0 ? t : 0);
    Task tk;
    tk.rm = rm | ((uint32_t)sm.hit_or[c] << 31);
    tk.spot = spot;
    candidate_window(a.cstart, a.p.n_contigs, spot, len, a.p.misalign_slop, &tk.wstart, &tk.blen);
    a.tasks[base + c] = tk;
    if (tk.blen > 0) stat_cells += (unsigned long long)tk.blen * (unsigned long long)len;
  }
  if (lane == 0) {
    a.cand_base[rm] = base;
    a.cand_n[rm] = (uint32_t)tot;
    stat_cand += (unsigned long long)tot;
    stat_pos += l_pos;
    stat_lookups += l_lookups;
  }
  __syncwarp();
  return true;
}

// CAP > 0: stores in shared memory; CAP = 0: in the per-warp global scratch.  Work items are all read-mates of the
// chunk (work_list == nullptr) or the ones an earlier pass could not hold; the ones this pass cannot hold go to next_list.
template <int WARPS, int CAP>
__global__ void __launch_bounds__(WARPS * 32) k_seed_rbi(SeedRbiArgs a) {
  extern __shared__ __align__(16) unsigned char rbi_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RbiWarpSmem& sm = reinterpret_cast<RbiWarpSmem*>(rbi_smem)[warp];
  const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  unsigned long long st_pos = 0, st_cand = 0, st_cells = 0, st_lookups = 0;
  const int n_work = a.work_list ? (int)*a.work_n : (a.paired ? 2 * a.n_reads : a.n_reads);
  for (int i = gw; i < n_work; i += nw) {
    const int w = a.work_list ? (int)a.work_list[i] : i;
    if (!a.work_list) {  // the next read-mate of this warp: its row and length are cold in HBM; start fetching them now
      const int wn = i + nw;
      if (wn < n_work && lane < 4) {
        const int rn = a.paired ? (wn >> 1) : wn, mn = a.paired ? (wn & 1) : 0;
        const char* nxt = a.reads[mn] + (size_t)rn * a.stride;
        if (lane < 3) {
          if (lane * 128 < a.stride) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + lane * 128));
        } else {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.len[mn] + rn));
        }
      }
    }
    bool fit;
    if (CAP == 0) {
      unsigned char* sc = a.big_scratch + (size_t)gw * PM_RBI_BIG_BYTES;
      uint32_t* g_pos = reinterpret_cast<uint32_t*>(sc);
      uint32_t* g_next = g_pos + PM_RBI_BIG_CAP;
      uint32_t* g_head = g_next + PM_RBI_BIG_CAP;
      uint8_t* g_seg = reinterpret_cast<uint8_t*>(g_head + PM_RBI_BIG_TAB);
      uint8_t* g_found = g_seg + PM_RBI_BIG_CAP;
      fit = rbi_map_read_mate<0, uint32_t>(a, sm, g_pos, g_head, g_next, g_seg, g_found, PM_RBI_BIG_CAP, PM_RBI_BIG_TAB - 1, w,
                                           lane, st_pos, st_cand, st_cells, st_lookups);
    } else {
      constexpr int C = CAP > 0 ? CAP : 1;
      RbiSmemStore<C>& fs = reinterpret_cast<RbiSmemStore<C>*>(rbi_smem + WARPS * sizeof(RbiWarpSmem))[warp];
      fit = rbi_map_read_mate<C, uint16_t>(a, sm, fs.pos, fs.head, fs.next, fs.seg, fs.found, a.fast_cap < C ? a.fast_cap : C,
                                           (uint32_t)(C - 1), w, lane, st_pos, st_cand, st_cells, st_lookups);
    }
    if (!fit && lane == 0 && a.next_list) a.next_list[atomicAdd(a.next_cursor, 1u)] = (uint32_t)w;
    __syncwarp();
  }
  // statistics: one atomic per counter per warp (lane 0 holds the totals)
  unsigned long long cells = st_cells;
  for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xFFFFFFFFu, cells, o);
  if (lane == 0) {
    atomicAdd(&a.counters->lookups, st_lookups);
    atomicAdd(&a.counters->mer_positions, st_pos);
    atomicAdd(&a.counters->candidates, st_cand);
    atomicAdd(&a.counters->sw_cells, cells);
  }
}

template <int WARPS, int CAP>
constexpr size_t seed_rbi_smem() {
  return WARPS * sizeof(RbiWarpSmem) + (CAP > 0 ? WARPS * sizeof(RbiSmemStore<(CAP > 0 ? CAP : 1)>) : 0);
}

}  // namespace pm
