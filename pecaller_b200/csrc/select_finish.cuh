// select_finish.cuh - per-read selection rules, pileup compaction, device index build (sm_100a).
#pragma once
#include "pemap_common.cuh"

namespace pm {

// ---------------------------------------------------------------------------------------------------
// Selection: one thread per read / pair.  Replaces map_everything 1084-1192 and find_mate_pairs 1313-1536
// (the SW calls themselves, 1361-1379, were done by the score kernel).  All comparisons are on the doubles
// the score kernel produced, with the reference's expressions and evaluation order.
// ---------------------------------------------------------------------------------------------------

struct SelectArgs {
  const Task* tasks;
  const TaskResult* results;
  const uint32_t* cand_base;
  const uint32_t* cand_n;
  const int* len[2];
  int n_reads;
  const uint32_t* read_list;   // optional: only these reads (replay of integer ties); count in *n_list
  const uint32_t* n_list;
  uint32_t* m1;
  uint32_t* m2;
  int* mapping_type;
  // detail (optional, may be null)
  int32_t* det_best;     // [2*n]
  int32_t* det_orient;   // [2*n]
  double* det_score;     // [2*n]
  Winner* winners;
  uint32_t* winner_cursor;
  DevParams p;
};

enum { T_UNIQUE_MATE = 0, T_UNIQUE_SLIP, T_UNIQUE_SINGLE, T_UNIQUE_MIS, T_NON_MATE, T_NON_MIS, T_FRAG_MIS, T_NON_NO,
       T_NEITHER_MAP };

// "only one mate has candidates" rule (1084-1128 / 1130-1174)
__device__ __forceinline__ int single_rule(const TaskResult* res, int n, int len, const DevParams& p, int* best) {
  const double good = __dmul_rn(__dmul_rn((double)len, p.min_align), p.match_bonus);
  double top = __dmul_rn(-p.go, (double)len);
  int count = 0;
  *best = -1;
  for (int q = 0; q < n; q++) {
    const double s = res[q].score;
    if (s > top && s >= good) {
      top = s;
      count = 1;
      *best = q;
    } else if (fabs(__dsub_rn(s, top)) < 0.0001 && count > 0) {
      count++;
    }
  }
  if (count == 0) { *best = -1; return T_NEITHER_MAP; }
  if (count == 1) return T_UNIQUE_SINGLE;
  *best = -1;
  return T_NON_NO;
}

// smax arrays of find_mate_pairs are dvector(0, max_hits) initialised to -1.0 (1346-1351)
__device__ __forceinline__ double smax_at(const TaskResult* res, int n, int i) { return i < n ? res[i].score : -1.0; }

__device__ __forceinline__ int pair_rule(const Task* ta, const TaskResult* ra, int n1, int l1, const Task* tb,
                                         const TaskResult* rb, int n2, int l3, const DevParams& p, int* keep1, int* keep2) {
  const double good1 = __dmul_rn(__dmul_rn((double)l1, p.min_align), p.match_bonus);
  const double good2 = __dmul_rn(__dmul_rn((double)l3, p.min_align), p.match_bonus);
  double tot_best = -1e5;
  int perfect = 0, slip = 0, sm1 = -1, sm2 = -1;
  *keep1 = *keep2 = -1;
  for (int w1 = 0; w1 < n1; w1++) {
    const double a1 = ra[w1].score;
    if (!(a1 >= good1)) continue;  // 1383
    const long long p1 = (long long)ta[w1].spot;
    const int or1 = (int)(ta[w1].rm >> 31);
    for (int w2 = 0; w2 < n2; w2++) {
      const double a2 = rb[w2].score;
      if (!(a2 >= good2)) continue;  // 1386
      long long dist = p1 - (long long)tb[w2].spot;  // 1388-1394: index coordinates, no contig check
      if (dist < 0) dist = -dist;
      const int or2 = (int)(tb[w2].rm >> 31);
      if (!(dist >= p.min_dist && dist <= p.max_dist && or1 != or2)) continue;
      const double inc = __dsub_rn(__dadd_rn(a1, a2), tot_best);  // 1400
      if (inc > 0.001) {
        perfect = 1;
        sm1 = w1;
        sm2 = w2;
        tot_best = __dadd_rn(a1, a2);
        slip = 1;
      } else if (inc > -0.001) {
        if (sm1 == w1 || sm2 == w2) slip++;
        perfect++;
      }
    }
  }
  if (perfect > 0) {  // 1424-1447
    if (perfect == 1) { *keep1 = sm1; *keep2 = sm2; return T_UNIQUE_MATE; }
    if (slip == perfect) { *keep1 = sm1; *keep2 = sm2; return T_UNIQUE_SLIP; }
    return T_NON_MATE;
  }
  int best1 = 0, best2 = 0, c1 = 0, c2 = 0;  // 1450-1469 (quirks kept: counters start at 0; smax2[best1])
  for (int i = 1; i < n1; i++) {
    if (ra[i].score > ra[best1].score) { best1 = i; c1 = 1; }
    else if (__dsub_rn(ra[i].score, ra[best1].score) > -0.0001) c1++;
  }
  for (int i = 1; i < n2; i++) {
    if (rb[i].score > rb[best2].score) { best2 = i; c2 = 1; }
    else if (__dsub_rn(rb[i].score, smax_at(rb, n2, best1)) > -0.0001) c2++;
  }
  const bool ok2 = (rb[best2].score >= good2) && (c2 < 2);
  if (ra[best1].score >= good1 && c1 < 2) {  // 1483-1527
    *keep1 = best1;
    if (ok2) { *keep2 = best2; return T_UNIQUE_MIS; }
    return T_UNIQUE_SINGLE;
  }
  if (ok2) { *keep2 = best2; return T_UNIQUE_SINGLE; }
  return T_NON_MIS;
}

__global__ void __launch_bounds__(128) k_select(SelectArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int r = idx;
  if (a.read_list) {
    if ((uint32_t)idx >= *a.n_list) return;
    r = (int)a.read_list[idx];
  } else if (idx >= a.n_reads) return;
  const int n1 = (int)a.cand_n[2 * r], n2 = a.p.pair_flag ? (int)a.cand_n[2 * r + 1] : 0;
  const uint32_t b1 = a.cand_base[2 * r], b2 = a.p.pair_flag ? a.cand_base[2 * r + 1] : 0;
  const int l1 = a.len[0][r], l3 = a.p.pair_flag ? a.len[1][r] : 0;
  int keep1 = -1, keep2 = -1, call;
  if (n1 > 0 && n2 == 0) call = single_rule(a.results + b1, n1, l1, a.p, &keep1);
  else if (n2 > 0 && n1 == 0) call = single_rule(a.results + b2, n2, l3, a.p, &keep2);
  else if (n1 > 0 && n2 > 0)
    call = pair_rule(a.tasks + b1, a.results + b1, n1, l1, a.tasks + b2, a.results + b2, n2, l3, a.p, &keep1, &keep2);
  else call = T_NEITHER_MAP;  // 1186-1192
  uint32_t m1 = 0, m2 = 0;
  if (keep1 >= 0) {  // 1207-1208: bn[start1[1]].pos + 1
    m1 = a.tasks[b1 + keep1].wstart + (uint32_t)a.results[b1 + keep1].maxi + 1u;
    uint32_t w = atomicAdd(a.winner_cursor, 1u);
    a.winners[w].task = b1 + (uint32_t)keep1;
    a.winners[w].rm = 2u * (uint32_t)r;
  }
  if (keep2 >= 0) {  // 1227-1228
    m2 = a.tasks[b2 + keep2].wstart + (uint32_t)a.results[b2 + keep2].maxi + 1u;
    uint32_t w = atomicAdd(a.winner_cursor, 1u);
    a.winners[w].task = b2 + (uint32_t)keep2;
    a.winners[w].rm = 2u * (uint32_t)r + 1u;
  }
  a.m1[r] = m1;
  a.m2[r] = m2;
  a.mapping_type[r] = call;
  if (a.det_best) {
    a.det_best[2 * r] = keep1;
    a.det_best[2 * r + 1] = keep2;
    a.det_orient[2 * r] = keep1 >= 0 ? (int)(a.tasks[b1 + keep1].rm >> 31) : -1;
    a.det_orient[2 * r + 1] = keep2 >= 0 ? (int)(a.tasks[b2 + keep2].rm >> 31) : -1;
    a.det_score[2 * r] = keep1 >= 0 ? a.results[b1 + keep1].score : 0.0;
    a.det_score[2 * r + 1] = keep2 >= 0 ? a.results[b2 + keep2].score : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------
// Pileup compaction: replaces the writer loop of main() (828-842).  Counters are uint32 on the device and are
// truncated to the reference's unsigned short once, here (sum mod 2^32 mod 2^16 == sum mod 2^16).
// ---------------------------------------------------------------------------------------------------

struct PileRecord {
  uint32_t pos;
  uint16_t c[6];
};

#define PM_COMPACT_BLOCK 256
#define PM_COMPACT_ITEMS 8

__device__ __forceinline__ bool site_covered(const uint32_t* c6) {
  uint32_t t = (c6[0] & 0xFFFFu) + (c6[1] & 0xFFFFu) + (c6[2] & 0xFFFFu) + (c6[3] & 0xFFFFu) + (c6[4] & 0xFFFFu) +
               (c6[5] & 0xFFFFu);
  return t > 0;  // 829-831
}

// pass 1: covered sites per tile of PM_COMPACT_BLOCK*PM_COMPACT_ITEMS sites.  Both passes work on a window of the
// genome, sites [site0, site0 + n_sites): pemap_finish_stream compacts window by window through a bounded buffer.
__global__ void __launch_bounds__(PM_COMPACT_BLOCK) k_compact_count(const uint32_t* counts, uint64_t site0, uint64_t n_sites,
                                                                    unsigned long long* tile_count) {
  const uint64_t tile0 = (uint64_t)blockIdx.x * (PM_COMPACT_BLOCK * PM_COMPACT_ITEMS);
  int n = 0;
  for (int k = 0; k < PM_COMPACT_ITEMS; k++) {
    uint64_t s = tile0 + (uint64_t)k * PM_COMPACT_BLOCK + threadIdx.x;
    if (s < n_sites) n += site_covered(counts + (site0 + s) * 6) ? 1 : 0;
  }
  n = __reduce_add_sync(0xFFFFFFFFu, n);
  __shared__ int ws[PM_COMPACT_BLOCK / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < PM_COMPACT_BLOCK / 32; i++) t += ws[i];
    tile_count[blockIdx.x] = (unsigned long long)t;
  }
}

// pass 2: tile_off = exclusive scan of tile_count (a device scan between the passes); write the window's records in order
__global__ void __launch_bounds__(PM_COMPACT_BLOCK) k_compact_write(const uint32_t* counts, uint64_t site0, uint64_t n_sites,
                                                                    const unsigned long long* tile_off, PileRecord* out) {
  const uint64_t tile0 = (uint64_t)blockIdx.x * (PM_COMPACT_BLOCK * PM_COMPACT_ITEMS);
  __shared__ uint32_t warp_tot[PM_COMPACT_BLOCK / 32];
  __shared__ uint64_t running;
  if (threadIdx.x == 0) running = tile_off[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < PM_COMPACT_ITEMS; k++) {
    const uint64_t s = tile0 + (uint64_t)k * PM_COMPACT_BLOCK + threadIdx.x;
    uint32_t c6[6] = {0, 0, 0, 0, 0, 0};
    bool cov = false;
    if (s < n_sites) {
#pragma unroll
      for (int i = 0; i < 6; i++) c6[i] = counts[(site0 + s) * 6 + i];
      cov = site_covered(c6);
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, cov);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    uint64_t base = running;
    for (int w = 0; w < warp; w++) base += warp_tot[w];
    if (cov) {
      PileRecord rec;
      rec.pos = (uint32_t)(site0 + s);
#pragma unroll
      for (int i = 0; i < 6; i++) rec.c[i] = (uint16_t)(c6[i] & 0xFFFFu);
      out[base + __popc(bal & ((1u << lane) - 1u))] = rec;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t t = 0;
      for (int w = 0; w < PM_COMPACT_BLOCK / 32; w++) t += warp_tot[w];
      running += t;
    }
    __syncthreads();
  }
}

// counts[i] += peer[i]; `peer` is another GPU's counter array mapped through NVLink peer access (16-byte loads)
__global__ void __launch_bounds__(256) k_add_peer_counts(uint32_t* counts, const uint32_t* peer, uint64_t n_words) {
  const uint64_t n4 = n_words >> 2;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 p = reinterpret_cast<const uint4*>(peer)[i];
    uint4 c = reinterpret_cast<uint4*>(counts)[i];
    c.x += p.x; c.y += p.y; c.z += p.z; c.w += p.w;
    reinterpret_cast<uint4*>(counts)[i] = c;
  }
  for (uint64_t i = (n4 << 2) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) counts[i] += peer[i];
}

// counts[w] += sum over peers of peer[w] for the words [first, first + n) of this GPU's slice: the peers' arrays are
// mapped through NVLink (CUDA IPC or in-process peer access), 16-byte loads, up to 15 peers
struct PeerPtrs {
  const uint32_t* p[15];
  int n;
};
// NP = number of peers (compile time: every peer's load is issued before the first add, and two 16-byte words per
// thread and peer are in flight - a remote load takes microseconds, a dependent chain of them is what made the first
// version run at 65 GB/s per GPU on 8 GPUs); NP = 0: any number, one load at a time.
template <int NP>
__global__ void __launch_bounds__(256) k_reduce_slice(uint32_t* counts, PeerPtrs peers, uint64_t first, uint64_t n_words) {
  const uint64_t n4 = n_words >> 2;  // first and n_words are multiples of 4 (slices are cut at compaction tiles)
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint4* mine = reinterpret_cast<uint4*>(counts + first);
  if (NP == 0) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      uint4 c = mine[i];
      for (int k = 0; k < peers.n; k++) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(peers.p[k] + first) + i);
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
      }
      mine[i] = c;
    }
    return;
  }
  constexpr int P = NP > 0 ? NP : 1;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const uint64_t j = i + stride;
    const bool two = j < n4;
    uint4 v[P][2];
#pragma unroll
    for (int k = 0; k < P; k++) {
      const uint4* src = reinterpret_cast<const uint4*>(peers.p[k] + first);
      v[k][0] = __ldcs(src + i);
      v[k][1] = two ? __ldcs(src + j) : make_uint4(0, 0, 0, 0);
    }
    uint4 c0 = mine[i], c1 = two ? mine[j] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < P; k++) {
      c0.x += v[k][0].x; c0.y += v[k][0].y; c0.z += v[k][0].z; c0.w += v[k][0].w;
      c1.x += v[k][1].x; c1.y += v[k][1].y; c1.z += v[k][1].z; c1.w += v[k][1].w;
    }
    mine[i] = c0;
    if (two) mine[j] = c1;
  }
}

// ---------------------------------------------------------------------------------------------------
// Device index build: what index_genome_whole.c computes (169-177 codes, 248-299 rolling k-mer with N reset,
// 213-216/271 index coordinates, 334-342 prefix table), as data-parallel passes.
// ---------------------------------------------------------------------------------------------------

// one thread per candidate k-mer start (real coordinate x); contig c of x found by bisection over real starts
__global__ void __launch_bounds__(256) k_index_kmers(const char* genome, uint64_t genome_size, const uint64_t* real_start,
                                                     int n_contigs, int bisulfite, uint32_t* key, uint32_t* val,
                                                     unsigned char* flag) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= genome_size) return;
  int lo = 0, hi = n_contigs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (real_start[mid] <= x) lo = mid; else hi = mid - 1;
  }
  const uint64_t cend = real_start[lo + 1];
  bool ok = x + 16 <= cend;
  uint32_t code = 0;
  if (ok) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const char ch = genome[x + i];
      if (ch == 'N') ok = false;
      uint32_t b = ch == 'G' ? 2u : ch == 'T' ? 3u : ch == 'C' ? (bisulfite ? 3u : 1u) : 0u;
      code = (code << 2) | b;
    }
  }
  flag[x] = ok ? 1 : 0;
  key[x] = code;
  val[x] = (uint32_t)(x - 15ull * (uint64_t)lo);  // index coordinate = real coordinate - 15 * contig (271)
}

// pos_index[w] = number of indexed k-mers with code < w, for w in [first, first+n)  (334-342)
__global__ void __launch_bounds__(256) k_index_prefix(const uint32_t* sorted_keys, uint64_t n_mers, uint32_t* pos_index,
                                                      uint64_t first, uint64_t n) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t w = first + t;
  uint64_t a = 0, b = n_mers;
  if (w > 0xFFFFFFFFull) a = n_mers;
  else {
    const uint32_t k = (uint32_t)w;
    while (a < b) {
      uint64_t m = (a + b) >> 1;
      if (sorted_keys[m] < k) a = m + 1; else b = m;
    }
  }
  pos_index[w] = (uint32_t)a;
}

}  // namespace pm
