// pemap_common.cuh - shared device types and helpers for the sm_100a PEMapper hot path.
// Reference line numbers are into wingolab-org/pecaller src/pemapper.c unless a file is named.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PM_MAX_SEG 19        // total_cuts+1 for reads up to 299 bp (pemapper.c:1573-1587)
#define PM_KV 49             // exact 16-mer + 48 one-substitution neighbours (fill_mers 1969-2003)
#define PM_SEG_CAP 8192      // >= 49*99 positions per segment, rounded to a power of two for the bitonic sort
#define PM_MAX_HITS 200      // max_hits (pemapper.c:162)
#define PM_DP_MAX 320        // padded read columns / window rows a DP kernel can be instantiated for

// -DPM_TIE_DEBUG: count why winners leave the integer traceback for the fp64 kernel (printed by pemap_destroy)
#ifdef PM_TIE_DEBUG
namespace pm { __device__ unsigned long long g_tie_why[32]; }
#define PM_WHY(x) atomicAdd(&pm::g_tie_why[x], 1ull)
#else
#define PM_WHY(x)
#endif

namespace pm {

struct DevParams {
  // scoring constants, IEEE doubles computed on the host with the reference's own expressions (2011, 2039-2040)
  double match, mism, go, ge;
  double min_align, match_bonus;
  int idepth, max_hits, too_many_spots, is_bisulfite, pair_flag, min_dist, max_dist, misalign_slop;
  int n_contigs;
  uint32_t genome_size_lo;  // genome_size fits in 32 bits for every genome the reference can index
};

// One (read-mate, candidate locus) alignment task = one smith_waterman_align call (1098, 1145, 1367, 1376).
struct Task {
  uint32_t rm;      // read-mate slot: 2*read + mate; bit 31 = orientation (0 forward, 1 reverse strand)
  uint32_t spot;    // candidate position in index coordinates (1664-1669)
  uint32_t wstart;  // window start in real coordinates (1055 / 1073)
  int32_t blen;     // window length nn (1058 / 1076)
};

struct TaskResult {
  double score;     // S[maxk][maxi][mm] (1747)
  int32_t maxi;     // start[1]
  int32_t maxk;     // start[0]
};

// Winner of a read-mate after the selection rules: which task gets the traceback.
struct Winner {
  uint32_t task;    // index into the task array
  uint32_t rm;      // read-mate slot
};

struct SeedCounters {  // device-side statistics, accumulated with atomics once per warp
  unsigned long long lookups, mer_positions, candidates, sw_cells, tb_cells, replayed, diag_traced, exact_traced, tb_cells_int,
      sw_cells_certified;  // cells of the candidates k_diag_certify resolved without the DP
};

__host__ __device__ __forceinline__ double dmax(double a, double b) { return (a > b) ? a : b; }  // maxim(), pemapper.c:36

// reverse_transcribe (2303-2337) for one character
__host__ __device__ __forceinline__ char rt_char(char ch) {
  switch (ch) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'W': return 'W';
    case 'S': return 'S';
    case 'K': return 'M';
    case 'M': return 'K';
    case 'Y': return 'R';
    case 'R': return 'Y';
    default: return 'N';
  }
}

// cv[] of fill_cv_mat (2379-2383): c/C=1, g/G=2, t/T=3, everything else (N included) = 0
__host__ __device__ __forceinline__ uint32_t base_code(char ch) {
  return (ch == 'C' || ch == 'c') ? 1u : (ch == 'G' || ch == 'g') ? 2u : (ch == 'T' || ch == 't') ? 3u : 0u;
}

// bonus matrix entry (init_bonus_matrices 2006-2035): equal chars or an N/n on either side match
__host__ __device__ __forceinline__ bool bases_match(char ref, char q, int bisulfite) {
  bool m = (ref == q) || ref == 'N' || ref == 'n' || q == 'N' || q == 'n';
  if (bisulfite) m = m || ((ref == 'C' || ref == 'c') && (q == 'T' || q == 't'));
  return m;
}

// find_chrom (2168-2186): bisection that starts at index 7 and is inclusive on both ends.
// pos must have at least 9 readable entries when n_contigs > 1 (the reference reads pos[7], pos[8]).
__host__ __device__ __forceinline__ int find_chrom(const uint32_t* pos, int n_contigs, uint32_t v) {
  int first = 0, last = n_contigs - 1, probe = 7;
  for (int guard = 0; guard < 64; guard++) {
    if (first == last) return first;
    uint32_t a = pos[probe], b = pos[probe + 1];
    if (a <= v && b >= v) return probe;
    if (a > v) last = probe - 1; else first = probe + 1;
    probe = (last + first) / 2;
    if (probe < 0 || probe >= n_contigs) return first < 0 ? 0 : (first >= n_contigs ? n_contigs - 1 : first);
  }
  return first;
}

// Dynamic work distribution for the persistent sub-warp kernels: the group's lane 0 takes the next item from a global
// counter and broadcasts it.  Items differ in cost (window length, walk length, tie certification), and with a few
// items per group a static grid-stride split leaves most of the last wave idle.
template <int G>
__device__ __forceinline__ uint32_t next_work_item(uint32_t* counter, unsigned gmask, int gl) {
  uint32_t v = 0;
  if (gl == 0) v = atomicAdd(counter, 1u);
  return __shfl_sync(gmask, v, 0, G);
}

// Same, for kernels whose groups end every item with a one-lane phase (the traceback walk): the sub-warps of a warp
// take consecutive items together, so that they stay in lock step through the wavefront (a warp whose sub-warps
// sit in different phases issues each phase for half of its lanes only).
// *first = the warp's first item (warp-uniform: the loop ends for the whole warp when it passes the end).
template <int G>
__device__ __forceinline__ uint32_t next_work_item_warp(uint32_t* counter, uint32_t* first) {
  __syncwarp();
  uint32_t v = 0;
  if ((threadIdx.x & 31) == 0) v = atomicAdd(counter, 32u / G);
  v = __shfl_sync(0xFFFFFFFFu, v, 0);
  *first = v;
  return v + (threadIdx.x & 31) / G;
}

// window set-up of map_everything (1052-1058): returns the contig, fills start (real coordinate) and blen
__host__ __device__ __forceinline__ int candidate_window(const uint32_t* cstart, int n_contigs, uint32_t spot, int len,
                                                          int slop, uint32_t* start, int32_t* blen) {
  int ch = find_chrom(cstart, n_contigs, spot);
  uint32_t extra = 15u * (uint32_t)ch;
  long long t = (long long)extra + (long long)spot - (long long)slop;
  if (t < 0) t = 0;
  long long lo = (long long)(uint32_t)(cstart[ch] + extra);
  uint32_t s = (uint32_t)(lo > t ? lo : t);
  uint32_t e1 = cstart[ch + 1] + extra, e2 = extra + spot + (uint32_t)len + (uint32_t)slop;
  uint32_t e = e1 < e2 ? e1 : e2;
  *start = s;
  *blen = (int32_t)(1u + e - s);
  return ch;
}

}  // namespace pm
