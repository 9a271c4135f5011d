// trace_int.cuh - integer traceback of the winners that are not pure diagonals (sm_100a).
//
// Replaces smith_waterman_backtrack (pemapper.c:1752-1965) for winners with gaps.  The DP of the winning
// (read, window) task is recomputed in exact integers (units of 1/36, sw_int16.cuh) with the same sub-warp
// wavefront as sw_wavefront.cuh, storing 6 decision bits per cell:
//     bits 0-1  A  = argmax_k S_k[i][j], priority 0 > 1 > 2                  (consulted from state 0, 1799-1813)
//     bit  2    X1 = S1[i][j] - ge > S0[i][j] - go                            (consulted from state 1, 1823-1831)
//     bit  3    X2 = S2[i][j] - ge > S0[i][j] - go                            (consulted from state 2, 1814-1822)
//     bit  4    the A decision compared equal integers   (S1 == S0, or S2 == max(S0, S1))
//     bit  5    an X decision compared equal integers    (S1 - ge == S0 - go, or S2 - ge == S0 - go)
// (layout in trace_walk.cuh).  The words of the PM_BAND_LANES lanes around the winner's end diagonal are kept in
// shared memory, so the walk never waits on global memory.
// The reference decides on rounded doubles with strict '>', so a comparison of rationally equal values may go
// either way.  Lane 0 therefore walks the path twice: a dry pass that only looks for a
// tie bit on a consulted decision or a step outside the band (then the winner is handed to the exact fp64
// traceback kernel instead), and, if there was none, the pass that applies the pileup increments.  Every consulted comparison then has integer
// operands that differ, i.e. doubles that differ by >= 1/36, and the walk is the reference's walk.
#pragma once
#include "pemap_common.cuh"
#include "trace_walk.cuh"

namespace pm {

struct TraceIntArgs {
  const Task* tasks;
  const TaskResult* results;   // maxk / maxi of the winner (written by the selection kernels)
  const Winner* winners;
  const uint32_t* n_items;
  Winner* exact_winners;       // winners that need the fp64 traceback
  uint32_t* exact_cursor;
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  PileSink sink;               // sink.pend: per group PM_DP_MAX bytes
  SeedCounters* counters;
  int band_half;               // lanes kept on each side of the end-diagonal lane (PM_BAND_LANES / 2)
  DevParams p;
};

template <int G, int WD>
__host__ __device__ constexpr int trace_rows() { return (G * WD + 24 < PM_DP_MAX) ? G * WD + 24 : PM_DP_MAX; }
template <int G, int WD>
__host__ __device__ constexpr size_t trace_band_bytes() { return (size_t)(128 / G) * trace_rows<G, WD>() * PM_BAND_LANES * 8; }

template <int G, int WD>
__global__ void __launch_bounds__(128) k_trace_i32(TraceIntArgs a) {
  static_assert(6 * WD <= 64, "decision bits of a lane's columns must fit one 64-bit word");
  constexpr int GPB = 128 / G;
  constexpr int ROWS = trace_rows<G, WD>();
  extern __shared__ unsigned long long s_band_i[];  // [GPB][ROWS][PM_BAND_LANES]
  __shared__ unsigned char s_win[GPB][ROWS];
  const int tid = threadIdx.x, grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) / G * G));
  const uint32_t n_items = *a.n_items;
  const uint32_t ggid = blockIdx.x * GPB + grp, n_groups = gridDim.x * GPB;
  unsigned char* win = s_win[grp];
  unsigned long long* band = s_band_i + (size_t)grp * ROWS * PM_BAND_LANES;
  const int bis = a.p.is_bisulfite;
  PileSink sink = a.sink;
  sink.pend = a.sink.pend + (size_t)ggid * PM_DP_MAX;

  for (uint32_t item = ggid; item < n_items; item += n_groups) {
    const uint32_t task_id = a.winners[item].task;
    const Task tk = a.tasks[task_id];
    const TaskResult res = a.results[task_id];
    const int orient = (int)(tk.rm >> 31);
    const uint32_t rm = tk.rm & 0x7FFFFFFFu;
    const int mm = ((rm & 1u) ? a.len[1] : a.len[0])[rm >> 1];
    const char* read = ((rm & 1u) ? a.reads[1] : a.reads[0]) + (size_t)(rm >> 1) * a.stride;
    const int nn = res.maxi < tk.blen ? res.maxi : tk.blen;  // rows below the winning cell are never consulted
    const int dend = res.maxi - mm;

    __syncwarp(gmask);
    // one-hot base codes: two bases match iff their codes share a bit; 0 = outside ACGTN (the scoring kernel sent
    // such reads to the fp64 path; their traceback goes there too)
    bool bad = false;
    for (int i = gl; i < nn; i += G) {
      const char ch = a.genome[(size_t)tk.wstart + i];
      unsigned c = ch == 'A' ? 1u : ch == 'C' ? (bis ? 10u : 2u) : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
      bad |= (c == 0u);
      win[i] = (unsigned char)c;
    }
    unsigned q[WD];
    int s0u[WD], s1u[WD], mu[WD];
    const int jbase = gl * WD;
#pragma unroll
    for (int c = 0; c < WD; c++) {
      const int j0 = jbase + c;
      unsigned qc = 0;
      if (j0 < mm) {
        const char ch = oriented_char(read, mm, orient, j0);
        qc = ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
        bad |= (qc == 0u);
      }
      q[c] = qc;
      const int b = -(72 + j0);  // S*[0][j] = -(go + (j-1) ge), j = j0 + 1 (2073-2081)
      s0u[c] = b;
      s1u[c] = b;
      mu[c] = b;
    }
    bad = __any_sync(gmask, bad);
    int out_s0 = 0, out_s2 = 0, out_m = 0;
    __syncwarp(gmask);

    const int steps = nn > 0 ? nn + G - 1 : 0;
    for (int s = 0; s < steps; s++) {
      int l_s0 = __shfl_up_sync(gmask, out_s0, 1, G);
      int l_s2 = __shfl_up_sync(gmask, out_s2, 1, G);
      int diag = __shfl_up_sync(gmask, out_m, 1, G);
      if (gl == 0) {  // column 0 (2062-2081)
        l_s0 = 0;
        l_s2 = -72;
        diag = 0;
      }
      const int i = s - gl + 1;
      if (i >= 1 && i <= nn) {
        const unsigned rc = win[i - 1];
        unsigned long long dword = 0;
#pragma unroll
        for (int c = 0; c < WD; c++) {
          const int s2 = max(l_s0 - 72, l_s2 - 1);
          const int s1 = max(s0u[c] - 72, s1u[c] - 1);
          const int s0 = diag + ((q[c] & rc) ? 36 : -12);
          diag = mu[c];
          const int m01 = max(s0, s1);
          const int m = max(m01, s2);
          const unsigned ak = (s2 > m01) ? 2u : (s1 > s0) ? 1u : 0u;
          const int opn = s0 - 72;
          const unsigned bits = ak | ((s1 - 1 > opn) ? 4u : 0u) | ((s2 - 1 > opn) ? 8u : 0u) |
                                ((s1 == s0 || s2 == m01) ? 16u : 0u) | ((s1 - 1 == opn || s2 - 1 == opn) ? 32u : 0u);
          dword |= (unsigned long long)bits << (6 * c);
          s0u[c] = s0;
          s1u[c] = s1;
          mu[c] = m;
          l_s0 = s0;
          l_s2 = s2;
        }
        out_s0 = l_s0;
        out_s2 = l_s2;
        out_m = diag;
        const int slot = gl - (band_center_lane<WD>(i, dend) - a.band_half);
        if (slot >= 0 && slot <= 2 * a.band_half) band[(i - 1) * PM_BAND_LANES + slot] = dword;
      }
    }
    __syncwarp(gmask);

    if (gl == 0 && nn > 0) {
      BandCell<WD, 6> cell;
      cell.band = band;
      cell.dend = dend;
      cell.half = a.band_half;
      int rc = bad ? PM_WALK_TIE : walk_path<false, 1>(cell, res.maxk, res.maxi, mm, read, mm, orient, tk.wstart, sink);
      if (rc == PM_WALK_OK) {
        walk_path<true, 0>(cell, res.maxk, res.maxi, mm, read, mm, orient, tk.wstart, sink);
        atomicAdd(&a.counters->tb_cells, (unsigned long long)nn * (unsigned long long)mm);
      } else {
        const uint32_t w = atomicAdd(a.exact_cursor, 1u);
        a.exact_winners[w] = a.winners[item];
        atomicAdd(&a.counters->exact_traced, 1ull);
      }
    }
    __syncwarp(gmask);
  }
}

}  // namespace pm
