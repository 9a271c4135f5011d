// sw_wavefront.cuh - exact (IEEE double) semi-global affine-gap Smith-Waterman, sub-warp wavefront (sm_100a).
//
// Replaces smith_waterman_align (pemapper.c:1694-1748) and, with TRACE, smith_waterman_backtrack (1752-1965).
// A group of G lanes owns one (read, window) task; lane t owns read columns t*WD+1 .. t*WD+WD and sweeps the
// window rows with a skew of one row per lane, so the three DP states of a column live in registers and the
// column boundary (S0, S2 of the row, M of the row above) moves to the next lane by shuffle.  The window is
// staged in shared memory.  Every cell evaluates the reference's own expressions in double:
//     S2[i][j] = max(S0[i][j-1]-go, S2[i][j-1]-ge)            (1710 / 1720)
//     S1[i][j] = max(S0[i-1][j]-go, S1[i-1][j]-ge)            (1711 / 1721)
//     S0[i][j] = max(S0,S1,S2)[i-1][j-1] + bump               (1713 / 1723; rounding is monotone, so adding
//                                                              bump after the max gives the same double)
// so scores, the last-column argmax (1717-1742) and the traceback predicates are bit-identical to the CPU.
// MODE 0 scores.  MODE 1/2 trace: 4 decision bits per cell (layout in trace_walk.cuh) are
// stored and walked (MODE 2: by the whole group, coop_walk in trace_walk.cuh; MODE 1: by lane 0), applying the pileup
// increments with atomics.  MODE 2 keeps the band around the
// winner's end diagonal in shared memory (a walk that leaves it hands the winner to the MODE 1 kernel through the
// `oob` list); MODE 1 keeps every lane's word in a global scratch and can walk anywhere.
#pragma once
#include "pemap_common.cuh"
#include "trace_walk.cuh"

namespace pm {

struct SwArgs {
  const Task* tasks;
  TaskResult* results;         // score kernel: written; trace kernels: read (maxk/maxi of the winner)
  const Winner* winners;       // trace kernels: the winners; score kernel with list_mode: the tasks to (re)score
  int list_mode;
  const uint32_t* n_items;     // device counter: number of tasks (score) or winners (trace)
  uint32_t* work;              // zeroed per launch: next item
  Winner* oob_winners;         // MODE 2: winners whose walk left the shared-memory band
  uint32_t* oob_cursor;
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  const double* border;        // S*[0][j], j < PM_DP_MAX (pemapper.c:2073-2081)
  unsigned long long* dirs;    // MODE 1 scratch: per group PM_DP_MAX * G words
  PileSink sink;               // sink.pend: per group PM_DP_MAX bytes
  SeedCounters* counters;
  int band_half;               // lanes kept on each side of the end-diagonal lane (PM_BAND_LANES / 2)
  DevParams p;
};

__device__ __forceinline__ char seq_char(const char* read, int len, int orient, int j0) {
  return oriented_char(read, len, orient, j0);
}

template <int G, int WD, int MODE>
__global__ void __launch_bounds__(128, MODE == 2 ? 4 : 1) k_sw_fp64(SwArgs a) {
  constexpr bool TRACE = MODE != 0;
  constexpr int GROUPS_PER_BLOCK = 128 / G;
  constexpr int ROWS = (G * WD + 24 < PM_DP_MAX) ? G * WD + 24 : PM_DP_MAX;
  extern __shared__ unsigned long long s_band_d[];  // MODE 2: [GROUPS_PER_BLOCK][ROWS][PM_BAND_LANES]
  __shared__ char s_win[GROUPS_PER_BLOCK][PM_DP_MAX];
  __shared__ unsigned char s_wk[GROUPS_PER_BLOCK][PM_DP_MAX];  // one-hot codes of the window (0 = outside ACGTN)
  __shared__ unsigned char s_q[TRACE ? GROUPS_PER_BLOCK : 1][TRACE ? PM_DP_MAX : 1];  // one-hot codes of the oriented read
  const int tid = threadIdx.x;
  const int grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) / G * G));
  const uint32_t n_items = *a.n_items;
  const uint32_t ggid = blockIdx.x * GROUPS_PER_BLOCK + grp;
  const double go = a.p.go, ge = a.p.ge, match = a.p.match, mism = a.p.mism;
  char* win = s_win[grp];
  unsigned long long* band = MODE == 2 ? s_band_d + (size_t)grp * ROWS * PM_BAND_LANES : nullptr;
  unsigned long long* dirs = MODE == 1 ? a.dirs + (size_t)ggid * PM_DP_MAX * G : nullptr;
  PileSink sink = a.sink;
  if (TRACE) sink.pend = a.sink.pend + (size_t)ggid * PM_DP_MAX;

  // TRACE: the sub-warps of a warp take consecutive winners and run the loops below with warp-uniform trip counts and
  // full-mask shuffles (width G), so that they stay in lock step through the wavefront; a sub-warp without a winner
  // runs empty rows.
  constexpr unsigned FULL = 0xFFFFFFFFu;
  for (;;) {
    uint32_t first = 0;
    uint32_t item = TRACE ? next_work_item_warp<G>(a.work, &first) : next_work_item<G>(a.work, gmask, gl);
    bool have = true;
    if (TRACE) {
      if (first >= n_items) break;
      have = item < n_items;
      if (!have) item = first;
    } else if (item >= n_items) {
      break;
    }
    const uint32_t task_id = (TRACE || a.list_mode) ? a.winners[item].task : item;
    const Task tk = a.tasks[task_id];
    const int orient = (int)(tk.rm >> 31);
    const uint32_t rm = tk.rm & 0x7FFFFFFFu;
    const int mate = (int)(rm & 1u);
    const uint32_t r = rm >> 1;
    const int mm = (mate ? a.len[1] : a.len[0])[r];
    const char* read = (mate ? a.reads[1] : a.reads[0]) + (size_t)r * a.stride;
    TaskResult res;
    res.score = 0.0;
    res.maxi = 0;
    res.maxk = 0;
    if (TRACE) res = a.results[task_id];
    // the walk never consults rows below the winning cell
    const int nn = !have ? 0 : TRACE ? (res.maxi < tk.blen ? res.maxi : tk.blen) : tk.blen;
    const int dend = res.maxi - mm;
    const unsigned wmask = TRACE ? FULL : gmask;
    int steps = nn > 0 ? nn + G - 1 : 0;
    if (TRACE && G < 32) steps = max(steps, __shfl_xor_sync(FULL, steps, 16));

    __syncwarp(gmask);
    const int bis = a.p.is_bisulfite;
    for (int i = gl; i < nn; i += G) {
      const char ch = a.genome[(size_t)tk.wstart + i];
      win[i] = ch;
      unsigned code = base_onehot(ch);
      if (bis && ch == 'C') code = 10u;  // reference C also matches read T (2026-2033)
      s_wk[grp][i] = (unsigned char)code;
    }
    // this lane's read characters (reverse strand = reverse_transcribe'd read, 1021/1098)
    // Bases match iff their one-hot codes share a bit (N = all bits); a code of 0 (any character outside ACGTN,
    // lower case included) sends the row to the character comparison of init_bonus_matrices (2006-2035).
    char q[WD];
    unsigned qk[WD];
    bool lane_other = false;
    const int jbase = gl * WD;  // columns jbase+1 .. jbase+WD
#pragma unroll
    for (int c = 0; c < WD; c++) {
      int j0 = jbase + c;
      q[c] = (j0 < mm) ? seq_char(read, mm, orient, j0) : (char)0;
      qk[c] = base_onehot(q[c]);
      lane_other |= (j0 < mm) && qk[c] == 0u;
      if (TRACE && j0 < PM_DP_MAX) s_q[grp][j0] = (unsigned char)qk[c];
    }
    // row 0 (init_penalty_matrices 2073-2081): S0 = S1 = S2 = M = border[j]
    double s0u[WD], s1u[WD], mu[WD];
#pragma unroll
    for (int c = 0; c < WD; c++) {
      double b = a.border[jbase + c + 1];
      s0u[c] = b;
      s1u[c] = b;
      mu[c] = b;
    }
    const int cmm = ((mm - 1) / WD == gl) ? (mm - 1) % WD : -1;  // which of my columns is the read's last
    double best = a.border[mm];  // S[0][0][mm] (1701-1703)
    int bk = 0, bi = 0;
    double out_s0 = 0.0, out_s2 = 0.0, out_m = 0.0;
    __syncwarp(gmask);

    for (int s = 0; s < steps; s++) {
      double in_s0 = __shfl_up_sync(wmask, out_s0, 1, G);
      double in_s2 = __shfl_up_sync(wmask, out_s2, 1, G);
      double in_m = __shfl_up_sync(wmask, out_m, 1, G);
      if (gl == 0) {  // column 0: S0[i][0] = 0, S2[i][0] = -go, M[i-1][0] = max(0, 0, -go) = 0 (2062-2081)
        in_s0 = 0.0;
        in_s2 = -1.0 * go;
        in_m = 0.0;
      }
      const int i = s - gl + 1;
      if (i >= 1 && i <= nn) {
        const char rc = win[i - 1];
        const unsigned rk = s_wk[grp][i - 1];
        const bool by_char = lane_other || rk == 0u;
        double l_s0 = in_s0, l_s2 = in_s2, diag = in_m;
        unsigned long long dword = 0;
#pragma unroll
        for (int c = 0; c < WD; c++) {
          const double s2 = dmax(l_s0 - go, l_s2 - ge);
          const double s1 = dmax(s0u[c] - go, s1u[c] - ge);
          const bool same = by_char ? bases_match(rc, q[c], bis) : (qk[c] & rk) != 0u;
          const double bump = same ? match : mism;
          const double s0 = diag + bump;
          diag = mu[c];
          // argmax with the traceback's priority 0 > 1 > 2 (1804-1811)
          double m = s0;
          int ak = 0;
          if (s1 > m) { m = s1; ak = 1; }
          if (s2 > m) { m = s2; ak = 2; }
          if (TRACE) {
            unsigned nib = (unsigned)ak | ((s1 - ge > s0 - go) ? 4u : 0u) | ((s2 - ge > s0 - go) ? 8u : 0u);
            dword |= (unsigned long long)nib << (4 * c);
          }
          if (!TRACE && c == cmm) {  // scan of the last column, states 0,1,2 in order with strict > (1724-1741)
            if (s0 > best) { best = s0; bk = 0; bi = i; }
            if (s1 > best) { best = s1; bk = 1; bi = i; }
            if (s2 > best) { best = s2; bk = 2; bi = i; }
          }
          s0u[c] = s0;
          s1u[c] = s1;
          mu[c] = m;
          l_s0 = s0;
          l_s2 = s2;
        }
        out_s0 = l_s0;
        out_s2 = l_s2;
        out_m = diag;  // M[i-1][last column of this lane]
        if (MODE == 1) dirs[(size_t)(i - 1) * G + gl] = dword;
        if (MODE == 2) {
          const int slot = gl - (band_center_lane<WD>(i, dend) - a.band_half);
          if (slot >= 0 && slot <= 2 * a.band_half) band[(i - 1) * PM_BAND_LANES + slot] = dword;
        }
      }
    }

    if (!TRACE) {
      if (cmm >= 0) {  // exactly one lane owns column mm; with nn <= 0 it still holds S[0][0][mm]
        res.score = best;
        res.maxi = bi;
        res.maxk = bk;
        a.results[task_id] = res;
      }
    } else {
      __syncwarp(gmask);
      // smith_waterman_backtrack (1752-1965) over the stored decisions
      if (MODE == 1) {
        if (gl == 0 && have && nn > 0) {
          FullCell<G, WD, 4> cell;
          cell.dirs = dirs;
          walk_path<true, 0>(cell, res.maxk, res.maxi, mm, read, mm, orient, tk.wstart, sink, s_q[grp], 0);
          atomicAdd(&a.counters->tb_cells, (unsigned long long)nn * (unsigned long long)mm);
        }
      } else if (have && nn > 0) {  // the whole group walks (coop_walk, trace_walk.cuh); the path segments go to the pend scratch
        BandCell<WD, 4> cell;
        cell.band = band;
        cell.dend = dend;
        cell.half = a.band_half;
        uint2* segs = reinterpret_cast<uint2*>(sink.pend);
        int nseg = 0;
        NoTie nt;
        int rc = coop_walk<G>(cell, nt, gmask, (tid & 31) / G * G, gl, res.maxk, res.maxi, mm, 0, segs, &nseg);
        if (rc == PM_WALK_OK && nseg < 0) rc = PM_WALK_OOB;  // more segments than the scratch holds
        if (rc == PM_WALK_OK) {
          coop_apply<G>(segs, nseg, gmask, gl, read, mm, orient, tk.wstart, sink, s_q[grp], 0);
          if (gl == 0) atomicAdd(&a.counters->tb_cells, (unsigned long long)nn * (unsigned long long)mm);
        } else if (gl == 0) {
          const uint32_t w = atomicAdd(a.oob_cursor, 1u);
          a.oob_winners[w] = a.winners[item];
        }
      }
      __syncwarp(gmask);
    }
  }
}

}  // namespace pm
