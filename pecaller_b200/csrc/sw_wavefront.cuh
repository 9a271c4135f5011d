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
// TRACE stores 4 decision bits per cell in a per-group scratch and lane 0 walks them,
// applying the pileup increments with atomics.
#pragma once
#include "pemap_common.cuh"

namespace pm {

struct SwArgs {
  const Task* tasks;
  TaskResult* results;         // score kernel: written; trace kernel: read (maxk/maxi of the winner)
  const Winner* winners;       // trace kernel: the winners; score kernel with list_mode: the tasks to (re)score
  int list_mode;
  const uint32_t* n_items;     // device counter: number of tasks (score) or winners (trace)
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  const double* border;        // S*[0][j], j < PM_DP_MAX (pemapper.c:2073-2081)
  uint32_t* counts;            // [genome_size][6] pileup counters
  unsigned long long* dirs;    // trace scratch: per group PM_DP_MAX * G words
  char* pend;                  // trace scratch: per group PM_DP_MAX bytes (pending insertion chars)
  unsigned char* ins_buf;      // insertion records: {u32 pos, u32 len, chars padded to 4}
  unsigned long long* ins_cursor;
  unsigned long long ins_cap;
  SeedCounters* counters;
  DevParams p;
};

__device__ __forceinline__ char seq_char(const char* read, int len, int orient, int j0) {  // j0 = 0-based read index
  return orient ? rt_char(read[len - 1 - j0]) : read[j0];
}

__device__ __forceinline__ void emit_insertion(const SwArgs& a, uint32_t site, const char* pend, int n) {
  unsigned long long need = 8ull + (unsigned long long)((n + 3) & ~3);
  unsigned long long off = atomicAdd(a.ins_cursor, need);
  if (off + need <= a.ins_cap) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(a.ins_buf + off);
    hdr[0] = site;
    hdr[1] = (uint32_t)n;
    unsigned char* dst = a.ins_buf + off + 8;
    for (int m = 0; m < n; m++) dst[m] = (unsigned char)pend[n - (m + 1)];  // 1892-1893: un-reverse
  }
  atomicAdd(&a.counts[(size_t)site * 6 + 5], 1u);  // no_ins++ (1903 / 1952)
}

template <int G, int WD, bool TRACE>
__global__ void __launch_bounds__(128) k_sw_fp64(SwArgs a) {
  constexpr int GROUPS_PER_BLOCK = 128 / G;
  __shared__ char s_win[GROUPS_PER_BLOCK][PM_DP_MAX];
  const int tid = threadIdx.x;
  const int grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) / G * G));
  const uint32_t n_items = *a.n_items;
  const uint32_t ggid = blockIdx.x * GROUPS_PER_BLOCK + grp, n_groups = gridDim.x * GROUPS_PER_BLOCK;
  const double go = a.p.go, ge = a.p.ge, match = a.p.match, mism = a.p.mism;
  char* win = s_win[grp];

  for (uint32_t item = ggid; item < n_items; item += n_groups) {
    const uint32_t task_id = (TRACE || a.list_mode) ? a.winners[item].task : item;
    const Task tk = a.tasks[task_id];
    const int orient = (int)(tk.rm >> 31);
    const uint32_t rm = tk.rm & 0x7FFFFFFFu;
    const int mate = (int)(rm & 1u);
    const uint32_t r = rm >> 1;
    const int mm = a.len[mate][r];
    const char* read = a.reads[mate] + (size_t)r * a.stride;
    const int nn = tk.blen;

    __syncwarp(gmask);
    for (int i = gl; i < nn; i += G) win[i] = a.genome[(size_t)tk.wstart + i];
    // this lane's read characters (reverse strand = reverse_transcribe'd read, 1021/1098)
    char q[WD];
    const int jbase = gl * WD;  // columns jbase+1 .. jbase+WD
#pragma unroll
    for (int c = 0; c < WD; c++) {
      int j0 = jbase + c;
      q[c] = (j0 < mm) ? seq_char(read, mm, orient, j0) : (char)0;
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
    unsigned long long* dirs = TRACE ? a.dirs + (size_t)ggid * PM_DP_MAX * G : nullptr;
    __syncwarp(gmask);

    const int steps = nn > 0 ? nn + G - 1 : 0;
    for (int s = 0; s < steps; s++) {
      double in_s0 = __shfl_up_sync(gmask, out_s0, 1, G);
      double in_s2 = __shfl_up_sync(gmask, out_s2, 1, G);
      double in_m = __shfl_up_sync(gmask, out_m, 1, G);
      if (gl == 0) {  // column 0: S0[i][0] = 0, S2[i][0] = -go, M[i-1][0] = max(0, 0, -go) = 0 (2062-2081)
        in_s0 = 0.0;
        in_s2 = -1.0 * go;
        in_m = 0.0;
      }
      const int i = s - gl + 1;
      if (i >= 1 && i <= nn) {
        const char rc = win[i - 1];
        double l_s0 = in_s0, l_s2 = in_s2, diag = in_m;
        unsigned long long dword = 0;
#pragma unroll
        for (int c = 0; c < WD; c++) {
          const double s2 = dmax(l_s0 - go, l_s2 - ge);
          const double s1 = dmax(s0u[c] - go, s1u[c] - ge);
          const double bump = bases_match(rc, q[c], a.p.is_bisulfite) ? match : mism;
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
          if (c == cmm) {  // scan of the last column, states 0,1,2 in order with strict > (1724-1741)
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
        if (TRACE) dirs[(size_t)(i - 1) * G + gl] = dword;
      }
    }

    if (!TRACE) {
      if (cmm >= 0) {  // exactly one lane owns column mm; with nn <= 0 it still holds S[0][0][mm]
        TaskResult res;
        res.score = best;
        res.maxi = bi;
        res.maxk = bk;
        a.results[task_id] = res;
      }
    } else {
      __syncwarp(gmask);
      if (gl == 0 && nn > 0) {
        // smith_waterman_backtrack (1752-1965) over the stored decisions
        const TaskResult res = a.results[task_id];
        char* pend = a.pend + (size_t)ggid * PM_DP_MAX;
        int n_pend = 0;
        int k = res.maxk, i = res.maxi, j = mm, i1 = 0, j1 = 0;
        while (i > 0 && j > 0) {
          i1 = i - 1;
          j1 = j - 1;
          int pi, pj, pk = 0;
          if (k == 0) {
            pi = i1; pj = j1;
            if (pi > 0 && pj > 0) pk = (int)((dirs[(size_t)(pi - 1) * G + (pj - 1) / WD] >> (4 * ((pj - 1) % WD))) & 3ull);
          } else if (k == 2) {
            pi = i; pj = j1;
            if (pj > 0) pk = ((dirs[(size_t)(pi - 1) * G + (pj - 1) / WD] >> (4 * ((pj - 1) % WD))) & 8ull) ? 2 : 0;
          } else {
            pi = i1; pj = j;
            if (pi > 0) pk = ((dirs[(size_t)(pi - 1) * G + (pj - 1) / WD] >> (4 * ((pj - 1) % WD))) & 4ull) ? 1 : 0;
          }
          const uint32_t site = tk.wstart + (uint32_t)i1;
          if (pi != i) {
            if (pj != j) {  // 1846-1858
              const char ch = seq_char(read, mm, orient, j1);
              const int col = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
              if (col >= 0) atomicAdd(&a.counts[(size_t)site * 6 + col], 1u);
            } else {
              atomicAdd(&a.counts[(size_t)site * 6 + 4], 1u);  // 1868
            }
            if (n_pend > 0) emit_insertion(a, site, pend, n_pend);  // 1871-1904
            n_pend = 0;
          } else {
            pend[n_pend++] = seq_char(read, mm, orient, j1);  // 1910-1911
          }
          i = pi; j = pj; k = pk;
        }
        if (n_pend > 0 && i >= 1) emit_insertion(a, tk.wstart + (uint32_t)i1, pend, n_pend);  // 1918-1958
        atomicAdd(&a.counters->tb_cells, (unsigned long long)nn * (unsigned long long)mm);
      }
      __syncwarp(gmask);
    }
  }
}

}  // namespace pm
