// sw_int16.cuh - integer Smith-Waterman scoring with packed s16x2 DPX instructions (sm_100a) and the integer
// version of the selection rules, with tie flags that send a read to the exact fp64 path.
//
// All of the reference's scores are multiples of 1/36 (match +1, mismatch -1/3, gap open 2, gap extend 1/36;
// pemapper.c:2011, 2039-2040), so the DP is rationally exact in integer units of 1/36: +36, -12, -72, -1.
// Two alignment tasks are packed per 32-bit lane (low / high 16 bits) and advance with VIADDMNMX.S16x2,
// VIMNMX3.S16x2 and VIMNMX.U16x2; values are stored biased by +1024 so that every half stays positive and the
// constant subtractions can be plain 32-bit adds without borrows between halves.
//
// k_sw_i16 keeps every state in the frame T[i][j] = S[i][j] + i + j: the gap-extension steps S - 1 one column or one
// row further then cost nothing (T2[i][j] = max(T0[i][j-1] - 71, T2[i][j-1]) is ONE VIADDMNMX with no separate
// subtraction, likewise T1), the diagonal step becomes T0 = TM[i-1][j-1] + 2 + bump = (TM - 10) + 48 * match, and all
// comparisons inside a cell are unchanged.  The three-input DPX instructions issue at half rate on sm_100a
// (profiles/README_r02.md), so per pair of cells the loop is 2 x 2 (VIADDMNMX) + 8 single-rate slots instead of + 10.
// Only the last-column scan compares cells of different rows: it subtracts the row index again.
//
// The reference compares IEEE doubles with strict '>' (1101, 1147, 1455, 1463, 1724-1741, 1805-1811), and
// rationally equal scores need not be equal doubles.  Every comparison whose outcome
// could depend on that is detected here (equal integers) and the read is replayed by the fp64 kernels; all
// other comparisons have operands at least 1/36 apart and resolve identically in double.
#pragma once
#include "pemap_common.cuh"
#include "select_finish.cuh"
#include "sw_wavefront.cuh"

#define PM_IBIAS 1024

namespace pm {

struct ITaskResult {
  int32_t score36;  // score in units of 1/36
  uint16_t maxi;    // start[1]
  uint8_t maxk;     // start[0]
  uint8_t flags;    // 1: the last-column argmax has a rational tie; 2: a base outside ACGTN (codes not valid);
                    // 4: the traceback from (0, maxi, mm) is a pure diagonal decided by strict integer inequalities
};

struct SwIntArgs {
  const Task* tasks;
  ITaskResult* results;
  const uint32_t* n_items;   // number of tasks to score (entries of `list` when it is set)
  const uint32_t* list;      // task ids left over by k_diag_certify, or nullptr: every task
  uint32_t* work;   // zeroed per launch: next pair to score
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  int lane_mm;  // CMM >= 0 only: lane that owns the last read column
  DevParams p;
};

// 4-bit one-hot codes: two bases match iff their codes share a bit (N = all bits).  0 = not representable.
__device__ __forceinline__ uint32_t ref_code(char ch, int bis) {
  switch (ch) {
    case 'A': return 1u;
    case 'C': return bis ? 10u : 2u;  // bisulfite: reference C also matches read T (2026-2033)
    case 'G': return 4u;
    case 'T': return 8u;
    case 'N': case 'n': return 15u;
    default: return 0u;
  }
}
__device__ __forceinline__ uint32_t read_code(char ch) {
  switch (ch) {
    case 'A': return 1u;
    case 'C': return 2u;
    case 'G': return 4u;
    case 'T': return 8u;
    case 'N': case 'n': return 15u;
    default: return 0u;
  }
}

__device__ __forceinline__ void track_best(int v0, int v1, int v2, int dq, int i, int& best, int& bk, int& bi, int& tie,
                                           int& bdq) {
  // scan of the last column, states 0,1,2 in order with strict '>' (1724-1741); equal values are remembered
  if (v0 > best) { best = v0; bk = 0; bi = i; tie = 0; bdq = dq; } else if (v0 == best) tie = 1;
  if (v1 > best) { best = v1; bk = 1; bi = i; tie = 0; } else if (v1 == best) tie = 1;
  if (v2 > best) { best = v2; bk = 2; bi = i; tie = 0; } else if (v2 == best) tie = 1;
}

// CMM >= 0: every task of the launch has (len-1) % WD == CMM and (len-1) / WD == lane_mm (uniform read length);
// CMM < 0: per-task lengths.
// ROWCAP > 0: windows of up to ROWCAP rows (BASELINE configs[3] scores reads against 1000-bp windows; the mapper's
// own windows are len + 21 rows).  The rows stream through the wavefront, so only the staged window grows.
template <int G, int WD, int CMM, int ROWCAP = 0>
__global__ void __launch_bounds__(128, 6) k_sw_i16(SwIntArgs a) {
  constexpr int GPB = 128 / G;
  constexpr int ROWS = ROWCAP > 0 ? ROWCAP : ((G * WD + 24 < PM_DP_MAX) ? G * WD + 24 : PM_DP_MAX);  // window rows: nn <= len + 21
  // uniform read length: the last read column of every row is parked and scanned after the sweep; with long windows
  // (ROWCAP) that buffer would not fit, the owner lane then tracks the maximum row by row at its compile-time column
  constexpr bool PARK = CMM >= 0 && ROWCAP == 0;
  __shared__ uint32_t s_win[GPB][ROWS];
  __shared__ __align__(16) uint32_t s_last[PARK ? GPB : 1][PARK ? 4 * ROWS : 4];  // last read column of every row
  const int tid = threadIdx.x, grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) / G * G));
  const uint32_t n_items = *a.n_items, n_pairs = (n_items + 1) >> 1;
  uint32_t* win = s_win[grp];
  uint32_t* last = s_last[PARK ? grp : 0];
  constexpr uint32_t K1 = 0x00010001u, K10 = 0x000A000Au, NEG71 = 0xFFB9FFB9u;
  constexpr uint32_t BIASP = (PM_IBIAS << 16) | PM_IBIAS;

  for (;;) {
    const uint32_t pair = next_work_item<G>(a.work, gmask, gl);
    if (pair >= n_pairs) break;
    uint32_t idA = 2 * pair, idB = (2 * pair + 1 < n_items) ? 2 * pair + 1 : 2 * pair;
    if (a.list) {
      idA = a.list[idA];
      idB = a.list[idB];
    }
    const Task tA = a.tasks[idA], tB = a.tasks[idB];
    const int orA = (int)(tA.rm >> 31), orB = (int)(tB.rm >> 31);
    const uint32_t rmA = tA.rm & 0x7FFFFFFFu, rmB = tB.rm & 0x7FFFFFFFu;
    const int mmA = a.len[rmA & 1][rmA >> 1], mmB = a.len[rmB & 1][rmB >> 1];
    const char* readA = a.reads[rmA & 1] + (size_t)(rmA >> 1) * a.stride;
    const char* readB = a.reads[rmB & 1] + (size_t)(rmB >> 1) * a.stride;
    const int nnA = tA.blen > 0 ? tA.blen : 0, nnB = tB.blen > 0 ? tB.blen : 0;
    const int nn = nnA > nnB ? nnA : nnB;

    __syncwarp(gmask);
    bool badA = false, badB = false;
    for (int i = gl; i < nn; i += G) {
      uint32_t cA = 0, cB = 0;
      if (i < nnA) { cA = ref_code(a.genome[(size_t)tA.wstart + i], a.p.is_bisulfite); badA |= (cA == 0); }
      if (i < nnB) { cB = ref_code(a.genome[(size_t)tB.wstart + i], a.p.is_bisulfite); badB |= (cB == 0); }
      win[i] = cA | (cB << 16);
    }
    // dvu: "diagonal quality" E[i][j] = min over the cells (i-t, j-t), t >= 0, of (S0 - max(S1, S2)) clipped at 0:
    // >= 1 iff the traceback arriving there in state 0 keeps taking state 0 by strict inequalities (1804-1811)
    uint32_t q[WD], s0u[WD], s1u[WD], mu[WD], dvu[WD];
    const int jbase = gl * WD;
#pragma unroll
    for (int c = 0; c < WD; c++) {
      const int j0 = jbase + c;
      uint32_t cA = 0, cB = 0;
      if (j0 < mmA) { cA = read_code(seq_char(readA, mmA, orA, j0)); badA |= (cA == 0); }
      if (j0 < mmB) { cB = read_code(seq_char(readB, mmB, orB, j0)); badB |= (cB == 0); }
      q[c] = cA | (cB << 16);
      // row 0: S*[0][j] = -(72 + j - 1) (2073-2081), i.e. T = -71 in every column; mu holds max(T0,T1,T2) - 10
      const uint32_t b = (uint32_t)(PM_IBIAS - 71);
      s0u[c] = b | (b << 16);
      s1u[c] = s0u[c];
      mu[c] = s0u[c] - K10;
      dvu[c] = 0xFFFFFFFFu;  // row 0 is never consulted by the walk (loop ends at i == 0)
    }
    badA = __any_sync(gmask, badA);
    badB = __any_sync(gmask, badB);

    int colA, colB;
    bool ownA, ownB;
    if (CMM >= 0) {
      colA = colB = CMM;
      ownA = ownB = (gl == a.lane_mm);
    } else {
      ownA = ((mmA - 1) / WD == gl);
      ownB = ((mmB - 1) / WD == gl);
      colA = ownA ? (mmA - 1) % WD : -1;
      colB = ownB ? (mmB - 1) % WD : -1;
    }
    // the running maximum of the last column is kept as S + mm (+ bias): T minus the row index; S[0][0][mm] (1701-1703)
    int bestA = PM_IBIAS - 71, bkA = 0, biA = 0, tieA = 0, dqA = 0;
    int bestB = PM_IBIAS - 71, bkB = 0, biB = 0, tieB = 0, dqB = 0;
    uint32_t out_s0 = 0, out_s2 = 0, out_m = 0, out_dv = 0xFFFFFFFFu;
    __syncwarp(gmask);

    const int steps = nn > 0 ? nn + G - 1 : 0;
    for (int s = 0; s < steps; s++) {
      uint32_t l_s0 = __shfl_up_sync(gmask, out_s0, 1, G);
      uint32_t l_s2 = __shfl_up_sync(gmask, out_s2, 1, G);
      uint32_t diag = __shfl_up_sync(gmask, out_m, 1, G);
      uint32_t dvd = __shfl_up_sync(gmask, out_dv, 1, G);
      const int i = s - gl + 1;
      if (gl == 0) {  // column 0 of row i: S0 = 0, S2 = -72, M(row above) = 0 (2062-2081) -> T0 = i, T2 = i - 72, TM - 10 = i - 11
        const uint32_t ik = BIASP + (uint32_t)i * K1;
        l_s0 = ik;
        l_s2 = ik - 0x00480048u;
        diag = ik - 0x000B000Bu;
        dvd = 0xFFFFFFFFu;  // the walk stops at j == 0
      }
      if (i >= 1 && i <= nn) {
        const uint32_t rc = win[i - 1];
#pragma unroll
        for (int c = 0; c < WD; c++) {
          const uint32_t s2 = __viaddmax_s16x2(l_s0, NEG71, l_s2);           // 1710 / 1720
          const uint32_t s1 = __viaddmax_s16x2(s0u[c], NEG71, s1u[c]);       // 1711 / 1721
          const uint32_t m01 = __vminu2(q[c] & rc, K1);                      // 1 where the bases match
          const uint32_t s0 = m01 * 48u + diag;                              // (TM - 10) + 48 or + 0 (1713 / 1723)
          diag = mu[c];
          const uint32_t m12 = __vmaxs2(s1, s2);
          const uint32_t m = __vmaxs2(s0, m12);
          if (PARK) {
            if (c == CMM && ownA) {  // uniform read length: park the last column, scan it after the sweep
              *reinterpret_cast<uint4*>(&last[4 * (i - 1)]) = make_uint4(s0, s1, s2, dvd);
            }
          } else if (CMM >= 0) {
            if (c == CMM && ownA) {
              if (i <= nnA)
                track_best((int)(s0 & 0xFFFFu) - i, (int)(s1 & 0xFFFFu) - i, (int)(s2 & 0xFFFFu) - i, (int)(dvd & 0xFFFFu), i,
                           bestA, bkA, biA, tieA, dqA);
              if (i <= nnB)
                track_best((int)(s0 >> 16) - i, (int)(s1 >> 16) - i, (int)(s2 >> 16) - i, (int)(dvd >> 16), i, bestB, bkB, biB,
                           tieB, dqB);
            }
          } else {
            if (c == colA && ownA && i <= nnA)
              track_best((int)(s0 & 0xFFFFu) - i, (int)(s1 & 0xFFFFu) - i, (int)(s2 & 0xFFFFu) - i, (int)(dvd & 0xFFFFu), i,
                         bestA, bkA, biA, tieA, dqA);
            if (c == colB && ownB && i <= nnB)
              track_best((int)(s0 >> 16) - i, (int)(s1 >> 16) - i, (int)(s2 >> 16) - i, (int)(dvd >> 16), i, bestB, bkB, biB,
                         tieB, dqB);
          }
          // both halves of m - m12 are >= 0, so the plain subtraction does not borrow across halves
          const uint32_t e = __vminu2(dvd, m - m12);
          dvd = dvu[c];
          dvu[c] = e;
          s0u[c] = s0;
          s1u[c] = s1;
          mu[c] = m - K10;
          l_s0 = s0;
          l_s2 = s2;
        }
        out_s0 = l_s0;
        out_s2 = l_s2;
        out_m = diag;
        out_dv = dvd;
      }
    }
    if (PARK) {
      // scan of the last column (1717-1742) by the whole group.  key = value<<10 | (1023 - order) keeps the FIRST
      // row/state of the maximum (strict '>' in scan order); the mirrored key keeps the LAST one; they differ
      // exactly when the maximum occurs more than once.  order = 3*i + k <= 3*320+2 < 1024.
      __syncwarp(gmask);
      uint32_t fA = 0, lA = 0, fB = 0, lB = 0;
      if (gl == 0) {  // S[0][0][mm]: i = 0, k = 0
        fA = ((uint32_t)bestA << 10) | 1023u;
        lA = ((uint32_t)bestA << 10);
        fB = ((uint32_t)bestB << 10) | 1023u;
        lB = ((uint32_t)bestB << 10);
      }
      for (int r = gl; r < nn; r += G) {
        const uint32_t w0 = last[4 * r], w1 = last[4 * r + 1], w2 = last[4 * r + 2];
        const uint32_t o = 3u * (uint32_t)(r + 1), ri = (uint32_t)(r + 1);  // parked values are T: take the row index off
        if (r < nnA) {
          const uint32_t v0 = ((w0 & 0xFFFFu) - ri) << 10, v1 = ((w1 & 0xFFFFu) - ri) << 10, v2 = ((w2 & 0xFFFFu) - ri) << 10;
          fA = max(fA, max(v0 | (1023u - o), max(v1 | (1022u - o), v2 | (1021u - o))));
          lA = max(lA, max(v0 | o, max(v1 | (o + 1u), v2 | (o + 2u))));
        }
        if (r < nnB) {
          const uint32_t v0 = ((w0 >> 16) - ri) << 10, v1 = ((w1 >> 16) - ri) << 10, v2 = ((w2 >> 16) - ri) << 10;
          fB = max(fB, max(v0 | (1023u - o), max(v1 | (1022u - o), v2 | (1021u - o))));
          lB = max(lB, max(v0 | o, max(v1 | (o + 1u), v2 | (o + 2u))));
        }
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) {
        fA = max(fA, __shfl_xor_sync(gmask, fA, off, G));
        lA = max(lA, __shfl_xor_sync(gmask, lA, off, G));
        fB = max(fB, __shfl_xor_sync(gmask, fB, off, G));
        lB = max(lB, __shfl_xor_sync(gmask, lB, off, G));
      }
      const uint32_t oA = 1023u - (fA & 1023u), oB = 1023u - (fB & 1023u);
      bestA = (int)(fA >> 10); biA = (int)(oA / 3u); bkA = (int)(oA % 3u); tieA = (lA & 1023u) != oA;
      bestB = (int)(fB >> 10); biB = (int)(oB / 3u); bkB = (int)(oB % 3u); tieB = (lB & 1023u) != oB;
      dqA = biA > 0 ? (int)(last[4 * (biA - 1) + 3] & 0xFFFFu) : 0;
      dqB = biB > 0 ? (int)(last[4 * (biB - 1) + 3] >> 16) : 0;
    }
    if (ownA) {
      ITaskResult r;
      r.score36 = bestA - PM_IBIAS - mmA;
      r.maxi = (uint16_t)biA;
      r.maxk = (uint8_t)bkA;
      r.flags = (uint8_t)((tieA ? 1 : 0) | (badA ? 2 : 0) | ((bkA == 0 && biA > 0 && dqA >= 1) ? 4 : 0));
      a.results[idA] = r;
    }
    if (ownB && idB != idA) {
      ITaskResult r;
      r.score36 = bestB - PM_IBIAS - mmB;
      r.maxi = (uint16_t)biB;
      r.maxk = (uint8_t)bkB;
      r.flags = (uint8_t)((tieB ? 1 : 0) | (badB ? 2 : 0) | ((bkB == 0 && biB > 0 && dqB >= 1) ? 4 : 0));
      a.results[idB] = r;
    }
  }
}

// (read i, window i) -> Task i, and back: the microbenchmark entry pemap_sw_score_device
__global__ void __launch_bounds__(256) k_sw_bench_tasks(int n, const uint32_t* win_start, const int* win_len, Task* tasks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Task t;
  t.rm = 2u * (uint32_t)i;  // mate 0, forward orientation
  t.spot = win_start[i];
  t.wstart = win_start[i];
  t.blen = win_len[i];
  tasks[i] = t;
}
__global__ void __launch_bounds__(256) k_sw_bench_results(int n, const ITaskResult* res, int32_t* score36, int32_t* maxi, int32_t* maxk,
                                                          int32_t* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ITaskResult r = res[i];
  score36[i] = r.score36;
  maxi[i] = r.maxi;
  maxk[i] = r.maxk;
  if (flags) flags[i] = r.flags;
}

// ---------------------------------------------------------------------------------------------------
// k_diag_certify: candidates whose result follows from their ungapped diagonals alone skip the DP.
//
// In units of 1/36 every DP value of column j is either the sum of one whole diagonal started at column 0
// (S0[i-j][0] = 0: the free start of the semi-global alignment) or the value of a path with a gap or a row-0 border
// start, and those are <= 36 j - 72 (induction over 1710-1713 with the borders 2062-2081: every read base adds at
// most +36 and the first gap or border costs at least 72).  A diagonal with at most one mismatch scores
// >= 36 j - 48 at every column, i.e. 24 above anything with a gap.  So if, among the K + 1 = nn - mm + 1 whole
// diagonals of the window, EXACTLY ONE has <= 2 mismatches (say offset o, m mismatches; for m = 2 with the extra
// condition spelled out in the kernel) then, with no rounding involved anywhere:
//   * on that diagonal S0[j+o][j] is its prefix sum and S1, S2 are at least 24 below it: the traceback from its
//     last cell is the pure state-0 diagonal decided by strict inequalities (the `flags & 4` of k_sw_i16);
//   * the last-column scan (1717-1742) has a unique maximum 36 mm - 48 m at (k, i) = (0, o + mm): every other
//     whole diagonal has >= 3 mismatches (<= 36 mm - 144) and everything else is below it too.
// That is the ITaskResult k_sw_i16 would write.  Anything else (no such diagonal, two of them, a character outside
// ACGT, a window shorter than the read) goes on the list of tasks k_sw_i16 scores.
// One warp per task, 32 consecutive tasks per warp: lane o tries the first 8 columns of diagonal o, the surviving
// diagonals are counted in full by all lanes; the left-over tasks of a warp cost one atomic.
// ---------------------------------------------------------------------------------------------------

struct CertifyArgs {
  const Task* tasks;
  ITaskResult* results;
  const uint32_t* n_items;
  uint32_t* list;          // tasks that need the DP
  uint32_t* list_cursor;
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  unsigned long long* cells_certified;  // sum of nn * mm over the certified tasks
  DevParams p;
};

__device__ __forceinline__ bool is_acgt(unsigned ch) {
  return (ch & 0xE0u) == 0x40u && ((0x0010008Au >> (ch & 31u)) & 1u);  // 'A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54
}

// 8 consecutive bytes starting at byte `off` of a word array (little endian: byte 0 of .x is the first)
__device__ __forceinline__ uint2 fetch8(const uint32_t* p, int off) {
  const int wi = off >> 2, sh = (off & 3) * 8;
  const uint32_t a = p[wi], b = p[wi + 1], c = p[wi + 2];
  return make_uint2(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh));
}
// bit 7 of every byte of x that is not zero
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {
  return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

template <int WARPS, bool BIS>
__global__ void __launch_bounds__(WARPS * 32) k_diag_certify(CertifyArgs a) {
  __shared__ uint32_t s_w[WARPS][(PM_DP_MAX + 32) / 4 + 2];
  __shared__ uint32_t s_r[WARPS][PM_DP_MAX / 4 + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_items = *a.n_items;
  const uint32_t gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  unsigned char* w = reinterpret_cast<unsigned char*>(s_w[warp]);
  unsigned char* q = reinterpret_cast<unsigned char*>(s_r[warp]);
  constexpr bool bis = BIS;
  unsigned long long cells = 0;

  for (uint32_t t0 = gw * 32u; t0 < n_items; t0 += nw * 32u) {
    bool mine_needs_dp = false;   // lane t: does task t0 + t need the DP ?
    const uint32_t t_end = (n_items - t0 < 32u) ? n_items - t0 : 32u;
    for (uint32_t t = 0; t < t_end; t++) {
      const uint32_t id = t0 + t;
      const Task tk = a.tasks[id];
      const int orient = (int)(tk.rm >> 31);
      const uint32_t rm = tk.rm & 0x7FFFFFFFu;
      const int mm = a.len[rm & 1][rm >> 1];
      const char* read = a.reads[rm & 1] + (size_t)(rm >> 1) * a.stride;
      const int nn = tk.blen;
      const int K = nn - mm;                       // whole diagonals: offsets 0 .. K
      bool ok = K >= 0 && K < 32 && mm >= 16 && nn <= PM_DP_MAX;
      __syncwarp();
      if (ok) {
        bool good = true;
        for (int i = lane; i < nn; i += 32) {
          const unsigned ch = (unsigned char)a.genome[(size_t)tk.wstart + i];
          good &= is_acgt(ch);
          w[i] = (unsigned char)ch;
        }
        for (int j = lane; j < mm; j += 32) {      // the oriented read (reverse strand = reverse_transcribe, 2303-2337)
          unsigned ch = (unsigned char)read[orient ? mm - 1 - j : j];
          good &= is_acgt(ch);
          if (orient) ch ^= (ch & 2u) ? 0x04u : 0x15u;   // A <-> T, C <-> G
          q[j] = (unsigned char)ch;
        }
        ok = __all_sync(0xFFFFFFFFu, good);
      }
      __syncwarp();
      int o_star = -1, m_star = 0;
      if (ok) {
        // first and last 8 columns of diagonal `lane`: mismatches, leading and trailing runs of matches
        int mis = 3, pre = 0, suf = 0;
        if (lane <= K) {
          if (BIS) {
            mis = 0;
            bool run = true, runs = true;
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const unsigned rc = w[lane + j], qc = q[j];
              const bool mt = rc == qc || (rc == 'C' && qc == 'T');
              mis += !mt;
              run = run && mt;
              pre += run;
              const unsigned rs = w[lane + mm - 1 - j], qs = q[mm - 1 - j];
              runs = runs && (rs == qs || (rs == 'C' && qs == 'T'));
              suf += runs;
            }
          } else {  // 8 characters at a time: xor, per-byte non-zero flags
            const uint2 wf = fetch8(s_w[warp], lane), qf = fetch8(s_r[warp], 0);
            const uint2 wl = fetch8(s_w[warp], lane + mm - 8), ql = fetch8(s_r[warp], mm - 8);
            const uint32_t f0 = nonzero_bytes(wf.x ^ qf.x), f1 = nonzero_bytes(wf.y ^ qf.y);
            const uint32_t l0 = nonzero_bytes(wl.x ^ ql.x), l1 = nonzero_bytes(wl.y ^ ql.y);
            mis = __popc(f0) + __popc(f1);
            pre = f0 ? ((__ffs((int)f0) - 1) >> 3) : f1 ? 4 + ((__ffs((int)f1) - 1) >> 3) : 8;
            suf = l1 ? (__clz((int)l1) >> 3) : l0 ? 4 + (__clz((int)l0) >> 3) : 8;
          }
        }
        unsigned cand = __ballot_sync(0xFFFFFFFFu, mis < 3);
        int n_good = 0, p1 = 0;
        while (cand) {                              // usually one survivor: count it in full
          const int o = __ffs((int)cand) - 1;
          cand &= cand - 1u;
          int m = 0, first = mm + 1;                // first = 1-based column of the first mismatch
          for (int j = lane; j < mm; j += 32) {
            const unsigned rc = w[o + j], qc = q[j];
            if (!(rc == qc || (bis && rc == 'C' && qc == 'T'))) {
              m++;
              first = min(first, j + 1);
            }
          }
          m = __reduce_add_sync(0xFFFFFFFFu, m);
          if (m < 3) {
            n_good++;
            o_star = o;
            m_star = m;
            p1 = __reduce_min_sync(0xFFFFFFFFu, first);
          }
        }
        ok = n_good == 1;
        if (ok && m_star == 2) {
          // two mismatches score 36 mm - 96: a path with one gap and no mismatch (<= 36 * matched - 72) could reach that.
          // Such a path is a run of matches from column 1 on one diagonal and a run of matches up to its last column on
          // another; with A / B the longest leading / trailing runs of the OTHER diagonals it cannot reach any cell of
          // this diagonal or the last column with as much when A <= p1, B <= mm - p1 and A + B <= mm - 1 (p1 = column
          // of the first mismatch); two gaps, or a gap and a mismatch, cost >= 120.  tools/certify_bruteforce.py checks
          // both rules against the full DP.
          const bool other = lane <= K && lane != o_star;
          const int A = __reduce_max_sync(0xFFFFFFFFu, other ? pre : 0), B = __reduce_max_sync(0xFFFFFFFFu, other ? suf : 0);
          ok = A < 8 && B < 8 && A <= p1 && B <= mm - p1 && A + B <= mm - 1;   // a run of 8 may be longer: leave it to the DP
        }
      }
      if (ok) {
        if (lane == 0) {
          ITaskResult r;
          r.score36 = 36 * mm - 48 * m_star;
          r.maxi = (uint16_t)(o_star + mm);
          r.maxk = 0;
          r.flags = 4;
          a.results[id] = r;
          cells += (unsigned long long)nn * (unsigned long long)mm;
        }
      } else if (lane == (int)t) {
        mine_needs_dp = true;
      }
    }
    const unsigned need = __ballot_sync(0xFFFFFFFFu, mine_needs_dp);
    if (need) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.list_cursor, (uint32_t)__popc(need));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (mine_needs_dp) a.list[base + __popc(need & ((1u << lane) - 1u))] = t0 + (uint32_t)lane;
    }
  }
  if (lane == 0 && cells) atomicAdd(a.cells_certified, cells);
}

// ---------------------------------------------------------------------------------------------------
// Integer selection: same control flow as select_finish.cuh (1084-1192, 1313-1536) on exact integers.
// `amb` is raised whenever the double-precision outcome is not implied by the integers.
// ---------------------------------------------------------------------------------------------------

struct SelectIntArgs {
  const Task* tasks;
  const ITaskResult* ires;
  TaskResult* results64;      // maxi / maxk of the winners are copied here for the traceback kernel
  const uint32_t* cand_base;
  const uint32_t* cand_n;
  const int* len[2];
  int n_reads;
  uint32_t* m1;
  uint32_t* m2;
  int* mapping_type;
  Winner* winners;            // winners that need the full traceback (gaps, or a rational tie on the way)
  uint32_t* winner_cursor;
  Winner* diag_winners;       // winners whose traceback is a pure diagonal (ITaskResult.flags & 4)
  uint32_t* diag_cursor;
  uint32_t* replay_reads;     // reads whose outcome needs the fp64 path
  uint32_t* replay_read_cursor;
  Winner* replay_tasks;       // every task of those reads
  uint32_t* replay_task_cursor;
  SeedCounters* counters;
  const char* reads[2];       // for diag_score64
  int stride;
  const char* genome;
  const double* border;
  DevParams p;
};

// The reference's double for a task whose integer result has a unique last-column maximum in state 0 (flags & 1 clear),
// only ACGTN bases (flags & 2 clear) and, behind that cell, a state-0 diagonal on which S0 beats S1 and S2 by at least
// 1/36 at every cell (flags & 4).  Rounding errors are ~1e-13, so the double DP takes the same maxima:
// S0[i][j] = S0[i-1][j-1] + bump all the way (1713 / 1723), started from M[i0][0] = 0 (2062-2081) or, when the diagonal
// leaves through row 0, from the border S[0][j0] (2073-2081), and the last-column scan (1717-1742) returns that cell.
// The sum below is therefore the bit pattern k_sw_fp64 would produce, at mm additions instead of mm * nn cells.
__device__ __forceinline__ double diag_score64(const Task& tk, int maxi, int mm, const char* read, const char* genome,
                                               const double* border, const DevParams& p) {
  const int orient = (int)(tk.rm >> 31);
  const int j0 = maxi >= mm ? 0 : mm - maxi;   // the diagonal's cell before its first step: (i0, j0)
  const int i0 = maxi - (mm - j0);
  double s = j0 == 0 ? 0.0 : border[j0];
  for (int j = j0 + 1; j <= mm; j++) {
    const char rc = genome[(size_t)tk.wstart + (size_t)(i0 + (j - j0) - 1)];
    const char q = seq_char(read, mm, orient, j - 1);
    s = s + (bases_match(rc, q, p.is_bisulfite) ? p.match : p.mism);
  }
  return s;
}

// s >= good_score ?  returns 1 / 0, or -1 when the rational score is (numerically) on the threshold
__device__ __forceinline__ int ge_good(int s36, double good36) {
  const double d = (double)s36 - good36;
  if (d > 1e-6) return 1;
  if (d < -1e-6) return 0;
  return -1;
}

// Single-end rule (1101-1127) on integers.  Scores below the integer maximum differ from it by >= 1/36 as doubles
// too, so they can neither end up as `top` nor be counted as ties of it: the double-precision outcome is a function
// of the candidates AT the maximum alone (their order and their rounded scores).  Hence
//   maximum clearly below the threshold            -> NEITHER_MAP, no double needed
//   maximum clearly above it, one candidate has it -> UNIQUE_SINGLE (fp64 only if its own last-column argmax tied)
//   several candidates at the maximum, or the maximum on the threshold -> *amb, and only those candidates are
//   re-scored in fp64 (*narrow = their score); a base outside ACGTN anywhere -> *amb with every candidate re-scored.
__device__ __forceinline__ int single_rule_int(const ITaskResult* res, int n, int len, const DevParams& p, int* best,
                                               bool* amb, int* narrow) {
  const double good36 = (double)len * p.min_align * p.match_bonus * 36.0;
  int smax = -0x7FFFFFFF, nmax = 0, first = -1;
  bool odd = false;
  for (int q = 0; q < n; q++) {
    const int s = res[q].score36;
    odd |= (res[q].flags & 2) != 0;
    if (s > smax) { smax = s; nmax = 1; first = q; }
    else if (s == smax) nmax++;
  }
  *best = -1;
  *narrow = smax;
  if (odd) { *amb = true; *narrow = -0x7FFFFFFF; return T_NEITHER_MAP; }
  const int ge = ge_good(smax, good36);
  if (ge == 0) return T_NEITHER_MAP;
  if (ge < 0 || nmax > 1) { *amb = true; return T_NEITHER_MAP; }
  *best = first;
  if (res[first].flags & 1) *amb = true;
  return T_UNIQUE_SINGLE;
}

__device__ __forceinline__ int pair_rule_int(const Task* ta, const ITaskResult* ra, int n1, int l1, const Task* tb,
                                             const ITaskResult* rb, int n2, int l3, const DevParams& p, int* keep1,
                                             int* keep2, bool* amb) {
  const double good1 = (double)l1 * p.min_align * p.match_bonus * 36.0, good2 = (double)l3 * p.min_align * p.match_bonus * 36.0;
  for (int i = 0; i < n1; i++)
    if ((ra[i].flags & 2) || ge_good(ra[i].score36, good1) < 0) *amb = true;
  for (int i = 0; i < n2; i++)
    if ((rb[i].flags & 2) || ge_good(rb[i].score36, good2) < 0) *amb = true;
  long long tot_best = -100000LL * 36;
  int perfect = 0, slip = 0, sm1 = -1, sm2 = -1;
  *keep1 = *keep2 = -1;
  for (int w1 = 0; w1 < n1; w1++) {
    if (ge_good(ra[w1].score36, good1) <= 0) continue;
    const long long p1 = (long long)ta[w1].spot;
    const int or1 = (int)(ta[w1].rm >> 31);
    for (int w2 = 0; w2 < n2; w2++) {
      if (ge_good(rb[w2].score36, good2) <= 0) continue;
      long long dist = p1 - (long long)tb[w2].spot;
      if (dist < 0) dist = -dist;
      if (!(dist >= p.min_dist && dist <= p.max_dist && or1 != (int)(tb[w2].rm >> 31))) continue;
      const long long sum = (long long)ra[w1].score36 + rb[w2].score36;
      if (sum > tot_best) {  // inc > 0.001: sums differ by at least 1/36
        perfect = 1;
        sm1 = w1;
        sm2 = w2;
        tot_best = sum;
        slip = 1;
      } else if (sum == tot_best) {  // inc > -0.001: equal sums differ by rounding only
        if (sm1 == w1 || sm2 == w2) slip++;
        perfect++;
      }
    }
  }
  if (perfect > 0) {
    if (perfect == 1 || slip == perfect) {
      *keep1 = sm1;
      *keep2 = sm2;
      if ((ra[sm1].flags | rb[sm2].flags) & 1) *amb = true;
      return perfect == 1 ? T_UNIQUE_MATE : T_UNIQUE_SLIP;
    }
    return T_NON_MATE;
  }
  int best1 = 0, best2 = 0, c1 = 0, c2 = 0;
  bool tie1 = false, tie2 = false;
  for (int i = 1; i < n1; i++) {
    if (ra[i].score36 > ra[best1].score36) { best1 = i; c1 = 1; tie1 = false; }
    else if (ra[i].score36 == ra[best1].score36) { c1++; tie1 = true; }  // '>' on doubles is undecided here
  }
  for (int i = 1; i < n2; i++) {
    if (rb[i].score36 > rb[best2].score36) { best2 = i; c2 = 1; tie2 = false; }
    else {
      if (rb[i].score36 == rb[best2].score36) tie2 = true;
      const int other = best1 < n2 ? rb[best1].score36 : -36;  // smax2[best1] (sic, 1468); unset entries are -1.0
      if (rb[i].score36 >= other) c2++;
    }
  }
  if (tie1 || tie2) *amb = true;
  const bool ok2 = ge_good(rb[best2].score36, good2) > 0 && c2 < 2;
  if (ge_good(ra[best1].score36, good1) > 0 && c1 < 2) {
    *keep1 = best1;
    if (ra[best1].flags & 1) *amb = true;
    if (ok2) {
      *keep2 = best2;
      if (rb[best2].flags & 1) *amb = true;
      return T_UNIQUE_MIS;
    }
    return T_UNIQUE_SINGLE;
  }
  if (ok2) {
    *keep2 = best2;
    if (rb[best2].flags & 1) *amb = true;
    return T_UNIQUE_SINGLE;
  }
  return T_NON_MIS;
}

__global__ void __launch_bounds__(128) k_select_int(SelectIntArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_reads) return;
  const int n1 = (int)a.cand_n[2 * r], n2 = a.p.pair_flag ? (int)a.cand_n[2 * r + 1] : 0;
  const uint32_t b1 = a.cand_base[2 * r], b2 = a.p.pair_flag ? a.cand_base[2 * r + 1] : 0;
  const int l1 = a.len[0][r], l3 = a.p.pair_flag ? a.len[1][r] : 0;
  int keep1 = -1, keep2 = -1, call;
  bool amb = false;
  int narrow = -0x7FFFFFFF;  // single-end rule: only candidates with this integer score need their double
  if (n1 > 0 && n2 == 0) call = single_rule_int(a.ires + b1, n1, l1, a.p, &keep1, &amb, &narrow);
  else if (n2 > 0 && n1 == 0) call = single_rule_int(a.ires + b2, n2, l3, a.p, &keep2, &amb, &narrow);
  else if (n1 > 0 && n2 > 0)
    call = pair_rule_int(a.tasks + b1, a.ires + b1, n1, l1, a.tasks + b2, a.ires + b2, n2, l3, a.p, &keep1, &keep2, &amb);
  else call = T_NEITHER_MAP;
  // a task of a replayed read whose double follows from its diagonal alone (diag_score64) is settled here
  auto settle = [&](const uint32_t id, const uint32_t rmv) {
    const ITaskResult ir = a.ires[id];
    if ((ir.flags & 7) != 4) return false;
    const int mm = a.len[rmv & 1][rmv >> 1];
    TaskResult t64;
    t64.score = diag_score64(a.tasks[id], (int)ir.maxi, mm, a.reads[rmv & 1] + (size_t)(rmv >> 1) * a.stride, a.genome, a.border, a.p);
    t64.maxi = ir.maxi;
    t64.maxk = ir.maxk;
    a.results64[id] = t64;
    return true;
  };
  if (amb && narrow != -0x7FFFFFFF) {
    // single-end rule with a rational tie at the top: the exact selection (k_select) needs the reference's doubles of
    // the tied candidates only; every other candidate gets its rational score, which orders it exactly as its double
    // would (>= 1/36 below the top) and is never consulted for anything else
    const uint32_t slot = atomicAdd(a.replay_read_cursor, 1u);
    a.replay_reads[slot] = (uint32_t)r;
    const int n = n1 > 0 ? n1 : n2;
    const uint32_t b = n1 > 0 ? b1 : b2, rmv = n1 > 0 ? 2u * r : 2u * r + 1u;
    int queued = 0;
    for (int q = 0; q < n; q++) {
      const ITaskResult ir = a.ires[b + q];
      queued += ir.score36 == narrow && (ir.flags & 7) != 4;
    }
    uint32_t at = queued ? atomicAdd(a.replay_task_cursor, (uint32_t)queued) : 0u;
    for (int q = 0; q < n; q++) {
      const ITaskResult ir = a.ires[b + q];
      if (ir.score36 == narrow) {
        if (!settle(b + q, rmv)) {
          a.replay_tasks[at].task = b + q;
          a.replay_tasks[at].rm = rmv;
          at++;
        }
      } else {
        TaskResult t64;
        t64.score = (double)ir.score36 / 36.0;
        t64.maxi = ir.maxi;
        t64.maxk = ir.maxk;
        a.results64[b + q] = t64;
      }
    }
    atomicAdd(&a.counters->replayed, 1ull);
    return;
  }
  if (amb) {  // hand the whole read (pair) to the exact path
    const uint32_t slot = atomicAdd(a.replay_read_cursor, 1u);
    a.replay_reads[slot] = (uint32_t)r;
    int queued = 0;
    for (int q = 0; q < n1; q++) queued += (a.ires[b1 + q].flags & 7) != 4;
    for (int q = 0; q < n2; q++) queued += (a.ires[b2 + q].flags & 7) != 4;
    uint32_t at = queued ? atomicAdd(a.replay_task_cursor, (uint32_t)queued) : 0u;
    for (int q = 0; q < n1; q++)
      if (!settle(b1 + q, 2u * r)) { a.replay_tasks[at].task = b1 + q; a.replay_tasks[at].rm = 2u * r; at++; }
    for (int q = 0; q < n2; q++)
      if (!settle(b2 + q, 2u * r + 1u)) { a.replay_tasks[at].task = b2 + q; a.replay_tasks[at].rm = 2u * r + 1u; at++; }
    atomicAdd(&a.counters->replayed, (unsigned long long)((n1 > 0) + (n2 > 0)));
    return;
  }
  uint32_t m1 = 0, m2 = 0;
  if (keep1 >= 0) {
    const ITaskResult ir = a.ires[b1 + keep1];
    m1 = a.tasks[b1 + keep1].wstart + (uint32_t)ir.maxi + 1u;
    TaskResult t64;
    t64.score = (double)ir.score36 / 36.0;
    t64.maxi = ir.maxi;
    t64.maxk = ir.maxk;
    a.results64[b1 + keep1] = t64;
    if (ir.flags & 4) {
      const uint32_t w = atomicAdd(a.diag_cursor, 1u);
      a.diag_winners[w].task = b1 + (uint32_t)keep1;
      a.diag_winners[w].rm = 2u * (uint32_t)r;
    } else {
      const uint32_t w = atomicAdd(a.winner_cursor, 1u);
      a.winners[w].task = b1 + (uint32_t)keep1;
      a.winners[w].rm = 2u * (uint32_t)r;
    }
  }
  if (keep2 >= 0) {
    const ITaskResult ir = a.ires[b2 + keep2];
    m2 = a.tasks[b2 + keep2].wstart + (uint32_t)ir.maxi + 1u;
    TaskResult t64;
    t64.score = (double)ir.score36 / 36.0;
    t64.maxi = ir.maxi;
    t64.maxk = ir.maxk;
    a.results64[b2 + keep2] = t64;
    if (ir.flags & 4) {
      const uint32_t w = atomicAdd(a.diag_cursor, 1u);
      a.diag_winners[w].task = b2 + (uint32_t)keep2;
      a.diag_winners[w].rm = 2u * (uint32_t)r + 1u;
    } else {
      const uint32_t w = atomicAdd(a.winner_cursor, 1u);
      a.winners[w].task = b2 + (uint32_t)keep2;
      a.winners[w].rm = 2u * (uint32_t)r + 1u;
    }
  }
  a.m1[r] = m1;
  a.m2[r] = m2;
  a.mapping_type[r] = call;
}

// ---------------------------------------------------------------------------------------------------
// Pileup of the pure-diagonal winners: smith_waterman_backtrack (1752-1965) when every step is the state-0
// diagonal step (1846-1858): read base j-1 is counted at window row i-1 for j = mm .. 1 while i >= 1.
// One warp per winner, consecutive lanes hit consecutive sites (24-byte stride) with fire-and-forget atomics.
// ---------------------------------------------------------------------------------------------------

struct DiagArgs {
  const Task* tasks;
  const ITaskResult* ires;
  const Winner* winners;
  const uint32_t* n_items;
  const char* reads[2];
  const int* len[2];
  int stride;
  uint32_t* counts;
  SeedCounters* counters;
};

__global__ void __launch_bounds__(256) k_apply_diag(DiagArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_items = *a.n_items;
  for (uint32_t item = gw; item < n_items; item += nw) {
    const uint32_t task_id = a.winners[item].task;
    const Task tk = a.tasks[task_id];
    const int orient = (int)(tk.rm >> 31);
    const uint32_t rm = tk.rm & 0x7FFFFFFFu;
    const int mm = ((rm & 1) ? a.len[1] : a.len[0])[rm >> 1];
    const char* read = ((rm & 1) ? a.reads[1] : a.reads[0]) + (size_t)(rm >> 1) * a.stride;
    const int maxi = (int)a.ires[task_id].maxi;
    const int steps = maxi < mm ? maxi : mm;  // the walk ends at i == 0 or j == 0
    for (int t = lane; t < steps; t += 32) {
      const int j1 = mm - 1 - t, i1 = maxi - 1 - t;
      const char ch = seq_char(read, mm, orient, j1);
      const int col = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
      if (col >= 0) atomicAdd(&a.counts[((size_t)tk.wstart + (size_t)i1) * 6 + col], 1u);
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&a.counters->diag_traced, (unsigned long long)n_items);
}

}  // namespace pm
