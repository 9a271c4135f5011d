// trace_int.cuh - integer traceback of the winners that are not pure diagonals (sm_100a).
//
// The walker and the tie certification come first, the kernel pair k_trace_dp16 + k_trace_walk16 (two winners per
// lane, deferred band pass, cooperative walk) further down.
//
// Replaces smith_waterman_backtrack (pemapper.c:1752-1965) for winners with gaps.  The DP of the winning
// (read, window) task is recomputed in exact integers (units of 1/36, sw_int16.cuh) with the same sub-warp
// wavefront as sw_wavefront.cuh, storing 6 decision bits per cell:
//     bits 0-1  A  = argmax_k S_k[i][j], priority 0 > 1 > 2                  (consulted from state 0, 1799-1813)
//     bit  2    X1 = S1[i][j] - ge > S0[i][j] - go                            (consulted from state 1, 1823-1831)
//     bit  3    X2 = S2[i][j] - ge > S0[i][j] - go                            (consulted from state 2, 1814-1822)
//     bit  4    S1 == S0          } the A decision compared equal integers
//     bit  5    S2 == max(S0,S1)  }
//     A == 3    marks a cell where an X decision compared equal integers (S1 - ge == S0 - go or S2 - ge == S0 - go)
// (layout in trace_walk.cuh).  The words of the PM_BAND_LANES lanes around the winner's end diagonal are kept in
// shared memory, so the walk never waits on global memory.
// The reference decides on rounded doubles with strict '>', so a comparison of rationally equal values may go
// either way.  Lane 0 therefore walks the path twice: a dry pass that only looks for a
// tie on a consulted decision or a step outside the band, and, if there was none, the pass that applies the pileup
// increments.  An A tie is first put to resolve_tie(): the tied values are traced back in lock step, and if their
// histories join again and are provably EQUAL doubles (three certificates, see resolve_tie) the reference's strict '>'
// keeps the lower state - exactly the integer argmax.  Typical case: a gap inside a homopolymer run or a short
// tandem repeat.  Only ties that cannot be certified hand the winner to the exact fp64 traceback kernel.  Every consulted comparison then has integer
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
  uint32_t* work;              // zeroed per launch: next winner (pair)
  Winner* exact_winners;       // winners that need the fp64 traceback
  uint32_t* exact_cursor;
  const char* reads[2];
  const int* len[2];
  int stride;
  const char* genome;
  PileSink sink;               // sink.pend: per group PM_DP_MAX bytes
  SeedCounters* counters;
  int band_half;               // lanes kept on each side of the end-diagonal lane (PM_BAND_LANES / 2)
  uint2* flagq;                // k_trace_dp16 -> k_trace_walk16: packed decision flags, [pair][winner][band row][lane]
  unsigned char* pair_codes;   // one-hot codes of every pair's windows and oriented reads (low nibble first winner)
  uint4* walk_meta;            // per winner, 2 quads: {wstart, rm | orient << 31, maxi | maxk << 16, mm},
                               // {score36, (dmid + 1024) | bad << 16, nn, task}: all the walk kernel needs, one round trip
  uint32_t* work_walk;         // zeroed per launch: next winner of the walk kernel
  DevParams p;
};

template <int G, int WD>
__host__ __device__ constexpr int trace_rows() { return (G * WD + 24 < PM_DP_MAX) ? G * WD + 24 : PM_DP_MAX; }
template <int G, int WD>
__host__ __device__ constexpr size_t trace_band_bytes() { return (size_t)(128 / G) * trace_rows<G, WD>() * PM_BAND_LANES * 8; }

// ---- exactness bookkeeping for resolve_tie ------------------------------------------------------------------
// Values are rationals x / 36; the doubles the reference holds differ from them by a few ulps at most.
// binade36: number of powers of two 2^e (e >= -5) that are <= |x| / 36; 0 for x = 0.  A rational exactly on a
// power of two may have its double on either side: on_boundary36.
__device__ __forceinline__ int binade36(int x) {
  const unsigned a = (unsigned)(x < 0 ? -x : x);
  if (a < 9u) return a >= 5u ? 4 : a >= 3u ? 3 : (int)a;  // thresholds 36 * 2^e rounded up: 1, 2, 3, 5
  return 5 + (31 - __clz(a / 9u));                         // 9, 18, 36, 72, ...
}
__device__ __forceinline__ bool on_boundary36(int x) {
  const unsigned a = (unsigned)(x < 0 ? -x : x);
  if (a < 9u || a % 9u) return false;
  const unsigned q = a / 9u;
  return (q & (q - 1u)) == 0u;
}
// adding an integer-valued double (+-1.0, -2.0) to a double is exact unless the result lands in a higher binade
__device__ __forceinline__ bool exact_step(int from36, int to36) {
  return binade36(to36) <= binade36(from36) - (on_boundary36(from36) ? 1 : 0);
}

struct ChainPos {
  int i, j, k, r;   // state k of cell (i, j) holds the rational value r / 36
  int end;          // 0 walking, 1 ended on an exact constant (column 0), 2 ended on the row-0 border cell (0, j)
  bool exact;       // every operation so far was an exact double addition (rule 1)
  bool same;        // every value so far lies strictly inside the binade of the tied value (rule 2)
  bool pure;        // every operation so far was +-1.0 or -2.0, whatever the binades (rule 3)
  unsigned sig;     // the rounding operations so far, in order: base-3 digits 1 = mismatch (-1/3), 2 = extension (-1/36)
  int n_round;
};

struct TieCtx {
  const unsigned char* win;    // one-hot reference codes of the window rows (shared memory), one byte per row
  const unsigned char* qcode;  // one-hot codes of the oriented read (shared memory), one byte per column
  int shift;                   // the winner's code sits in bits shift..shift+3 of each byte
};

__device__ __forceinline__ bool cells_match(const TieCtx& t, int i, int j) {
  return (((unsigned)t.qcode[j - 1] & (unsigned)t.win[i - 1]) >> t.shift & 15u) != 0u;
}

// set of states holding the maximum of a cell, from its decision bits (A != 3)
__device__ __forceinline__ int top_set(int c) {
  const int ak = c & 3;
  if (ak == 2) return 4;
  int t = 1 << ak;
  if (c & 16) t |= 3;   // S1 == S0 (then ak == 0)
  if (c & 32) t |= 4;   // S2 == max(S0, S1)
  return t;
}

#define PM_TIE_LIST 12
#define PM_TIE_BUDGET 384
#define PM_TIE_ROUNDS 18       // 3^18 < 2^32

// book one operation "prev -> p.r" of a history: is_round = it adds a rounded constant (digit 1 or 2)
__device__ __forceinline__ bool chain_book(ChainPos& p, int prev, int digit, int b0) {
  if (digit) {
    p.exact = false;
    p.pure = false;
    if (++p.n_round > PM_TIE_ROUNDS) p.same = false;
    p.sig = p.sig * 3u + (unsigned)digit;
  } else if (!exact_step(prev, p.r)) {
    p.exact = false;
  }
  if (binade36(prev) != b0 || on_boundary36(prev)) p.same = false;
  return p.exact || p.same || p.pure;
}

// One backward step of a value's history.  Returns false when neither certificate can hold any more (see
// resolve_tie), the step leaves the band, or it consults a decision that is itself undecidable here.
// Sub-ties met on the way (a predecessor cell whose maximum is shared) are appended to the work list.
template <class Cell>
__device__ __forceinline__ bool chain_step(const Cell& cell, const TieCtx& t, ChainPos& p, int b0, int* wl_i, int* wl_j,
                                           int& wl_n) {
  if (p.j == 0) { p.end = 1; return true; }             // S0[i][0] = S1[i][0] = 0, S2[i][0] = -go: exact constants
  if (p.i == 0) { p.end = 2; p.pure = false; return true; }  // S*[0][j] = -(go + (j-1) ge): one rounded constant per column
  if (p.k == 0) {
    const bool match = cells_match(t, p.i, p.j);
    const int prev = p.r - (match ? 36 : -12);          // M[i-1][j-1]
    if (!chain_book(p, prev, match ? 0 : 1, b0)) { PM_WHY(match ? 11 : 1); return false; }
    p.i--; p.j--; p.r = prev;
    if (p.i == 0 || p.j == 0) { p.k = 0; return true; } // M of a border cell: every state there is handled above
    const int c = cell(p.i, p.j);
    if (c < 0 || (c & 3) == 3) { PM_WHY(2); return false; }
    const int ts = top_set(c);
    if (ts & (ts - 1)) {                                // shared maximum: certify it later
      bool seen = false;
      for (int q = 0; q < wl_n; q++) seen |= (wl_i[q] == p.i && wl_j[q] == p.j);
      if (!seen) {
        if (wl_n >= PM_TIE_LIST) { PM_WHY(3); return false; }
        wl_i[wl_n] = p.i; wl_j[wl_n] = p.j; wl_n++;
      }
    }
    p.k = c & 3;
    return true;
  }
  // gap states: S1[i][j] = max(S0[i-1][j] - go, S1[i-1][j] - ge), S2[i][j] = max(S0[i][j-1] - go, S2[i][j-1] - ge)
  const int pi = p.k == 1 ? p.i - 1 : p.i, pj = p.k == 1 ? p.j : p.j - 1;
  if (pi == 0 || pj == 0) {
    // from a border cell: opening from S0 = 0 (column 0) is exact; everything else involves a rounded value
    if (p.k == 2 && pj == 0) {                          // S2[i][1] = max(0 - go, -go - ge) = -go, exact
      const int prev = p.r + 72;
      if (!chain_book(p, prev, 0, b0)) { PM_WHY(4); return false; }
      p.i = pi; p.j = pj; p.k = 0; p.r = prev;
      return true;
    }
    { PM_WHY(5); return false; }
  }
  const int c = cell(pi, pj);
  if (c < 0 || (c & 3) == 3 || (c & (p.k == 1 ? 64 : 128))) { PM_WHY(2); return false; }
  if (c & (p.k == 1 ? 4 : 8)) {                         // extension: - 1/36 rounds; the gap state continues
    const int prev = p.r + 1;
    if (!chain_book(p, prev, 2, b0)) { PM_WHY(6); return false; }
    p.i = pi; p.j = pj; p.r = prev;
    return true;
  }
  const int prev = p.r + 72;                            // opening: - 2.0
  if (!chain_book(p, prev, 0, b0)) { PM_WHY(7); return false; }
  p.i = pi; p.j = pj; p.k = 0; p.r = prev;
  return true;
}

// Are the doubles of the states sharing the maximum of cell (i, j), of rational value r36 / 36, provably equal?
// The tied values are traced back in lock step until their histories join.  Three certificates:
//  rule 1  both histories consist of exact double additions only (+-1.0, -2.0, never into a higher binade): both
//          values are the common ancestor (or an exact border constant) plus the same integer;
//  rule 2  both histories apply the same rounded constants (-1/3, -1/36) in the same order, merely interleaved
//          differently with integer additions, and every value involved lies strictly inside one binade: there
//          fl(y + c) - (y + c) depends only on y modulo the (common) ulp, which integer shifts leave alone (they are
//          even multiples of the ulp, so round-half-even is preserved too); by induction the two values are equal;
//  rule 3  both histories consist of +-1.0 / -2.0 only, in any binade, and so does EVERYTHING below the point where
//          they join, down to an exact column-0 constant (0 or -2.0): then every value involved is an integer-valued
//          double and every operation on it is exact.  The part below the join is the rest of the walk itself, so
//          the certificate is conditional (return value 2) and the walker gives the winner up if a later step is a
//          mismatch, a gap extension or a start on a rounded row-0 border.  Ties near the start of the read, where
//          the binades of rule 1 are a few cells wide, are the typical case.
// Returns 0 (cannot certify), 1 (equal doubles) or 2 (equal doubles provided the rest of the walk is all-integer;
// only when allow_cond).
template <class Cell>
__device__ int resolve_tie(const Cell& cell, const TieCtx& t, int i, int j, int r36, bool allow_cond) {
  int wl_i[PM_TIE_LIST], wl_j[PM_TIE_LIST], wl_r[PM_TIE_LIST];
  int wl_n = 1, budget = PM_TIE_BUDGET;
  bool cond = false;
  wl_i[0] = i; wl_j[0] = j; wl_r[0] = r36;
  for (int w = 0; w < wl_n; w++) {
    const int c = cell(wl_i[w], wl_j[w]);
    if (c < 0 || (c & 3) == 3) { PM_WHY(2); return 0; }
    const int ts = top_set(c);
    const int first = __ffs(ts) - 1;
    const int b0 = binade36(wl_r[w]);
    const bool b0_ok = !on_boundary36(wl_r[w]);
    // every other state of the top set against the lowest one
    for (int other = first + 1; other < 3; other++) {
      if (!(ts & (1 << other))) continue;
      ChainPos A, B;
      A.i = B.i = wl_i[w]; A.j = B.j = wl_j[w]; A.r = B.r = wl_r[w];
      A.k = first; B.k = other; A.end = B.end = 0;
      A.exact = B.exact = true;
      A.pure = B.pure = true;
      A.same = B.same = b0_ok;
      A.sig = B.sig = 0u;
      A.n_round = B.n_round = 0;
      for (;;) {
        const bool joined = (!A.end && !B.end && A.i == B.i && A.j == B.j && A.k == B.k) ||
                            (A.end == 2 && B.end == 2 && A.j == B.j);  // same cell and state / same row-0 border cell
        if (joined) {
          if (A.exact && B.exact) break;
          if (A.same && B.same && A.sig == B.sig && A.n_round == B.n_round) break;
          if (allow_cond && A.pure && B.pure && !A.end) {  // rule 3: the walker checks the part below the join
            cond = true;
            break;
          }
          { PM_WHY(8); return 0; }
        }
        if (A.end && B.end) {
          // two exact column-0 constants plus integers: whole histories of integer-valued doubles
          if (A.end == 1 && B.end == 1 && ((A.exact && B.exact) || (A.pure && B.pure))) break;
          { PM_WHY(9); return 0; }
        }
        if (--budget < 0) { PM_WHY(10); return 0; }
        // advance the one farther from the origin (the only one that can still reach the other)
        const bool stepA = !A.end && (B.end || A.i + A.j > B.i + B.j || (A.i + A.j == B.i + B.j && A.i >= B.i));
        ChainPos& p = stepA ? A : B;
        const int n0 = wl_n;
        if (!chain_step(cell, t, p, b0, wl_i, wl_j, wl_n)) return 0;
        if (wl_n > n0) wl_r[n0] = p.r;                 // value of the shared maximum just queued
      }
    }
  }
  return cond ? 2 : 1;
}

// dry walk of the integer kernel: like walk_path<false> but tracks the rational value along the path so that A ties
// can be put to resolve_tie()
template <class Cell>
__device__ int walk_check_int(const Cell& cell, const TieCtx& t, int k, int i, int j, int r36) {
  while (i > 0 && j > 0) {
    int pi, pj, pk = 0;
    if (k == 0) {
      pi = i - 1; pj = j - 1;
      r36 -= cells_match(t, i, j) ? 36 : -12;          // value of M[i-1][j-1]
      if (pi > 0 && pj > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if ((c & 3) == 3) return PM_WALK_TIE;
        if ((c & 48) && (top_set(c) & (top_set(c) - 1)) && !resolve_tie(cell, t, pi, pj, r36, false)) return PM_WALK_TIE;
        pk = c & 3;
      }
    } else if (k == 2) {
      pi = i; pj = j - 1;
      if (pj > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if ((c & 3) == 3) return PM_WALK_TIE;
        pk = (c & 8) ? 2 : 0;
      }
      r36 += pk == 2 ? 1 : 72;
    } else {
      pi = i - 1; pj = j;
      if (pi > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if ((c & 3) == 3) return PM_WALK_TIE;
        pk = (c & 4) ? 1 : 0;
      }
      r36 += pk == 1 ? 1 : 72;
    }
    i = pi; j = pj; k = pk;
  }
  return PM_WALK_OK;
}

// ---------------------------------------------------------------------------------------------------
// Packed integer traceback: k_trace_dp16 + k_trace_walk16.
//
// TWO winners are packed per lane (s16x2, biased by PM_TBIAS like sw_int16.cuh) and the decision flags are extracted
// with SWAR compares: for halves a, b in [0, 2^15), (a + 0x8000 - b) has bit 15 set iff a >= b, and neither half
// borrows from the other.  Six flags per cell and half:
//     F0  S1 > S0                 F1  S2 > max(S0,S1)     F2  X1     F3  X2
//     F4  S1 == S0, or X1 tie     F5  S2 == max(S0,S1), or X2 tie
// (S1 >= S0 implies X1, so F4 without F2 is free to mean "S1 - ge == S0 - go"; likewise F5 without F3 for X2.  The
// accessor turns them into bits 6 / 7; the A decision of such a cell stays usable.)
//
// Only the cells near the winners' end diagonal are ever consulted by the walk: lane l (columns WD*l+1 .. WD*l+WD)
// is "in the band" for the nb = (half+1)*WD rows starting at r0(l) = WD*l - WD/2 + dmid + 1, and at any step of the
// wavefront only one or two lanes of a group are.  Computing the flags inside the wavefront makes all G lanes pay
// for them (3/4 of the instructions of a row), so k_trace_dp16 splits the work:
//   1. main pass - the plain scoring recurrence (10 instructions per packed cell) over the whole matrix; a lane
//      entering its band rows parks its column state (S0, S1, M of the row above: 3*WD words) and, for every band
//      row, the three words that arrive from its left neighbour (S0, S2 of the row, M of the row above);
//   2. band pass - every lane recomputes ITS OWN band rows from the parked inputs, now with the flags: no lane
//      depends on another one any more, all G lanes are busy, and the flags cost nb rows per lane instead of
//      nn + G - 1.  The values are the same integers, so the flags are the ones a single pass would produce.
// Shared memory per group: [2*WD + ceil(WD/4)][G] uint4: band row r of lane l holds {l_s0, l_s2, diag, top-state
// word r}; the last WD top-state words sit in the extra quads.  The packed flags of a band row,
// {A: F0|F1<<WD|F2<<2WD, A: F3|F4<<WD|F5<<2WD, B: ..., B: ...}, go to global memory ([pair][row][lane], one
// coalesced line per row and group) for the walk kernel.
//
// The walk is a different kind of program (branchy, latency-bound, few registers), and one kernel holding both
// phases overflowed the instruction caches (44 % of its stall samples were instruction fetch).  k_trace_walk16
// gives G/2 lanes to every winner: they stage the winner's flag words in shared memory and walk cooperatively
// (trace_walk.cuh), with the tie certification above.
// ---------------------------------------------------------------------------------------------------
#define PM_TBIAS 1024

template <int G, int WD>
__host__ __device__ constexpr int defer_quads() { return 2 * WD + (WD + 3) / 4; }
template <int G, int WD>
__host__ __device__ constexpr size_t trace_dp16_smem() { return (size_t)(128 / G) * defer_quads<G, WD>() * G * 16; }
template <int G, int WD>
__host__ __device__ constexpr int trace_code_bytes() { return (trace_rows<G, WD>() + G * WD + 15) / 16 * 16; }
template <int G, int WD>
__host__ __device__ constexpr size_t trace_walk16_smem() { return (size_t)4 * (2 * WD * G * 8 + trace_code_bytes<G, WD>()); }
template <int G, int WD>
__host__ __device__ constexpr size_t trace_flag_bytes_per_pair() { return (size_t)2 * WD * G * 16; }

// accessor over one winner's flag words: [band row][lane] uint2
template <int G, int WD>
struct LaneBandCell {
  const uint2* fl;
  int dmid, nb;          // nb = band rows per lane = (half + 1) * WD
  __device__ __forceinline__ int operator()(int pi, int pj) const {
    const int l = (pj - 1) / WD, c = (pj - 1) - l * WD;
    const int r = pi - (WD * l - WD / 2 + dmid + 1);
    if (r < 0 || r >= nb) return -1;
    const uint2 w = fl[r * G + l];
    const int f0 = (w.x >> c) & 1, f1 = (w.x >> (WD + c)) & 1, f2 = (w.x >> (2 * WD + c)) & 1;
    const int f3 = (w.y >> c) & 1, f4 = (w.y >> (WD + c)) & 1, f5 = (w.y >> (2 * WD + c)) & 1;
    // S1 >= S0 implies X1 and S2 >= max(S0,S1) implies X2, so "equal" without the X flag is free to mark an X tie:
    // bits 6 / 7 = the X1 / X2 comparison met equal integers (only a walk that ENTERS the cell in a gap state cares)
    return (f1 ? 2 : f0) | (f2 << 2) | (f3 << 3) | ((f4 & f2) << 4) | ((f5 & f3) << 5) | ((f4 & ~f2 & 1) << 6) |
           ((f5 & ~f3 & 1) << 7);
  }
};

struct IntTie {
  static constexpr bool kTrack = true;
  TieCtx t;
  __device__ __forceinline__ bool match(int i, int j) const { return cells_match(t, i, j); }
  template <class Cell>
  __device__ __forceinline__ int resolve(const Cell& cell, int pi, int pj, int r36) const {
    return resolve_tie(cell, t, pi, pj, r36, true);
  }
};

template <int G, int WD>
__global__ void __launch_bounds__(128) k_trace_dp16(TraceIntArgs a) {
  constexpr int GPB = 128 / G;
  constexpr int ROWS = trace_rows<G, WD>();
  constexpr int QUADS = defer_quads<G, WD>();
  extern __shared__ uint4 s_park[];           // [GPB][QUADS][G]
  constexpr int CB = trace_code_bytes<G, WD>();
  __shared__ __align__(16) unsigned char s_codes[GPB][CB];  // one-hot codes, low nibble first winner, high nibble second:
                                                            // [0, ROWS) window rows, [ROWS, ROWS + G*WD) read columns
  const int tid = threadIdx.x, grp = tid / G, gl = tid % G;
  const uint32_t n_items = *a.n_items, n_pairs = (n_items + 1) >> 1;
  unsigned char* win = s_codes[grp];
  unsigned char* qcodes = s_codes[grp] + ROWS;
  uint4* quads = s_park + (size_t)grp * QUADS * G;
  const int bis = a.p.is_bisulfite;
  const int half = a.band_half < 1 ? a.band_half : 1;  // lanes kept = half + 1
  const int nb = (half + 1) * WD;                      // band rows per lane
  constexpr uint32_t K1 = 0x00010001u, K12 = 0x000C000Cu, NEG72 = 0xFFB8FFB8u, H = 0x80008000u;
  constexpr uint32_t BIASP = (PM_TBIAS << 16) | PM_TBIAS;
  constexpr uint32_t HX = H + 70u * K1;  // x = S + 70 >= S0  <=>  S - ge > S0 - go
  constexpr unsigned FULL = 0xFFFFFFFFu;

  // The sub-warps of a warp take consecutive pairs and run every loop below with warp-uniform trip counts and
  // full-mask shuffles (width G), so that they stay in lock step: a sub-warp without a pair runs empty rows.
  for (;;) {
    uint32_t first;
    const uint32_t pair = next_work_item_warp<G>(a.work, &first);
    if (first >= n_pairs) break;
    const bool have = pair < n_pairs;
    const uint32_t itA = have ? 2 * pair : 0, itB = have ? ((2 * pair + 1 < n_items) ? 2 * pair + 1 : 2 * pair) : 0;
    const uint32_t idA = a.winners[itA].task, idB = a.winners[itB].task;
    const Task tA = a.tasks[idA], tB = a.tasks[idB];
    const TaskResult rA = a.results[idA], rB = a.results[idB];
    const int orA = (int)(tA.rm >> 31), orB = (int)(tB.rm >> 31);
    const uint32_t rmA = tA.rm & 0x7FFFFFFFu, rmB = tB.rm & 0x7FFFFFFFu;
    const int mmA = ((rmA & 1u) ? a.len[1] : a.len[0])[rmA >> 1], mmB = ((rmB & 1u) ? a.len[1] : a.len[0])[rmB >> 1];
    const char* readA = ((rmA & 1u) ? a.reads[1] : a.reads[0]) + (size_t)(rmA >> 1) * a.stride;
    const char* readB = ((rmB & 1u) ? a.reads[1] : a.reads[0]) + (size_t)(rmB >> 1) * a.stride;
    const int nnA = have ? (rA.maxi < tA.blen ? rA.maxi : tA.blen) : 0;  // rows below the winning cell are never consulted
    const int nnB = have ? (rB.maxi < tB.blen ? rB.maxi : tB.blen) : 0;
    const int nn = nnA > nnB ? nnA : nnB;
    const int dmid = ((rA.maxi - mmA) + (rB.maxi - mmB)) >> 1;  // one band for both winners
    int nn_w = nn;
    if (G < 32) nn_w = max(nn_w, __shfl_xor_sync(FULL, nn_w, 16));

    __syncwarp();
    bool badA = false, badB = false;  // a base outside ACGTN: the walk kernel hands such winners to the fp64 path
    for (int i = gl; i < nn_w; i += G) {
      uint32_t cA = 0, cB = 0;
      if (i < nnA) {
        const char ch = a.genome[(size_t)tA.wstart + i];
        cA = ch == 'A' ? 1u : ch == 'C' ? (bis ? 10u : 2u) : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
        badA |= (cA == 0u);
      }
      if (i < nnB) {
        const char ch = a.genome[(size_t)tB.wstart + i];
        cB = ch == 'A' ? 1u : ch == 'C' ? (bis ? 10u : 2u) : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
        badB |= (cB == 0u);
      }
      if (i < nn) win[i] = (unsigned char)(cA | (cB << 4));
    }
    uint32_t q[WD], s0u[WD], s1u[WD], mu[WD];
    const int jbase = gl * WD;
#pragma unroll
    for (int c = 0; c < WD; c++) {
      const int j0 = jbase + c;
      uint32_t cA = 0, cB = 0;
      if (j0 < mmA) {
        const char ch = oriented_char(readA, mmA, orA, j0);
        cA = ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
        badA |= (cA == 0u);
      }
      if (j0 < mmB) {
        const char ch = oriented_char(readB, mmB, orB, j0);
        cB = ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
        badB |= (cB == 0u);
      }
      q[c] = cA | (cB << 16);
      qcodes[j0] = (unsigned char)(cA | (cB << 4));
      const uint32_t b = (uint32_t)(PM_TBIAS - 72 - j0);  // S*[0][j] = -(72 + j - 1), j = j0 + 1 (2073-2081)
      s0u[c] = b | (b << 16);
      s1u[c] = s0u[c];
      mu[c] = s0u[c] - K12;  // mu holds max(S0,S1,S2) - 12
    }
    {
      const unsigned gm = (G == 32) ? FULL : (((1u << G) - 1u) << ((tid & 31) / G * G));
      badA = (__ballot_sync(FULL, badA) & gm) != 0u;
      badB = (__ballot_sync(FULL, badB) & gm) != 0u;
    }
    uint32_t out_s0 = 0, out_s2 = 0, out_m = 0;
    __syncwarp();
    if (have && gl == 0) {  // what the walk kernel needs of the two winners (written now: the registers die here)
      uint4* meta = a.walk_meta + (size_t)pair * 4;
      meta[0] = make_uint4(tA.wstart, tA.rm, (uint32_t)rA.maxi | ((uint32_t)rA.maxk << 16), (uint32_t)mmA);
      meta[1] = make_uint4((uint32_t)(int)lrint(rA.score * 36.0), (uint32_t)(dmid + 1024) | (badA ? (1u << 16) : 0u), (uint32_t)nnA, idA);
      meta[2] = make_uint4(tB.wstart, tB.rm, (uint32_t)rB.maxi | ((uint32_t)rB.maxk << 16), (uint32_t)mmB);
      meta[3] = make_uint4((uint32_t)(int)lrint(rB.score * 36.0), (uint32_t)(dmid + 1024) | (badB ? (1u << 16) : 0u), (uint32_t)nnB, idB);
      atomicAdd(&a.counters->tb_cells_int, (unsigned long long)nnA * (unsigned long long)mmA +
                                               (itB != itA ? (unsigned long long)nnB * (unsigned long long)mmB : 0ull));
    }
    if (have) {  // the codes travel with the flags
      uint4* dst = reinterpret_cast<uint4*>(a.pair_codes + (size_t)pair * CB);
      const uint4* srcq = reinterpret_cast<const uint4*>(s_codes[grp]);
      for (int x = gl; x < CB / 16; x += G) dst[x] = srcq[x];
    }

    const int r0 = WD * gl - WD / 2 + dmid + 1;  // first band row of this lane
    // ---- main pass: values only
    const int steps_w = nn_w > 0 ? nn_w + G - 1 : 0;
    for (int s = 0; s < steps_w; s++) {
      uint32_t l_s0 = __shfl_up_sync(FULL, out_s0, 1, G);
      uint32_t l_s2 = __shfl_up_sync(FULL, out_s2, 1, G);
      uint32_t diag = __shfl_up_sync(FULL, out_m, 1, G);
      if (gl == 0) {  // column 0: S0 = 0, S2 = -72, M(row above) - 12 = -12 (2062-2081)
        l_s0 = BIASP;
        l_s2 = BIASP - 0x00480048u;
        diag = BIASP - K12;
      }
      const int i = s - gl + 1;
      if (i >= 1 && i <= nn) {
        const unsigned r = (unsigned)(i - r0);
        if (r < (unsigned)nb) {
          uint32_t* slot = reinterpret_cast<uint32_t*>(&quads[r * G + gl]);
          if (r == 0u && i >= 2) {  // entering the band: park the column state of row i - 1
#pragma unroll
            for (int c = 0; c < WD; c++) {
              reinterpret_cast<uint32_t*>(&quads[c * G + gl])[3] = s0u[c];
              reinterpret_cast<uint32_t*>(&quads[(WD + c) * G + gl])[3] = s1u[c];
              reinterpret_cast<uint32_t*>(&quads[(2 * WD + c / 4) * G + gl])[c & 3] = mu[c];
            }
          }
          *reinterpret_cast<uint2*>(slot) = make_uint2(l_s0, l_s2);
          slot[2] = diag;
        }
        const uint32_t rb = win[i - 1];
        const uint32_t rc = (rb & 15u) | ((rb >> 4) << 16);
#pragma unroll
        for (int c = 0; c < WD; c++) {
          const uint32_t s2 = __viaddmax_s16x2(l_s0, NEG72, l_s2 - K1);      // 1710 / 1720
          const uint32_t s1 = __viaddmax_s16x2(s0u[c], NEG72, s1u[c] - K1);  // 1711 / 1721
          const uint32_t m01 = __vminu2(q[c] & rc, K1);                      // 1 where the bases match
          const uint32_t s0 = m01 * 48u + diag;                              // (M - 12) + 48 or + 0 (1713 / 1723)
          diag = mu[c];
          const uint32_t m = __vimax3_s16x2(s0, s1, s2);                      // one VIMNMX3.S16x2
          s0u[c] = s0;
          s1u[c] = s1;
          mu[c] = m - K12;
          l_s0 = s0;
          l_s2 = s2;
        }
        out_s0 = l_s0;
        out_s2 = l_s2;
        out_m = diag;
      }
    }
    __syncwarp();
    // ---- band pass: this lane's band rows again, with the flags
    {
#pragma unroll
      for (int c = 0; c < WD; c++) {
        if (r0 >= 2) {
          s0u[c] = reinterpret_cast<const uint32_t*>(&quads[c * G + gl])[3];
          s1u[c] = reinterpret_cast<const uint32_t*>(&quads[(WD + c) * G + gl])[3];
          mu[c] = reinterpret_cast<const uint32_t*>(&quads[(2 * WD + c / 4) * G + gl])[c & 3];
        } else {  // the band starts at row 1: row 0 borders (2073-2081)
          const uint32_t b = (uint32_t)(PM_TBIAS - 72 - (jbase + c));
          s0u[c] = b | (b << 16);
          s1u[c] = s0u[c];
          mu[c] = s0u[c] - K12;
        }
      }
      uint2* outA = a.flagq + (size_t)pair * (2 * 2 * WD * G);  // [winner][band row][lane]
      uint2* outB = outA + 2 * WD * G;
      for (int r = 0; r < nb; r++) {  // warp-uniform trip count; rows outside 1 .. nn are never consulted
        const int i = r0 + r;
        if (i >= 1 && i <= nn) {
          const uint4 in = quads[r * G + gl];
          uint32_t l_s0 = in.x, l_s2 = in.y, diag = in.z;
          const uint32_t rb = win[i - 1];
          const uint32_t rc = (rb & 15u) | ((rb >> 4) << 16);
          uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
#pragma unroll
          for (int c = 0; c < WD; c++) {
            const uint32_t s2 = __viaddmax_s16x2(l_s0, NEG72, l_s2 - K1);      // 1710 / 1720
            const uint32_t s1 = __viaddmax_s16x2(s0u[c], NEG72, s1u[c] - K1);  // 1711 / 1721
            const uint32_t m01 = __vminu2(q[c] & rc, K1);                      // 1 where the bases match
            const uint32_t s0 = m01 * 48u + diag;                              // (M - 12) + 48 or + 0
            diag = mu[c];
            const uint32_t x01 = __vmaxs2(s0, s1);
            const uint32_t m = __vmaxs2(x01, s2);
            // SWAR decisions, bit 15 of each half
            const uint32_t g = s0 + H - s1, g2 = s1 + H - s0;                  // S0 >= S1, S1 >= S0
            const uint32_t hh = x01 + H - s2, h2 = s2 + H - x01;               // max01 >= S2, S2 >= max01
            const uint32_t w1 = s1 + HX - s0, w2 = s2 + HX - s0;               // X1, X2
            const uint32_t tx1 = (w1 + K1) & ~w1, tx2 = (w2 + K1) & ~w2;       // S1 + 71 == S0, S2 + 71 == S0
            a0 = (a0 >> 1) | (~g & H);
            a1 = (a1 >> 1) | (~hh & H);
            a2 = (a2 >> 1) | (w1 & H);
            a3 = (a3 >> 1) | (w2 & H);
            a4 = (a4 >> 1) | (((g & g2) | tx1) & H);
            a5 = (a5 >> 1) | (((hh & h2) | tx2) & H);
            s0u[c] = s0;
            s1u[c] = s1;
            mu[c] = m - K12;
            l_s0 = s0;
            l_s2 = s2;
          }
          // flags sit in bits 16-WD .. 15 (first winner) and 32-WD .. 31 (second) of a0 .. a5
          constexpr uint32_t FM = (1u << WD) - 1u;
          const uint32_t wa0 = ((a0 >> (16 - WD)) & FM) | (((a1 >> (16 - WD)) & FM) << WD) | (((a2 >> (16 - WD)) & FM) << (2 * WD));
          const uint32_t wa1 = ((a3 >> (16 - WD)) & FM) | (((a4 >> (16 - WD)) & FM) << WD) | (((a5 >> (16 - WD)) & FM) << (2 * WD));
          const uint32_t wb0 = (a0 >> (32 - WD)) | ((a1 >> (32 - WD)) << WD) | ((a2 >> (32 - WD)) << (2 * WD));
          const uint32_t wb1 = (a3 >> (32 - WD)) | ((a4 >> (32 - WD)) << WD) | ((a5 >> (32 - WD)) << (2 * WD));
          outA[r * G + gl] = make_uint2(wa0, wa1);
          outB[r * G + gl] = make_uint2(wb0, wb1);
        }
      }
    }
    __syncwarp();
  }
}

// one winner per warp: stage its flag words and base codes, walk, apply or hand over to the exact kernel
template <int G, int WD>
__global__ void __launch_bounds__(128) k_trace_walk16(TraceIntArgs a) {
  constexpr int NL = 32;
  constexpr int WPB = 4;
  constexpr int ROWS = trace_rows<G, WD>();
  constexpr int CB = trace_code_bytes<G, WD>();
  constexpr int FLW = 2 * WD * G;             // flag words (uint2) of one winner
  extern __shared__ uint4 s_walk[];           // [WPB]{FLW uint2, CB bytes}
  const int tid = threadIdx.x, wk = tid / NL, sl = tid % NL;
  constexpr unsigned smask = 0xFFFFFFFFu;
  const uint32_t n_items = *a.n_items;
  const uint32_t gwk = blockIdx.x * WPB + wk;
  const int half = a.band_half < 1 ? a.band_half : 1;
  const int nb = (half + 1) * WD;
  uint4* mine = s_walk + (size_t)wk * ((FLW * 8 + CB) / 16);
  uint2* fl = reinterpret_cast<uint2*>(mine);
  unsigned char* codes = reinterpret_cast<unsigned char*>(mine + FLW * 8 / 16);
  PileSink sink = a.sink;
  sink.pend = a.sink.pend + (size_t)gwk * PM_DP_MAX;   // the walker's path segments
  uint2* segs = reinterpret_cast<uint2*>(sink.pend);

  // winners are taken 8 at a time: one atomic per winner on a single address serialises in L2
  uint32_t next_item = 0, batch_left = 0;
  for (;;) {
    if (batch_left == 0u) {
      uint32_t v = 0;
      if (sl == 0) v = atomicAdd(a.work_walk, 8u);
      next_item = __shfl_sync(smask, v, 0);
      batch_left = 8u;
    }
    const uint32_t item = next_item++;
    batch_left--;
    if (item >= n_items) break;
    const uint32_t pair = item >> 1;
    const int hi = (int)(item & 1u);
    const uint4 m0 = __ldg(a.walk_meta + (size_t)item * 2), m1 = __ldg(a.walk_meta + (size_t)item * 2 + 1);
    const uint32_t wstart = m0.x;
    const int orient = (int)(m0.y >> 31);
    const uint32_t rm = m0.y & 0x7FFFFFFFu;
    const int maxi = (int)(m0.z & 0xFFFFu), maxk = (int)(m0.z >> 16), mm = (int)m0.w;
    const char* read = ((rm & 1u) ? a.reads[1] : a.reads[0]) + (size_t)(rm >> 1) * a.stride;
    const int r36 = (int)m1.x;
    const int dmid = (int)(m1.y & 0xFFFFu) - 1024;
    const bool bad = (m1.y >> 16) & 1u;  // bases outside ACGTN: the scoring kernel sent such reads to the fp64 path already
    const int nn = (int)m1.z;

    __syncwarp();
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.flagq + ((size_t)pair * 2 + hi) * FLW);
      for (int x = sl; x < nb * G / 2; x += NL) mine[x] = __ldg(src + x);
      const uint4* csrc = reinterpret_cast<const uint4*>(a.pair_codes + (size_t)pair * CB);
      for (int x = sl; x < CB / 16; x += NL) mine[FLW * 8 / 16 + x] = __ldg(csrc + x);
    }
    __syncwarp();

    int rc = PM_WALK_TIE, nseg = 0;
    if (!bad && nn > 0) {
      LaneBandCell<G, WD> cell;
      cell.fl = fl;
      cell.dmid = dmid;
      cell.nb = nb;
      IntTie tie;
      tie.t.win = codes;
      tie.t.qcode = codes + ROWS;
      tie.t.shift = 4 * hi;
      rc = coop_walk<NL>(cell, tie, smask, 0, sl, maxk, maxi, mm, r36, segs, &nseg);
      if (rc == PM_WALK_OK && nseg < 0) rc = PM_WALK_TIE;  // more segments than the scratch holds: exact kernel
    }
    if (rc == PM_WALK_OK) {
      coop_apply<NL>(segs, nseg, smask, sl, read, mm, orient, wstart, sink, codes + ROWS, 4 * hi);
    } else if (sl == 0) {
      const uint32_t w = atomicAdd(a.exact_cursor, 1u);
      a.exact_winners[w] = a.winners[item];
      atomicAdd(&a.counters->exact_traced, 1ull);
    }
    __syncwarp();
  }
}

}  // namespace pm
