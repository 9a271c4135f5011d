// trace_walk.cuh - the traceback walk (smith_waterman_backtrack, pemapper.c:1752-1965) over stored decision bits,
// shared by the integer and the fp64 traceback kernels, and the band geometry of the shared-memory decision store.
//
// Decision bits of cell (i, j), i = window row 1..nn, j = read column 1..mm:
//     bits 0-1  A  = argmax_k S_k[i][j], priority 0 > 1 > 2     consulted from state 0 at (i+1, j+1)   (1799-1813)
//     bit  2    X1 = S1[i][j] - ge > S0[i][j] - go              consulted from state 1 at (i+1, j)     (1823-1831)
//     bit  3    X2 = S2[i][j] - ge > S0[i][j] - go              consulted from state 2 at (i, j+1)     (1814-1822)
//     bit  4    (integer kernel) the A decision compared equal integers
//     bit  5    (integer kernel) an X decision compared equal integers
//
// Band store: a lane writes one 64-bit word per row (its WD columns).  Only the PM_BAND_LANES lanes around the lane
// that owns the column of the winner's end diagonal (j = i - (maxi - mm)) are kept, in shared memory; a walk that
// leaves the band (net indel drift > WD columns) reports PM_WALK_OOB and the winner is redone by the kernel that
// keeps every lane's word in global memory.
#pragma once
#include "pemap_common.cuh"

#define PM_BAND_LANES 3
#define PM_WALK_OK 0
#define PM_WALK_TIE 1
#define PM_WALK_OOB 2

namespace pm {

struct PileSink {
  uint32_t* counts;            // [genome_size][6]
  unsigned char* ins_buf;      // insertion records: {u32 pos, u32 len, chars padded to 4}
  unsigned long long* ins_cursor;
  unsigned long long ins_cap;
  char* pend;                  // scratch for the pending insertion characters of this walk (PM_DP_MAX bytes)
};

__device__ __forceinline__ char oriented_char(const char* read, int len, int orient, int j0) {  // j0 = 0-based read index
  return orient ? rt_char(read[len - 1 - j0]) : read[j0];
}

__device__ __forceinline__ void sink_insertion(const PileSink& s, uint32_t site, int n) {
  unsigned long long need = 8ull + (unsigned long long)((n + 3) & ~3);
  unsigned long long off = atomicAdd(s.ins_cursor, need);
  if (off + need <= s.ins_cap) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(s.ins_buf + off);
    hdr[0] = site;
    hdr[1] = (uint32_t)n;
    unsigned char* dst = s.ins_buf + off + 8;
    for (int m = 0; m < n; m++) dst[m] = (unsigned char)s.pend[n - (m + 1)];  // 1892-1893: un-reverse
  }
  atomicAdd(&s.counts[(size_t)site * 6 + 5], 1u);  // no_ins++ (1903 / 1952)
}

// lane that owns the end-diagonal column of row i (may be negative or >= G); floor division
template <int WD>
__device__ __forceinline__ int band_center_lane(int i, int dend) {
  return (i - dend - 1 + 64 * WD) / WD - 64;
}

// Cell: int operator()(int pi, int pj) -> decision bits of cell (pi, pj), or -1 when the cell is not stored.
// APPLY = false: dry run that only reports PM_WALK_TIE / PM_WALK_OOB; APPLY = true: the pileup increments.
// one-hot base code (A 1, C 2, G 4, T 8, N 15, anything else 0) -> pileup column, -1 = counted nowhere (1846-1858)
__device__ __forceinline__ int code_column(unsigned code) {
  return code == 1u ? 0 : code == 2u ? 1 : code == 4u ? 2 : code == 8u ? 3 : -1;
}
__device__ __forceinline__ unsigned base_onehot(char ch) {
  return ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : (ch == 'N' || ch == 'n') ? 15u : 0u;
}

// qcode: one-hot codes of the oriented read in shared memory, one byte per column, the code in bits qshift..qshift+3
// (the walk is a chain of dependent steps on one lane: it must not wait on global memory); the characters themselves
// are only fetched for insertions.
template <bool APPLY, int TIE_BITS, class Cell>
__device__ __forceinline__ int walk_path(const Cell& cell, int k, int i, int j, const char* read, int mm, int orient,
                                         uint32_t wstart, const PileSink& sink, const unsigned char* qcode, int qshift) {
  int n_pend = 0, i1 = 0, j1 = 0;
  while (i > 0 && j > 0) {
    i1 = i - 1;
    j1 = j - 1;
    int pi, pj, pk = 0;
    if (k == 0) {
      pi = i1; pj = j1;
      if (pi > 0 && pj > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if (TIE_BITS && (c & 16)) return PM_WALK_TIE;
        pk = c & 3;
      }
    } else if (k == 2) {
      pi = i; pj = j1;
      if (pj > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if (TIE_BITS && (c & 32)) return PM_WALK_TIE;
        pk = (c & 8) ? 2 : 0;
      }
    } else {
      pi = i1; pj = j;
      if (pi > 0) {
        const int c = cell(pi, pj);
        if (c < 0) return PM_WALK_OOB;
        if (TIE_BITS && (c & 32)) return PM_WALK_TIE;
        pk = (c & 4) ? 1 : 0;
      }
    }
    if (APPLY) {
      const uint32_t site = wstart + (uint32_t)i1;
      if (pi != i) {
        if (pj != j) {  // 1846-1858
          const int col = code_column((qcode[j1] >> qshift) & 15u);
          if (col >= 0) atomicAdd(&sink.counts[(size_t)site * 6 + col], 1u);
        } else {
          atomicAdd(&sink.counts[(size_t)site * 6 + 4], 1u);  // 1868
        }
        if (n_pend > 0) sink_insertion(sink, site, n_pend);  // 1871-1904
        n_pend = 0;
      } else {
        sink.pend[n_pend++] = oriented_char(read, mm, orient, j1);  // 1910-1911
      }
    }
    i = pi; j = pj; k = pk;
  }
  if (APPLY && n_pend > 0 && i >= 1) sink_insertion(sink, wstart + (uint32_t)i1, n_pend);  // 1918-1958
  return PM_WALK_OK;
}

// ---------------------------------------------------------------------------------------------------
// Cooperative walk.  The serial walk above is a chain of ~2 * len dependent shared-memory look-ups on one lane;
// most of it is a state-0 diagonal.  Here NL consecutive lanes own one winner: in state 0 lane t looks at the cell
// t steps down the diagonal, a ballot finds the first cell that is not a plain "stay in state 0" decision, and the
// walk jumps there in one iteration.  Gap steps and cells that need the tie certification are handled one at a
// time, exactly as walk_path / walk_check_int do.  Nothing is applied during the walk: the path is recorded as
// segments (type, i, j, length) and, once the whole walk is known to be decidable here, applied by all NL lanes
// (coop_apply).  Every lane of the walker carries the same (k, i, j): control flow is uniform within the walker.
// ---------------------------------------------------------------------------------------------------
#define PM_WALK_SEGS 40   // recorded segments per walker: PM_DP_MAX bytes of scratch / 8

// sink_insertion for an insertion whose characters are read columns j_lo .. j_lo + n - 1 (0-based, ascending):
// the un-reversed buffer of 1892-1893
__device__ __forceinline__ void sink_insertion_direct(const PileSink& s, uint32_t site, const char* read, int mm,
                                                      int orient, int j_lo, int n) {
  unsigned long long need = 8ull + (unsigned long long)((n + 3) & ~3);
  unsigned long long off = atomicAdd(s.ins_cursor, need);
  if (off + need <= s.ins_cap) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(s.ins_buf + off);
    hdr[0] = site;
    hdr[1] = (uint32_t)n;
    unsigned char* dst = s.ins_buf + off + 8;
    for (int m = 0; m < n; m++) dst[m] = (unsigned char)oriented_char(read, mm, orient, j_lo + m);
  }
  atomicAdd(&s.counts[(size_t)site * 6 + 5], 1u);  // no_ins++ (1903 / 1952)
}

struct NoTie {  // the fp64 kernels: every decision bit is the reference's own double comparison
  static constexpr bool kTrack = false;
  __device__ __forceinline__ bool match(int, int) const { return false; }
  template <class Cell>
  __device__ __forceinline__ int resolve(const Cell&, int, int, int) const { return 1; }
};

// Tie: kTrack (follow the rational value r36 of the path), match(i, j), resolve(cell, pi, pj, r36) -> 0 (undecidable
// here), 1 (decided), 2 (decided provided every later step of the walk is an exact integer operation: a match or a
// gap opening, ending on a column-0 border; trace_int.cuh rule 3).
// smask: the walker's lanes within the warp, base: its first lane, sl: this lane's index in the walker.
// segs: PM_WALK_SEGS records written by lane 0 of the walker; *n_segs = number written, or -1 when the path has more
// segments than that (the caller then applies it with the serial walk_path<true>).
template <int NL, class Cell, class Tie>
__device__ int coop_walk(const Cell& cell, const Tie& tie, unsigned smask, int base, int sl, int k, int i, int j, int r36,
                         uint2* segs, int* n_segs) {
  int nseg = 0, ot = -1, oi = 0, oj = 0, ol = 0;  // ot..ol: the open segment
  bool over = false;
  bool need_pure = false;  // a conditional certificate is outstanding
  auto flush = [&]() {
    if (ot >= 0) {
      if (nseg < PM_WALK_SEGS) {
        if (sl == 0) segs[nseg] = make_uint2((unsigned)oi | ((unsigned)oj << 16), (unsigned)ol | ((unsigned)ot << 16));
        nseg++;
      } else {
        over = true;
      }
    }
  };
  auto add = [&](int type, int si, int sj, int len) {
    if (type == ot) { ol += len; return; }   // consecutive steps of one type are contiguous
    flush();
    ot = type; oi = si; oj = sj; ol = len;
  };
  while (i > 0 && j > 0) {
    if (k == 0) {
      // iteration t of the serial walk would stand at (i - t, j - t) in state 0 and consult cell (i-t-1, j-t-1)
      const int ci = i - sl, cj = j - sl;
      int c = 0;
      bool clean = false, mt = false;
      if (Tie::kTrack && ci > 0 && cj > 0) mt = tie.match(ci, cj);
      if (ci > 1 && cj > 1) {
        c = cell(ci - 1, cj - 1);
        clean = c >= 0 && (c & 0x33) == 0;     // stays in state 0, no equal integers involved
      }
      const unsigned dirty = (__ballot_sync(smask, !clean) >> base) & ((NL == 32) ? 0xFFFFFFFFu : ((1u << NL) - 1u));
      const unsigned mbits = Tie::kTrack ? (__ballot_sync(smask, mt) >> base) : 0u;
      const int ts = dirty ? __ffs((int)dirty) - 1 : NL;   // iterations 0 .. ts-1 are plain diagonal steps
      if (Tie::kTrack) {
        const int nm = __popc(mbits & ((ts >= 32) ? 0xFFFFFFFFu : ((1u << ts) - 1u)));
        if (need_pure && nm != ts) { PM_WHY(21); return PM_WALK_TIE; }   // a mismatch below a conditional certificate
        r36 -= 36 * nm - 12 * (ts - nm);
      }
      if (ts > 0) add(0, i, j, ts);
      i -= ts;
      j -= ts;
      if (ts < NL) {  // the iteration at (i, j): its cell is the one lane ts looked at
        const int cc = __shfl_sync(smask, c, base + ts);
        if (Tie::kTrack) {
          if (need_pure && !((mbits >> ts) & 1u)) { PM_WHY(21); return PM_WALK_TIE; }
          r36 -= ((mbits >> ts) & 1u) ? 36 : -12;   // value of M[i-1][j-1]
        }
        int pk = 0;
        if (i > 1 && j > 1) {
          if (cc < 0) { PM_WHY(16); return PM_WALK_OOB; }
          if ((cc & 3) == 3) { PM_WHY(17); return PM_WALK_TIE; }
          if (Tie::kTrack && (cc & 48)) {
            const int t5 = ((cc & 3) == 2) ? 4 : ((1 << (cc & 3)) | ((cc & 16) ? 3 : 0) | ((cc & 32) ? 4 : 0));  // top_set
            if (t5 & (t5 - 1)) {
              int ok = 1;
              if (sl == 0) ok = tie.resolve(cell, i - 1, j - 1, r36);
              ok = __shfl_sync(smask, ok, base);
              if (!ok) { PM_WHY(18); return PM_WALK_TIE; }
              if (ok == 2) need_pure = true;
            }
          }
          pk = cc & 3;
        }
        add(0, i, j, 1);
        i--;
        j--;
        k = pk;
      }
    } else if (k == 2) {
      const int pj = j - 1;
      int pk = 0;
      if (pj > 0) {
        const int c = cell(i, pj);
        if (c < 0) { PM_WHY(16); return PM_WALK_OOB; }
        if ((c & 3) == 3 || (c & 128)) { PM_WHY(19); return PM_WALK_TIE; }   // the X2 comparison met equal integers
        pk = (c & 8) ? 2 : 0;
      }
      if (Tie::kTrack) {
        if (need_pure && pk == 2) { PM_WHY(22); return PM_WALK_TIE; }   // a gap extension (- 1/36) rounds
        r36 += pk == 2 ? 1 : 72;
      }
      add(2, i, j, 1);
      j = pj;
      k = pk;
    } else {
      const int pi = i - 1;
      int pk = 0;
      if (pi > 0) {
        const int c = cell(pi, j);
        if (c < 0) { PM_WHY(16); return PM_WALK_OOB; }
        if ((c & 3) == 3 || (c & 64)) { PM_WHY(20); return PM_WALK_TIE; }    // the X1 comparison met equal integers
        pk = (c & 4) ? 1 : 0;
      }
      if (Tie::kTrack) {
        if (need_pure && pk == 1) { PM_WHY(22); return PM_WALK_TIE; }
        r36 += pk == 1 ? 1 : 72;
      }
      add(1, i, j, 1);
      i = pi;
      k = pk;
    }
  }
  // the walk ended on the row-0 border -(go + (j-1) ge): a rounded constant unless j == 1
  if (Tie::kTrack && need_pure && i == 0 && j > 1) { PM_WHY(23); return PM_WALK_TIE; }
  flush();
  *n_segs = over ? -1 : nseg;
  return PM_WALK_OK;
}

// the pileup increments of a recorded path (1846-1958), by all NL lanes of the walker
template <int NL>
__device__ void coop_apply(const uint2* segs, int nseg, unsigned smask, int sl, const char* read, int mm, int orient,
                           uint32_t wstart, const PileSink& sink, const unsigned char* qcode, int qshift) {
  __syncwarp(smask);
  int pend_n = 0, pend_j = 0, pend_i = 0;
  for (int s = 0; s < nseg; s++) {
    const uint2 sg = segs[s];
    const int si = (int)(sg.x & 0xFFFFu), sj = (int)(sg.x >> 16), len = (int)(sg.y & 0xFFFFu), type = (int)(sg.y >> 16);
    if (type == 2) {  // insertion: read columns sj-1 down to sj-len wait for the next consumed reference base (1910-1911)
      pend_n = len;
      pend_j = sj;
      pend_i = si;
      continue;
    }
    if (pend_n > 0 && sl == 0) sink_insertion_direct(sink, wstart + (uint32_t)(si - 1), read, mm, orient, pend_j - pend_n, pend_n);
    pend_n = 0;
    for (int u = sl; u < len; u += NL) {
      const size_t site = (size_t)wstart + (size_t)(si - 1 - u);
      if (type == 0) {  // 1846-1858
        const int col = code_column((qcode[sj - 1 - u] >> qshift) & 15u);
        if (col >= 0) atomicAdd(&sink.counts[site * 6 + col], 1u);
      } else {
        atomicAdd(&sink.counts[site * 6 + 4], 1u);  // 1868
      }
    }
  }
  // the walk ended inside an insertion: attached to the row it stood on (1918-1958)
  if (pend_n > 0 && sl == 0) sink_insertion_direct(sink, wstart + (uint32_t)(pend_i - 1), read, mm, orient, pend_j - pend_n, pend_n);
  __syncwarp(smask);
}

// accessor over the shared-memory band store of one group: rows x PM_BAND_LANES 64-bit words
// (half = lanes kept on each side of the centre lane, <= PM_BAND_LANES / 2; smaller values only exist to test the
// out-of-band hand-over)
template <int WD, int BITS>
struct BandCell {
  const unsigned long long* band;
  int dend, half;
  __device__ __forceinline__ int operator()(int pi, int pj) const {
    const int l = (pj - 1) / WD, slot = l - (band_center_lane<WD>(pi, dend) - half);
    if (slot < 0 || slot > 2 * half) return -1;
    return (int)((band[(pi - 1) * PM_BAND_LANES + slot] >> (BITS * ((pj - 1) % WD))) & ((1ull << BITS) - 1ull));
  }
};

// accessor over the global store: rows x G words
template <int G, int WD, int BITS>
struct FullCell {
  const unsigned long long* dirs;
  __device__ __forceinline__ int operator()(int pi, int pj) const {
    return (int)((dirs[(size_t)(pi - 1) * G + (pj - 1) / WD] >> (BITS * ((pj - 1) % WD))) & ((1ull << BITS) - 1ull));
  }
};

}  // namespace pm
