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
