// seed_chain.cuh - k-mer seed lookup + co-linear chaining, one warp per read-mate (sm_100a).
//
// Replaces initial_map (pemapper.c:1539-1690): fill_mers 1969-2003, get_mers 2158-2165, the per-segment
// gather/veto/sort loop 1594-1640 and find_matches 2189-2289 for both strands, followed by the window
// set-up of map_everything (1047-1081).  Output is the candidate list in the reference's order.
//
// Memory behaviour: the 2*nseg*49 pos_index lookups of a read are spread over the 32 lanes and issued
// PM_SEED_UNROLL at a time (two 4-byte loads each, normally one 32-byte sector), so a warp keeps
// 2*32*PM_SEED_UNROLL independent loads in flight into the 16 GiB table; the second-level gathers from `mers`
// go straight into the warp's private list scratch.
// Random 8-byte gathers from HBM run at ~42 G lookups/s on a B200 whatever the fetch size (tools/gather_probe.cu),
// and ~97 % of the 48 one-substitution neighbours of a k-mer do not occur in a small genome.  FILT = true puts a
// word-blocked Bloom filter of the occupied k-mers (built at init from pos_index, kept resident in L2 through a
// persisting access-policy window) in front of the table: only k-mers the filter cannot rule out go to HBM.  The
// filter has no false negatives, so the lists, the >= too_many_spots veto and everything downstream are unchanged.  Lists are sorted by a warp bitonic network (registers
// for <=32 entries, scratch above) and chained with binary searches instead of the reference's cursors.
#pragma once
#include "pemap_common.cuh"

#ifndef PM_SEED_UNROLL
#define PM_SEED_UNROLL 8
#endif
#ifndef PM_SEED_CTAS
#define PM_SEED_CTAS 8
#endif                          // resident CTAs per SM without the filter; the scratch is sized for 8
#ifndef PM_SEED_CTAS_FILT
#define PM_SEED_CTAS_FILT 5
#endif                          // with the filter: fewer, fatter warps (more registers, more loads in flight each)
#ifndef PM_FILTER_LDCG
#define PM_FILTER_LDCG 1
#endif

namespace pm {

struct SeedArgs {
  const uint32_t* pos_index;   // 2^32+1
  const uint32_t* mers;
  const uint32_t* cstart;      // n_contigs+1 (padded to >= 9 entries)
  const char* reads[2];        // [n][stride] each
  const int* len[2];
  int stride;
  int n_reads;                 // reads (pairs) in this chunk
  int paired;
  uint32_t* scratch;           // per warp: 2*PM_MAX_SEG*PM_SEG_CAP words
  Task* tasks;
  uint32_t* task_cursor;
  uint32_t task_cap;
  uint32_t* cand_base;         // [2*n_reads]
  uint32_t* cand_n;            // [2*n_reads]
  SeedCounters* counters;
  const uint32_t* filter;      // 2^(32 - filter_shift) words, or nullptr
  int filter_shift;
  int filter_k;                // bits set per k-mer (1..4)
  DevParams p;
};

#define PM_SLIST 4             // positions of a segment kept in shared memory (complete list when segcnt <= PM_SLIST)
#define PM_SEED_DRAIN 128      // queued k-mers looked up per drain (4 per lane in flight)
#define PM_SEED_QUEUE (PM_SEED_DRAIN + 32 * PM_SEED_UNROLL)  // per-warp queue of k-mers that passed the filter

// word index and bit mask of a k-mer in the word-blocked (32-bit blocks) Bloom filter
__host__ __device__ __forceinline__ void filter_slot(uint32_t code, int shift, int k, uint32_t* word, uint32_t* mask) {
  uint32_t x = code * 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  *word = x >> shift;
  uint32_t m = 1u << (x & 31u);
  if (k > 1) m |= 1u << ((x >> 5) & 31u);
  if (k > 2) m |= 1u << ((x >> 10) & 31u);
  if (k > 3) m |= 1u << ((x >> 15) & 31u);
  *mask = m;
}

// one thread per 4 consecutive k-mer codes: occupied k-mers (count != 0 with get_mers' 32-bit wrap, 2163) are inserted
__global__ void __launch_bounds__(256) k_filter_build(const uint32_t* pos_index, uint32_t* filter, int shift, int k) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;  // 2^30 threads
  const uint32_t w0 = (uint32_t)(t << 2);
  const uint4 v = *reinterpret_cast<const uint4*>(pos_index + w0);
  const uint32_t nxt = pos_index[(uint32_t)(w0 + 4u)];  // wraps to pos_index[0] for the last group
  const uint32_t a[5] = {v.x, v.y, v.z, v.w, nxt};
#pragma unroll
  for (int i = 0; i < 4; i++)
    if (a[i + 1] != a[i]) {
      uint32_t word, mask;
      filter_slot(w0 + (uint32_t)i, shift, k, &word, &mask);
      if ((filter[word] & mask) != mask) atomicOr(filter + word, mask);
    }
}

struct SeedWarpSmem {
  char rd[2][PM_DP_MAX];       // forward read and its reverse_transcribe (C->T converted when bisulfite)
  uint32_t kcode[2 * PM_MAX_SEG];
  uint32_t segcnt[2 * PM_MAX_SEG];
  uint32_t veto[2 * PM_MAX_SEG];
  uint32_t hit_pos[PM_MAX_HITS];
  uint16_t hit_off[PM_MAX_HITS];
  uint8_t hit_or[PM_MAX_HITS];
  // head of every segment list: on a unique genome a segment has one or two positions, and sorting / chaining them
  // out of shared memory takes the L2 round trips off the warp's critical path; longer lists live in the global scratch
  uint32_t slist[2 * PM_MAX_SEG][PM_SLIST];
};

struct SeedQueueSmem {         // FILT only
  uint32_t code[PM_SEED_QUEUE];
  uint8_t ss[PM_SEED_QUEUE];
};

// pos_index lookup + mers gather of up to 32*PM_SEED_UNROLL queued k-mers (get_mers 2158-2165, loop 1594-1612)
__device__ __forceinline__ void seed_lookup_tile(const SeedArgs& a, SeedWarpSmem& sm, uint32_t* lists, const uint32_t* qcode,
                                                 const uint8_t* qss, int n, int lane, unsigned long long& st_pos) {
  uint32_t lo[PM_SEED_UNROLL], hi[PM_SEED_UNROLL];
  int ssv[PM_SEED_UNROLL];
#pragma unroll
  for (int u = 0; u < PM_SEED_UNROLL; u++) {
    const int q = u * 32 + lane;
    ssv[u] = -1;
    lo[u] = hi[u] = 0;
    if (q < n) {
      const uint32_t code = qcode[q];
      ssv[u] = qss[q];
      if ((code & 1u) == 0u) {
        const uint2 v = __ldcg(reinterpret_cast<const uint2*>(a.pos_index + code));
        lo[u] = v.x;
        hi[u] = v.y;
      } else {
        lo[u] = __ldcg(a.pos_index + code);
        hi[u] = __ldcg(a.pos_index + (uint32_t)(code + 1u));  // which+1 wraps in 32 bits (2163)
      }
    }
  }
#pragma unroll
  for (int u = 0; u < PM_SEED_UNROLL; u++) {
    if (ssv[u] >= 0) {
      uint32_t cnt = hi[u] - lo[u];
      if (cnt >= (uint32_t)a.p.too_many_spots) {
        sm.veto[ssv[u]] = 1;  // 1602-1606: one crowded k-mer empties the whole segment
      } else if (cnt) {
        uint32_t off = atomicAdd(&sm.segcnt[ssv[u]], cnt);
        uint32_t* dst = lists + (size_t)ssv[u] * PM_SEG_CAP + off;
        const uint32_t* src = a.mers + lo[u];
        for (uint32_t t = 0; t < cnt; t++) {  // the head of a list lives in shared memory only (copied out before a long sort)
          const uint32_t v = __ldcg(src + t);
          if (off + t < PM_SLIST) sm.slist[ssv[u]][off + t] = v;
          else dst[t] = v;
        }
        st_pos += cnt;
      }
    }
  }
}

// one-substitution neighbour v (1..48) of a packed 16-mer.  fill_mers (1969-2003) walks bytes, 2-bit fields and the
// three alternative bases; the set it produces is { code ^ (d << 2f) : f = 0..15, d = 1..3 }, and the order of the
// 49 lookups of a segment does not matter (their lists are concatenated and sorted, the veto is an OR).
__device__ __forceinline__ uint32_t kmer_variant(uint32_t code, int v) {
  if (v == 0) return code;
  const int f = (v - 1) / 3, d = v - 3 * f;  // d = 1..3
  return code ^ ((uint32_t)d << (2 * f));
}

__device__ __forceinline__ void warp_sort_list(uint32_t* lst, int n, int lane) {
  if (n <= 4) {  // the usual case on a unique genome: the exact hit plus a chance neighbour or two
    if (lane == 0) {
      uint32_t a = lst[0], b = lst[1], c = n > 2 ? lst[2] : 0xFFFFFFFFu, d = n > 3 ? lst[3] : 0xFFFFFFFFu, t;
      if (a > b) { t = a; a = b; b = t; }
      if (c > d) { t = c; c = d; d = t; }
      if (a > c) { t = a; a = c; c = t; }
      if (b > d) { t = b; b = d; d = t; }
      if (b > c) { t = b; b = c; c = t; }
      lst[0] = a;
      lst[1] = b;
      if (n > 2) lst[2] = c;
      if (n > 3) lst[3] = d;
    }
    __syncwarp();
    return;
  }
  if (n <= 32) {
    uint32_t v = lane < n ? lst[lane] : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
        bool up = (lane & k) == 0, lower = (lane & j) == 0;
        v = (lower == up) ? min(v, o) : max(v, o);
      }
    if (lane < n) lst[lane] = v;
    __syncwarp();
    return;
  }
  int P = 64;
  while (P < n) P <<= 1;
  for (int i = n + lane; i < P; i += 32) lst[i] = 0xFFFFFFFFu;
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        int x = i ^ j;
        if (x > i) {
          uint32_t a = lst[i], b = lst[x];
          bool up = (i & k) == 0;
          if ((a > b) == up) {
            lst[i] = b;
            lst[x] = a;
          }
        }
      }
      __syncwarp();
    }
}

// does the sorted list hold a position p with lo <= p <= hi ?
__device__ __forceinline__ bool list_has_in_range(const uint32_t* lst, int n, long long lo, long long hi) {
  if (hi < 0) return false;
  uint32_t ulo = lo < 0 ? 0u : (uint32_t)lo;
  int a = 0, b = n;
  while (a < b) {
    int m = (a + b) >> 1;
    if (lst[m] < ulo) a = m + 1; else b = m;
  }
  return a < n && (long long)lst[a] <= hi;
}

// FK = bits per k-mer of the Bloom filter (1..3), 0 = no filter
template <int WARPS, int FK>
__global__ void __launch_bounds__(WARPS * 32, FK > 0 ? PM_SEED_CTAS_FILT : PM_SEED_CTAS) k_seed_chain(SeedArgs a) {
  constexpr bool FILT = FK > 0;
  __shared__ SeedWarpSmem smem[WARPS];
  __shared__ SeedQueueSmem qsmem[FILT ? WARPS : 1];
  __shared__ uint32_t s_xm[64];   // FILT: xor mask of variant v (kmer_variant), so that the lookup loop has no divisions
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (FILT) {
    if (threadIdx.x < 64) s_xm[threadIdx.x] = (threadIdx.x > 0 && threadIdx.x < PM_KV) ? kmer_variant(0u, (int)threadIdx.x) : 0u;
    __syncthreads();
  }
  SeedWarpSmem& sm = smem[warp];
  const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  uint32_t* lists = a.scratch + (size_t)gw * (2 * PM_MAX_SEG * PM_SEG_CAP);
  const int n_work = a.paired ? 2 * a.n_reads : a.n_reads;
  unsigned long long st_lookups = 0, st_pos = 0, st_cand = 0, st_cells = 0;

  for (int w = gw; w < n_work; w += nw) {
    const int r = a.paired ? (w >> 1) : w, mate = a.paired ? (w & 1) : 0;
    const uint32_t rm = 2u * (uint32_t)r + (uint32_t)mate;
    const int len = a.len[mate][r];
    const char* read = a.reads[mate] + (size_t)r * a.stride;
    int tot = 0;
    {  // the next read-mate of this warp: its row and length are cold in HBM; start fetching them now (10 % of the
       // warp's time was the wait for the first load of the read)
      const int wn = w + nw;
      if (wn < n_work && lane < 4) {
        const int rn = a.paired ? (wn >> 1) : wn, mn = a.paired ? (wn & 1) : 0;
        const char* nxt = a.reads[mn] + (size_t)rn * a.stride;
        if (lane < 3) {
          if (lane * 128 < a.stride) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + lane * 128));
        } else {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a.len[mn] + rn));
        }
      }
    }

    bool ok = (len >= 16 && len < PM_DP_MAX - 21);
    // N filter (1552-1559) + forward / reverse-transcribed copies (1019-1021, 1561-1570)
    int n_count = 0;
    if (ok) {
      for (int i = lane; i < len; i += 32) {
        char ch = read[i];
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
        int strand = ss >= nseg, s = strand ? ss - nseg : ss;
        int off = (s < total_cuts) ? 16 * s : len - 16;
        const char* q = sm.rd[strand] + off;
        uint32_t code = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) code = (code << 2) | base_code(q[i]);
        sm.kcode[ss] = code;
        sm.segcnt[ss] = 0;
        sm.veto[ss] = 0;
      }
      __syncwarp();

      // ---- lookups (get_mers 2158-2165) and gathers (1594-1612, 1619-1637)
      const int L = 2 * nseg * PM_KV;
      if (FILT) {
        SeedQueueSmem& qs = qsmem[warp];
        int qn = 0;  // warp-uniform queue length
        // lookup q = 32 * step + lane is variant vq of (strand, segment) sq: both advance without a division
        int sq = 0, vq = lane;
        const int nss = 2 * nseg;
        for (int q0 = 0; q0 < L; q0 += 32 * PM_SEED_UNROLL) {
          uint32_t codes[PM_SEED_UNROLL], fw[PM_SEED_UNROLL], fm[PM_SEED_UNROLL];
          int ssu[PM_SEED_UNROLL];
#pragma unroll
          for (int u = 0; u < PM_SEED_UNROLL; u++) {
            fw[u] = 0;
            fm[u] = 1;
            codes[u] = 0;
            ssu[u] = sq;
            if (sq < nss) {
              codes[u] = sm.kcode[sq] ^ s_xm[vq];
              uint32_t x = codes[u] * 0x9E3779B1u;   // filter_slot with the number of bits fixed at compile time
              x ^= x >> 15;
              x *= 0x85EBCA77u;
              x ^= x >> 13;
              uint32_t m = 1u << (x & 31u);
              if (FK > 1) m |= 1u << ((x >> 5) & 31u);
              if (FK > 2) m |= 1u << ((x >> 10) & 31u);
              fm[u] = m;
              fw[u] = PM_FILTER_LDCG ? __ldcg(a.filter + (x >> a.filter_shift)) : a.filter[x >> a.filter_shift];  // L2-resident (persisting window)
            }
            vq += 32;
            if (vq >= PM_KV) {
              vq -= PM_KV;
              sq++;
            }
          }
#pragma unroll
          for (int u = 0; u < PM_SEED_UNROLL; u++) {
            const bool pass = (fw[u] & fm[u]) == fm[u];  // lanes past L hold fw = 0, fm = 1
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
            if (pass) {
              const int at = qn + __popc(bal & ((1u << lane) - 1u));
              qs.code[at] = codes[u];
              qs.ss[at] = (uint8_t)ssu[u];
            }
            qn += __popc(bal);
          }
          __syncwarp();
          while (qn >= PM_SEED_DRAIN) {  // drain from the tail
            qn -= PM_SEED_DRAIN;
            seed_lookup_tile(a, sm, lists, qs.code + qn, qs.ss + qn, PM_SEED_DRAIN, lane, st_pos);
            __syncwarp();
          }
        }
        if (qn > 0) seed_lookup_tile(a, sm, lists, qs.code, qs.ss, qn, lane, st_pos);
      } else {
        for (int q0 = 0; q0 < L; q0 += 32 * PM_SEED_UNROLL) {
          uint32_t lo[PM_SEED_UNROLL], hi[PM_SEED_UNROLL];
          int ssv[PM_SEED_UNROLL];
#pragma unroll
          for (int u = 0; u < PM_SEED_UNROLL; u++) {
            int q = q0 + u * 32 + lane;
            ssv[u] = -1;
            lo[u] = hi[u] = 0;
            if (q < L) {
              int ss = q / PM_KV, v = q - ss * PM_KV;
              uint32_t code = kmer_variant(sm.kcode[ss], v);
              ssv[u] = ss;
              // L1-bypassing loads (the line would never be reused)
              if ((code & 1u) == 0u) {
                const uint2 v = __ldcg(reinterpret_cast<const uint2*>(a.pos_index + code));
                lo[u] = v.x;
                hi[u] = v.y;
              } else {
                lo[u] = __ldcg(a.pos_index + code);
                hi[u] = __ldcg(a.pos_index + (uint32_t)(code + 1u));  // which+1 wraps in 32 bits (2163)
              }
            }
          }
#pragma unroll
          for (int u = 0; u < PM_SEED_UNROLL; u++) {
            if (ssv[u] >= 0) {
              uint32_t cnt = hi[u] - lo[u];
              if (cnt >= (uint32_t)a.p.too_many_spots) {
                sm.veto[ssv[u]] = 1;  // 1602-1606: one crowded k-mer empties the whole segment
              } else if (cnt) {
                uint32_t off = atomicAdd(&sm.segcnt[ssv[u]], cnt);
                uint32_t* dst = lists + (size_t)ssv[u] * PM_SEG_CAP + off;
                const uint32_t* src = a.mers + lo[u];
                for (uint32_t t = 0; t < cnt; t++) {
                  const uint32_t v = __ldcg(src + t);
                  if (off + t < PM_SLIST) sm.slist[ssv[u]][off + t] = v;
                  else dst[t] = v;
                }
                st_pos += cnt;
              }
            }
          }
        }
      }
      st_lookups += (lane == 0) ? (unsigned long long)L : 0ull;
      __syncwarp();
      for (int ss = lane; ss < 2 * nseg; ss += 32)
        if (sm.veto[ss]) sm.segcnt[ss] = 0;
      __syncwarp();
      // ---- ascending sort of every segment list (1613-1614, 1638-1639)
      for (int ss = 0; ss < 2 * nseg; ss++) {
        int n = (int)sm.segcnt[ss];
        if (n > PM_SLIST) {  // a long list is sorted and searched in the global scratch: its head joins it there
          if (lane < PM_SLIST) lists[(size_t)ss * PM_SEG_CAP + lane] = sm.slist[ss][lane];
          __syncwarp();
        }
        if (n > 1) warp_sort_list(n <= PM_SLIST ? sm.slist[ss] : lists + (size_t)ss * PM_SEG_CAP, n, lane);
      }
      __syncwarp();

      // ---- find_matches (2189-2289), forward strand then reverse strand (1656-1660)
      int min_match = total_cuts > 1 ? total_cuts : 1;  // 1642-1645
      if (total_cuts > 4) min_match = (4 * total_cuts) / 5;
      if (min_match > 4) min_match = 4;
      const int max_depth = total_cuts;
      for (int strand = 0; strand < 2; strand++) {
        if (strand == 1 && tot >= a.p.max_hits) break;  // 1658
        const uint32_t* cnts = sm.segcnt + strand * nseg;
        const uint32_t* slists = lists + (size_t)strand * nseg * PM_SEG_CAP;
        const uint32_t(*shead)[PM_SLIST] = sm.slist + strand * nseg;
        uint32_t ms = 10000;
        if (lane < nseg) ms = cnts[lane];
        ms = __reduce_min_sync(0xFFFFFFFFu, ms);
        if (ms > (uint32_t)a.p.max_hits) {  // 2203-2207: also wipes the other strand's hits
          tot = 0;
          continue;
        }
        bool done = false;
        for (int loop = 0; !done && loop <= 1 + max_depth - min_match; loop++) {
          const int off_loop = (loop < total_cuts) ? 16 * loop : len - 16;
          const int na = (int)cnts[loop];
          const uint32_t* anchors = na <= PM_SLIST ? shead[loop] : slists + (size_t)loop * PM_SEG_CAP;
          const int mo = (a.p.idepth - 4 > 2) ? a.p.idepth - 4 : 2;  // max_off (2196)
          // Two lane mappings with the same result.  Few anchors (the usual case: the exact hit and a chance
          // neighbour): one anchor at a time, the lanes check the later segments in parallel.  Many anchors
          // (repeats): 32 anchors at a time, each lane walks the later segments itself.
          const bool by_segment = na <= PM_SLIST;
          for (int i0 = 0; i0 < na && !done; i0 += by_segment ? 1 : 32) {
            unsigned mask;
            int found = 1;
            uint32_t av;
            if (by_segment) {
              av = anchors[i0];
              const int j = loop + 1 + lane;
              bool hit = false;
              if (j <= max_depth) {
                const int nj = (int)cnts[j];
                if (nj) {
                  const long long c = (long long)av + (((j < total_cuts) ? 16 * j : len - 16) - off_loop);
                  hit = list_has_in_range(nj <= PM_SLIST ? shead[j] : slists + (size_t)j * PM_SEG_CAP, nj, c - (mo - 1),
                                          c + (mo - 1));
                }
              }
              found = 1 + __popc(__ballot_sync(0xFFFFFFFFu, hit));
              mask = found >= min_match ? 1u : 0u;  // lane 0 stands for the anchor
            } else {
              const int i = i0 + lane;
              const bool valid = i < na;
              av = valid ? anchors[i] : 0u;
              if (valid)
                for (int j = loop + 1; j <= max_depth; j++) {
                  const int nj = (int)cnts[j];
                  if (nj == 0) continue;
                  const int d = ((j < total_cuts) ? 16 * j : len - 16) - off_loop;
                  // |(anchor - p) - (offsets[loop] - offsets[j])| < max_off = 12  (2244)
                  const long long c = (long long)av + d;
                  found += list_has_in_range(nj <= PM_SLIST ? shead[j] : slists + (size_t)j * PM_SEG_CAP, nj, c - (mo - 1),
                                             c + (mo - 1)) ? 1 : 0;
                }
              mask = __ballot_sync(0xFFFFFFFFu, valid && found >= min_match);
            }
            while (mask) {
              const int l = __ffs(mask) - 1;
              mask &= mask - 1;
              const int f = __shfl_sync(0xFFFFFFFFu, found, l);
              const uint32_t apos = __shfl_sync(0xFFFFFFFFu, av, l);
              if (f > min_match) {  // 2251-2260
                min_match = f;
                tot = 0;
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
                  break;
                }
              }
            }
          }
        }
      }
    }

    // ---- candidate windows (1047-1081) -> alignment tasks
    uint32_t base = 0;
    if (lane == 0 && tot > 0) base = atomicAdd(a.task_cursor, (uint32_t)tot);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (tot > 0 && base + (uint32_t)tot > a.task_cap) tot = 0;  // cannot happen: cap = 200 * work items
    for (int c = lane; c < tot; c += 32) {
      long long t = (long long)sm.hit_pos[c] - (long long)sm.hit_off[c];  // 1664-1669
      uint32_t spot = (uint32_t)(t > 0 ? t : 0);
      Task tk;
      tk.rm = rm | ((uint32_t)sm.hit_or[c] << 31);
      tk.spot = spot;
      candidate_window(a.cstart, a.p.n_contigs, spot, len, a.p.misalign_slop, &tk.wstart, &tk.blen);
      a.tasks[base + c] = tk;
      if (tk.blen > 0) st_cells += (unsigned long long)tk.blen * (unsigned long long)len;
    }
    if (lane == 0) {
      a.cand_base[rm] = base;
      a.cand_n[rm] = (uint32_t)tot;
      st_cand += (unsigned long long)tot;
    }
    __syncwarp();
  }
  // statistics: one atomic per counter per warp
  st_pos = __reduce_add_sync(0xFFFFFFFFu, (unsigned)st_pos);  // per-warp totals fit 32 bits per chunk
  unsigned long long cells_lo = st_cells;
  for (int o = 16; o > 0; o >>= 1) cells_lo += __shfl_xor_sync(0xFFFFFFFFu, cells_lo, o);
  if (lane == 0) {
    atomicAdd(&a.counters->lookups, st_lookups);
    atomicAdd(&a.counters->mer_positions, st_pos);
    atomicAdd(&a.counters->candidates, st_cand);
    atomicAdd(&a.counters->sw_cells, cells_lo);
  }
}

}  // namespace pm
