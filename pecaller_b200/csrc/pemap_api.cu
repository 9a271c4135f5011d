// pemap_api.cu - C-ABI (include/pemap.h) over the sm_100a kernels.  Host orchestration only: buffers, streams,
// launches, copies.  There is no CPU implementation of any stage in this library.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pemap.h"
#include "pemap_common.cuh"
#include "seed_chain.cuh"
#include "seed_rbi.cuh"
#include "select_finish.cuh"
#include "sw_wavefront.cuh"
#include "sw_int16.cuh"
#include "trace_int.cuh"

#define PEMAP_VERSION "pemap-b200 0.1 (sm_100a)"

namespace {

constexpr int kSeedWarps = 4;
constexpr int kRbiWarps = 8;     // k_seed_rbi: 8 warps per CTA, 9.3 KB of shared memory per warp -> 3 CTAs per SM
constexpr int kRbiMidWarps = 4;  // second pass: 27 KB per warp
constexpr int kRbiBigWarps = 4;
constexpr int kMaxDev = 16;  // per-device caches of launch configurations
constexpr uint64_t kInsCapDefault = 256ull << 20;

struct Timer {
  cudaEvent_t a, b;
};

}  // namespace

struct pemap_ctx {
  int device = 0;
  int sm_count = 148;
  pemap_params params;
  pm::DevParams dp;
  std::string err;
  int keep = 0;

  // index + genome in HBM
  uint32_t* d_pos_index = nullptr;
  uint32_t* d_mers = nullptr;
  uint64_t n_mers = 0;
  char* d_genome = nullptr;
  uint64_t genome_size = 0;
  uint32_t* d_cstart = nullptr;
  int n_contigs = 0;
  double* d_border = nullptr;
  // device-private rotated bucket index (seed_rbi.cuh): what the seed kernel reads; pos_index / mers are only kept
  // beside it while the genome is small (tests of the index builder, the legacy seed kernel as a cross-check)
  uint4* d_rbi_data[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t* d_rbi_dir[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t rbi_bytes = 0;
  int seed_legacy = 0;             // PEMAP_SEED=legacy: k_seed_chain over pos_index / mers
  int index_only = 0;              // PEMAP_INDEX_ONLY=1: the handle only builds and hands out the index
  uint32_t* d_big_list = nullptr;  // read-mates whose strand lists did not fit the first seed pass / the second
  uint32_t* d_big_list2 = nullptr;
  unsigned char* d_big_scratch = nullptr;
  int big_grid = 0;
  uint32_t* d_filter = nullptr;  // word-blocked Bloom filter of the occupied k-mers (seed_chain.cuh), or null
  int filter_shift = 0, filter_k = 0;
  size_t filter_bytes = 0;
  uint32_t* d_counts = nullptr;

  unsigned char* d_ins = nullptr;
  unsigned long long* d_ins_cursor = nullptr;
  uint64_t ins_cap = kInsCapDefault;

  // chunk buffers
  int chunk = 0;       // reads (pairs) per chunk
  int stride_cap = 0;  // bytes per read row in the device buffers
  struct Slot {  // double-buffered chunk I/O: chunk c+1 is copied in while chunk c computes
    char* d_reads[2] = {nullptr, nullptr};
    int* d_len[2] = {nullptr, nullptr};
    char* h_reads[2] = {nullptr, nullptr};  // pinned staging for pageable / pointer-array callers
    int* h_len[2] = {nullptr, nullptr};
    uint32_t *d_m1 = nullptr, *d_m2 = nullptr, *h_m1 = nullptr, *h_m2 = nullptr;
    int *d_type = nullptr, *h_type = nullptr;
    cudaEvent_t ev[7] = {};                 // stage boundaries of the chunk on the compute stream
    cudaEvent_t ev_h2d = nullptr, ev_d2h = nullptr;
    bool pending = false;
    int n = 0, first = 0;
    bool paired = false, direct = false;
  } slots[2];
  // 2-bit packed input (pemap_map_batch_packed): device copies of the packed rows, per slot and mate, allocated at the
  // first packed batch; cur_packed describes the chunk being submitted to run_chunk
  unsigned char* d_packed[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  size_t packed_cap = 0;
  struct {
    const unsigned char* rows[2] = {nullptr, nullptr};
    int stride = 0, code_words = 0, mask_words = 0;
  } cur_packed;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaStream_t s_aux = nullptr;            // the pure-diagonal pileup runs beside the integer traceback
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int diag_overlap = 1;                    // PEMAP_DIAG_OVERLAP=0: k_apply_diag on the compute stream
  pm::Task* d_tasks = nullptr;
  pm::TaskResult* d_results = nullptr;
  uint32_t task_cap = 0;
  uint32_t* d_cursors = nullptr;  // [0] tasks, [1] winners, [2] replay reads, [3] replay tasks, [4]/[5] min/max len,
                                  // [6] pure-diagonal winners, [7] winners handed to the fp64 traceback
  pm::Winner* d_diag_winners = nullptr;
  pm::Winner* d_exact_winners = nullptr;
  pm::Winner* d_oob_winners = nullptr;  // [8] winners whose walk left the shared-memory band
  pm::ITaskResult* d_ires = nullptr;
  uint32_t* d_replay_reads = nullptr;
  pm::Winner* d_replay_tasks = nullptr;
  int band_half = PM_BAND_LANES / 2;  // PEMAP_BAND_HALF=0/1 narrows the traceback band (tests of the hand-over path)
  uint32_t* d_sw_list = nullptr;  // tasks left for the DP scoring kernel by k_diag_certify
  int certify = 1;                // PEMAP_CERTIFY=0: score every candidate with the DP (cross-check)
  void* d_flagq = nullptr;      // packed decision flags between k_trace_dp16 and k_trace_walk16
  size_t flagq_bytes = 0;
  void* d_walk_meta = nullptr;   // per gapped winner: what k_trace_walk16 needs (written by k_trace_dp16)
  void* d_pair_codes = nullptr;  // one-hot base codes of every winner pair (window rows, read columns)
  size_t pair_codes_bytes = 0;
  int exact = 0;  // 1: fp64 kernels for everything (PEMAP_EXACT=1, PEMAP_KEEP_DETAIL, match_bonus != 1)
  uint32_t* d_cand_base = nullptr;
  uint32_t* d_cand_n = nullptr;
  pm::Winner* d_winners = nullptr;
  int32_t* d_det_best = nullptr;
  int32_t* d_det_orient = nullptr;
  double* d_det_score = nullptr;
  uint32_t* d_seed_scratch = nullptr;
  int seed_blocks = 0;
  unsigned long long* d_dirs = nullptr;
  char* d_pend = nullptr;
  int sw_blocks = 0;
  pm::SeedCounters* d_counters = nullptr;

  cudaStream_t stream = nullptr;

  // retained from the last batch
  std::vector<pemap_detail> detail;
  std::vector<uint32_t> cand_base_h, cand_n_h;  // per read-mate, offsets into cand_spot_h
  std::vector<uint32_t> cand_spot_h;
  std::vector<int8_t> cand_orient_h;

  // finish buffers
  std::vector<pemap_record> records;
  std::vector<pemap_insertion> ins;
  std::vector<char> ins_pool;
  // streaming finish: the counters are compacted window by window through two bounded device buffers and two pinned
  // host buffers (pemap_finish_stream); allocated at the first call
  uint64_t fin_sites = 0;                       // sites per window (multiple of the compaction tile)
  pm::PileRecord* d_fin_rec[2] = {nullptr, nullptr};
  pemap_record* h_fin_rec[2] = {nullptr, nullptr};
  unsigned long long *d_fin_cnt = nullptr, *d_fin_off = nullptr, *h_fin_total = nullptr;
  void* d_fin_tmp = nullptr;
  size_t fin_tmp_bytes = 0;
  cudaEvent_t ev_fin[2] = {nullptr, nullptr};
  // counter arrays of the other ranks' GPUs opened through CUDA IPC (pemap_reduce_scatter_ipc), kept until destroy
  std::vector<void*> ipc_open;
  std::vector<std::string> ipc_keys;
  // insertion records drained from the device append buffer so far (same {pos, len, chars} layout)
  std::vector<unsigned char> ins_raw;
  unsigned long long* h_ins_used = nullptr;     // pinned: cursor value after each chunk, one per slot
  bool want_drain = false;

  pemap_stats stats;
  struct pemap_lanes* lanes_ = nullptr;
};

namespace {

__global__ void k_len_range(const int* l1, const int* l2, int n, unsigned* out) {
  unsigned lo = 0xFFFFFFFFu, hi = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned a = (unsigned)l1[i];
    lo = min(lo, a);
    hi = max(hi, a);
    if (l2) {
      unsigned b = (unsigned)l2[i];
      lo = min(lo, b);
      hi = max(hi, b);
    }
  }
  lo = __reduce_min_sync(0xFFFFFFFFu, lo);
  hi = __reduce_max_sync(0xFFFFFFFFu, hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out, lo);
    atomicMax(out + 1, hi);
  }
}

int fail(pemap_ctx* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(h, e_ == cudaErrorMemoryAllocation ? PEMAP_ERR_NOMEM : PEMAP_ERR_CUDA,                  \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                    \
  } while (0)

}  // namespace

// The per-chunk working set exists twice ("lanes"): chunk c runs on lane c & 1, each lane on its own stream, so that the
// seed stage of one chunk (bound by HBM) shares the SMs with the scoring and traceback kernels of the previous one
// (bound by integer issue).  The context's plain fields are the CURRENT lane's; use_lane() swaps them.
#define PM_LANE_FIELDS(X)                                                                                              \
  X(stream) X(s_aux) X(ev_fork) X(ev_join) X(d_tasks) X(d_results) X(d_cursors) X(d_diag_winners) X(d_exact_winners)    \
  X(d_oob_winners) X(d_ires) X(d_replay_reads) X(d_replay_tasks) X(d_sw_list) X(d_flagq) X(flagq_bytes) X(d_walk_meta)  \
  X(d_pair_codes) X(pair_codes_bytes) X(d_cand_base) X(d_cand_n) X(d_winners) X(d_det_best) X(d_det_orient)             \
  X(d_det_score) X(d_seed_scratch) X(d_dirs) X(d_pend) X(d_big_list) X(d_big_list2) X(d_big_scratch)

struct pemap_lane {
#define X(f) decltype(pemap_ctx::f) f{};
  PM_LANE_FIELDS(X)
#undef X
};

struct pemap_lanes {  // kept beside the context (pemap_ctx::lanes_)
  pemap_lane saved[2];
  int cur = 0, n = 1;
  cudaEvent_t ev_main = nullptr, ev_side = nullptr;
};

namespace {

void use_lane(pemap_ctx* h, int k) {
  pemap_lanes& L = *h->lanes_;
  if (k == L.cur) return;
#define X(f) L.saved[L.cur].f = h->f; h->f = L.saved[k].f;
  PM_LANE_FIELDS(X)
#undef X
  L.cur = k;
}

// everything either lane has been given is done
int sync_lanes(pemap_ctx* h) {
  pemap_lanes& L = *h->lanes_;
  CK(cudaStreamSynchronize(h->stream));
  if (L.n > 1) CK(cudaStreamSynchronize(L.saved[L.cur ^ 1].stream));
  return PEMAP_OK;
}

// the side lane starts after what the main stream holds so far (resets, parameter uploads)
int fork_lanes(pemap_ctx* h) {
  pemap_lanes& L = *h->lanes_;
  if (L.n < 2) return PEMAP_OK;
  use_lane(h, 0);
  CK(cudaEventRecord(L.ev_main, h->stream));
  CK(cudaStreamWaitEvent(L.saved[1].stream, L.ev_main, 0));
  return PEMAP_OK;
}

// the main stream (the one pemap_stream hands out) continues after both lanes
int join_lanes(pemap_ctx* h) {
  pemap_lanes& L = *h->lanes_;
  use_lane(h, 0);
  if (L.n < 2) return PEMAP_OK;
  CK(cudaEventRecord(L.ev_side, L.saved[1].stream));
  CK(cudaStreamWaitEvent(h->stream, L.ev_side, 0));
  return PEMAP_OK;
}

}  // namespace

namespace {

void fill_dev_params(pemap_ctx* h) {
  const pemap_params& p = h->params;
  pm::DevParams& d = h->dp;
  const double mb = p.match_bonus;
  d.match = mb;
  d.mism = -1.0 / ((double)3.0 * mb);  // pemapper.c:2011
  d.go = 2.0 * mb;                     // pemapper.c:2039
  d.ge = mb / 36.0;                    // pemapper.c:2040
  d.min_align = p.min_align;
  d.match_bonus = mb;
  d.idepth = p.idepth;
  d.max_hits = p.max_hits;
  d.too_many_spots = p.too_many_spots;
  d.is_bisulfite = p.is_bisulfite;
  d.pair_flag = p.pair_flag;
  d.min_dist = p.min_dist;
  d.max_dist = p.max_dist;
  d.misalign_slop = p.misalign_slop;
  d.n_contigs = h->n_contigs;
  d.genome_size_lo = (uint32_t)h->genome_size;
}

int check_params(pemap_ctx* h, const pemap_params* p) {
  if (!p) return fail(h, PEMAP_ERR_ARG, "params is NULL");
  if (p->idepth != 16) return fail(h, PEMAP_ERR_UNSUPPORTED, "idepth must be 16 (index_genome_whole.c:149)");
  if (p->max_hits < 1 || p->max_hits > PM_MAX_HITS) return fail(h, PEMAP_ERR_UNSUPPORTED, "max_hits must be 1..200");
  if (p->too_many_spots < 1 || p->too_many_spots > 100)
    return fail(h, PEMAP_ERR_UNSUPPORTED, "too_many_spots must be 1..100");
  if (p->misalign_slop != 10) return fail(h, PEMAP_ERR_UNSUPPORTED, "misalign_slop must be 10 (MISALIGN_SLOP)");
  if (!(p->match_bonus > 0)) return fail(h, PEMAP_ERR_ARG, "match_bonus must be > 0");
  return PEMAP_OK;
}

int upload_border(pemap_ctx* h) {
  // S*[0][j] = -(gap_open + (j-1)*gap_extend), init_penalty_matrices pemapper.c:2073-2081, computed on the host
  // with the reference's own expression (no FMA: this file is built with -ffp-contract=off on the host side).
  std::vector<double> b(PM_DP_MAX + 1);
  const volatile double go = h->dp.go, ge = h->dp.ge;
  b[0] = 0.0;
  for (int j = 1; j <= PM_DP_MAX; j++) {
    volatile double prod = (double)(j - 1) * ge;
    volatile double sum = go + prod;
    b[j] = -sum;
  }
  if (!h->d_border) CK(cudaMalloc(&h->d_border, sizeof(double) * (PM_DP_MAX + 1)));
  CK(cudaMemcpy(h->d_border, b.data(), sizeof(double) * (PM_DP_MAX + 1), cudaMemcpyHostToDevice));
  return PEMAP_OK;
}

int open_device(pemap_ctx* h, int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(h, PEMAP_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0"));
  if (device < 0 || device >= n) return fail(h, PEMAP_ERR_ARG, "device ordinal out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(h, PEMAP_ERR_CUDA, "device is not sm_100 (this library has only sm_100a code)");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  // legacy seed kernel: random 8-byte gathers into the 16 GiB table want sector-sized (32 B) DRAM fetches; the rotated
  // bucket index is read as contiguous ~1 KB buckets, 128-byte pieces at a time: whole lines (PEMAP_L2_FETCH overrides)
  {
    size_t gran = 128;
    if (const char* s = getenv("PEMAP_SEED")) gran = strcmp(s, "legacy") == 0 ? 32 : 128;
    if (const char* s = getenv("PEMAP_L2_FETCH")) gran = (size_t)atoi(s);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (const char* s = getenv("PEMAP_DIAG_OVERLAP")) h->diag_overlap = atoi(s) != 0;
  for (auto& sl : h->slots) {
    for (auto& ev : sl.ev) CK(cudaEventCreate(&ev));
    CK(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
  }
  return PEMAP_OK;
}

// the working set of one lane (see pemap_lanes): everything a chunk's kernels write between its H2D and its D2H
int alloc_lane_scratch(pemap_ctx* h) {
  const size_t n = (size_t)h->chunk;
  CK(cudaMalloc(&h->d_tasks, (size_t)h->task_cap * sizeof(pm::Task)));
  CK(cudaMalloc(&h->d_results, (size_t)h->task_cap * sizeof(pm::TaskResult)));
  CK(cudaMalloc(&h->d_cursors, 128));
  CK(cudaMalloc(&h->d_sw_list, (size_t)h->task_cap * 4));
  CK(cudaMalloc(&h->d_ires, (size_t)h->task_cap * sizeof(pm::ITaskResult)));
  CK(cudaMalloc(&h->d_replay_reads, n * 4));
  CK(cudaMalloc(&h->d_replay_tasks, (size_t)h->task_cap * sizeof(pm::Winner)));
  CK(cudaMalloc(&h->d_cand_base, 2 * n * 4));
  CK(cudaMalloc(&h->d_cand_n, 2 * n * 4));
  CK(cudaMemset(h->d_cand_n, 0, 2 * n * 4));
  CK(cudaMemset(h->d_cand_base, 0, 2 * n * 4));
  CK(cudaMalloc(&h->d_winners, 2 * n * sizeof(pm::Winner)));
  CK(cudaMalloc(&h->d_diag_winners, 2 * n * sizeof(pm::Winner)));
  CK(cudaMalloc(&h->d_exact_winners, 2 * n * sizeof(pm::Winner)));
  CK(cudaMalloc(&h->d_oob_winners, 2 * n * sizeof(pm::Winner)));
  CK(cudaMalloc(&h->d_det_best, 2 * n * 4));
  CK(cudaMalloc(&h->d_det_orient, 2 * n * 4));
  CK(cudaMalloc(&h->d_det_score, 2 * n * 8));
  if (h->seed_legacy) {
    CK(cudaMalloc(&h->d_seed_scratch, (size_t)h->seed_blocks * kSeedWarps * 2 * PM_MAX_SEG * PM_SEG_CAP * 4));
  } else {  // second seed pass (strand lists that do not fit shared memory): one CTA per SM, per-warp lists in HBM
    CK(cudaMalloc(&h->d_big_scratch, (size_t)h->big_grid * kRbiBigWarps * PM_RBI_BIG_BYTES));
    CK(cudaMalloc(&h->d_big_list, 2 * n * 4));
    CK(cudaMalloc(&h->d_big_list2, 2 * n * 4));
  }
  const size_t max_groups = (size_t)h->sw_blocks * (128 / 16);
  // trace scratch: groups * G == sw_blocks * 128 lanes for every instantiation, PM_DP_MAX rows of one word per lane
  CK(cudaMalloc(&h->d_dirs, (size_t)h->sw_blocks * 128 * PM_DP_MAX * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->d_pend, 2 * max_groups * PM_DP_MAX));
  return PEMAP_OK;
}

int alloc_chunk_buffers(pemap_ctx* h) {
  int chunk = 1 << 19;  // reads (pairs) per pass of the kernel chain: larger chunks shorten the persistent kernels' tails
  h->lanes_ = new pemap_lanes();
  h->lanes_->n = 1;  // a second lane measured no gain (profiles/README_r02.md): opt-in
  if (const char* s = getenv("PEMAP_LANES")) h->lanes_->n = atoi(s) >= 2 ? 2 : 1;
  {  // two lanes of 512 Ki pairs take ~2 x 16.5 GB; beside a human-sized genome's counters halve the chunk instead
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && h->lanes_->n > 1 && free_b < (44ull << 30)) chunk = 1 << 18;
    cudaGetLastError();
  }
  if (const char* s = getenv("PEMAP_CHUNK")) chunk = std::max(1024, atoi(s));
  h->chunk = chunk;
  h->stride_cap = PM_DP_MAX;
  const size_t n = (size_t)chunk;
  for (auto& sl : h->slots) {
    for (int m = 0; m < 2; m++) {
      CK(cudaMalloc(&sl.d_reads[m], n * h->stride_cap));
      CK(cudaMalloc(&sl.d_len[m], n * sizeof(int)));
      CK(cudaHostAlloc(&sl.h_reads[m], n * h->stride_cap, cudaHostAllocDefault));
      CK(cudaHostAlloc(&sl.h_len[m], n * sizeof(int), cudaHostAllocDefault));
    }
    CK(cudaHostAlloc(&sl.h_m1, n * 4, cudaHostAllocDefault));
    CK(cudaHostAlloc(&sl.h_m2, n * 4, cudaHostAllocDefault));
    CK(cudaHostAlloc(&sl.h_type, n * 4, cudaHostAllocDefault));
    CK(cudaMalloc(&sl.d_m1, n * 4));
    CK(cudaMalloc(&sl.d_m2, n * 4));
    CK(cudaMalloc(&sl.d_type, n * 4));
  }
  h->task_cap = (uint32_t)std::min<size_t>(2 * n * PM_MAX_HITS, 0x7FFFFFFFull);
  if (const char* s = getenv("PEMAP_CERTIFY")) h->certify = atoi(s) != 0;
  if (const char* s = getenv("PEMAP_EXACT")) h->exact = atoi(s) != 0;
  if (const char* s = getenv("PEMAP_BAND_HALF")) h->band_half = std::min(PM_BAND_LANES / 2, std::max(0, atoi(s)));
  h->seed_blocks = h->sm_count * 8;
  h->big_grid = 2 * h->sm_count;
  h->sw_blocks = h->sm_count * 6;  // upper bound of CTAs per SM of the wavefront kernels (scratch is sized for it)
  {
    int rc = alloc_lane_scratch(h);  // lane 0: the context's own fields
    if (rc) return rc;
    pemap_lanes& L = *h->lanes_;
    if (L.n > 1) {
      use_lane(h, 1);
      CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      rc = alloc_lane_scratch(h);
      use_lane(h, 0);
      if (rc) return rc;
      CK(cudaEventCreateWithFlags(&L.ev_main, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&L.ev_side, cudaEventDisableTiming));
    }
  }
  CK(cudaMalloc(&h->d_counters, sizeof(pm::SeedCounters)));
  CK(cudaMemset(h->d_counters, 0, sizeof(pm::SeedCounters)));
  if (const char* s = getenv("PEMAP_INS_MB")) h->ins_cap = (uint64_t)std::max(1, atoi(s)) << 20;
  if (const char* s = getenv("PEMAP_INS_BYTES")) h->ins_cap = (uint64_t)std::max(1024, atoi(s));  // tests of the spill path
  CK(cudaMalloc(&h->d_ins, h->ins_cap));
  CK(cudaMalloc(&h->d_ins_cursor, 8));
  CK(cudaMemset(h->d_ins_cursor, 0, 8));
  CK(cudaHostAlloc(&h->h_ins_used, 2 * 8, cudaHostAllocDefault));
  h->h_ins_used[0] = h->h_ins_used[1] = 0;
  return PEMAP_OK;
}

// Bloom filter of the occupied k-mers, sized to stay resident in L2 (persisting access-policy window on the stream).
// Worth it while the genome is small against the 2^32 k-mer space; PEMAP_FILTER=0 disables it.
int build_filter(pemap_ctx* h) {
  if (const char* s = getenv("PEMAP_FILTER"))
    if (atoi(s) == 0) return PEMAP_OK;
  if (h->n_mers == 0 || h->n_mers > (1ull << 29)) return PEMAP_OK;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  size_t cap = 64ull << 20;
  if (const char* s = getenv("PEMAP_FILTER_MB")) cap = (size_t)std::max(1, atoi(s)) << 20;
  if (prop.persistingL2CacheMaxSize > 0) cap = std::min(cap, (size_t)prop.persistingL2CacheMaxSize);
  int lg = 18;  // 32-bit words: 2^18 * 4 B = 1 MB minimum
  while ((4ull << (lg + 1)) <= cap && (32ull << lg) < 16ull * h->n_mers) lg++;  // up to 16 bits per k-mer
  while ((4ull << lg) > cap && lg > 10) lg--;
  const double bits_per_key = (double)(32ull << lg) / (double)h->n_mers;
  if (bits_per_key < 1.5) return PEMAP_OK;   // too dense to reject much: go straight to the table
  h->filter_bytes = 4ull << lg;
  h->filter_shift = 32 - lg;
  h->filter_k = bits_per_key >= 12.0 ? 3 : bits_per_key >= 3.0 ? 2 : 1;
  CK(cudaMalloc(&h->d_filter, h->filter_bytes));
  CK(cudaMemsetAsync(h->d_filter, 0, h->filter_bytes, h->stream));
  pm::k_filter_build<<<1u << 22, 256, 0, h->stream>>>(h->d_pos_index, h->d_filter, h->filter_shift, h->filter_k);
  CK(cudaGetLastError());
  if (prop.persistingL2CacheMaxSize > 0) {
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t)prop.persistingL2CacheMaxSize, h->filter_bytes));
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    av.accessPolicyWindow.base_ptr = h->d_filter;
    av.accessPolicyWindow.num_bytes = std::min(h->filter_bytes, (size_t)prop.accessPolicyMaxWindowSize);
    av.accessPolicyWindow.hitRatio = 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaGetLastError();
  }
  if (getenv("PEMAP_VERBOSE"))
    fprintf(stderr, "pemap: k-mer filter %zu MB, %d bits/k-mer set, %.1f bits per k-mer; persisting L2 max %d MB\n",
            h->filter_bytes >> 20, h->filter_k, bits_per_key, prop.persistingL2CacheMaxSize >> 20);
  return PEMAP_OK;
}

void read_index_env(pemap_ctx* h) {
  if (const char* s = getenv("PEMAP_SEED")) h->seed_legacy = strcmp(s, "legacy") == 0;
  if (const char* s = getenv("PEMAP_INDEX_ONLY")) h->index_only = atoi(s) != 0;
}

// pos_index / mers (28 GB on a human-sized genome) are the input format, not what the seed kernel reads: beside the
// rotated bucket index they are kept only while small (index tests, PEMAP_SEED=legacy); PEMAP_KEEP_INDEX=0/1 overrides
void drop_file_index_if_large(pemap_ctx* h) {
  bool keep = h->n_mers <= (1ull << 30);
  if (const char* s = getenv("PEMAP_KEEP_INDEX")) keep = atoi(s) != 0;
  if (keep) return;
  cudaFree(h->d_pos_index);
  cudaFree(h->d_mers);
  h->d_pos_index = nullptr;
  h->d_mers = nullptr;
}

// Rotated bucket index from the code-sorted (code, position) list of the indexed k-mers (= pos_index / mers order).
// `code` is overwritten (sort output); `val` is only read.
int build_rbi(pemap_ctx* h, uint32_t* code, const uint32_t* val, uint64_t n) {
  const uint32_t T = (uint32_t)h->params.too_many_spots;
  uint32_t p0 = 0, plast = 0;
  CK(cudaMemcpy(&p0, h->d_pos_index, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&plast, h->d_pos_index + 0xFFFFFFFFull, 4, cudaMemcpyDeviceToHost));
  const uint32_t last_cnt = p0 - plast;  // get_mers(0xFFFFFFFF): which + 1 wraps in 32 bits (pemapper.c:2163)
  const uint64_t true_last = n - plast;
  if (last_cnt < T && (uint64_t)last_cnt > true_last)
    return fail(h, PEMAP_ERR_UNSUPPORTED, "index within 100 entries of 2^32: get_mers(0xFFFFFFFF) reads past the mers array");
  uint32_t *s1 = nullptr, *s2 = nullptr, *keyg = nullptr, *o2 = nullptr, *bstart = nullptr, *units = nullptr;
  unsigned char* flag = nullptr;
  unsigned long long* d_nsel = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    void* v[] = {s1, s2, keyg, o2, bstart, units, flag, d_nsel, tmp};
    for (void* p : v)
      if (p) cudaFree(p);
  };
#define CKB(call)                                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      cleanup();                                                                                          \
      return fail(h, e_ == cudaErrorMemoryAllocation ? PEMAP_ERR_NOMEM : PEMAP_ERR_CUDA,                  \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                    \
    }                                                                                                     \
  } while (0)
  const size_t n1 = (size_t)n + 1;
  CKB(cudaMalloc(&s1, n1 * 4));
  CKB(cudaMalloc(&s2, n1 * 4));
  CKB(cudaMalloc(&flag, n1));
  CKB(cudaMalloc(&d_nsel, 8));
  const unsigned nblk = (unsigned)((n + 255) / 256);
  if (n) pm::k_rbi_flag<<<nblk, 256, 0, h->stream>>>(code, h->d_pos_index, n, T, last_cnt, flag);
  size_t need = 0, tmp_bytes = 0;
  const long long n_items = (long long)n;
  cub::DeviceSelect::Flagged(nullptr, need, code, flag, s1, d_nsel, n_items, h->stream);
  tmp_bytes = need;
  cub::DeviceRadixSort::SortPairs(nullptr, need, s1, code, s2, s2, n_items + 1, 0, 32, h->stream);
  tmp_bytes = std::max(tmp_bytes, need);
  cub::DeviceScan::ExclusiveSum(nullptr, need, s1, s2, (1 << 24) + 1, h->stream);
  tmp_bytes = std::max(tmp_bytes, need);
  CKB(cudaMalloc(&tmp, tmp_bytes));
  need = tmp_bytes;
  CKB(cub::DeviceSelect::Flagged(tmp, need, code, flag, s1, d_nsel, n_items, h->stream));
  need = tmp_bytes;
  CKB(cub::DeviceSelect::Flagged(tmp, need, val, flag, s2, d_nsel, n_items, h->stream));
  unsigned long long ne = 0;
  CKB(cudaMemcpyAsync(&ne, d_nsel, 8, cudaMemcpyDeviceToHost, h->stream));
  CKB(cudaStreamSynchronize(h->stream));
  cudaFree(flag);
  flag = nullptr;
  if (ne) pm::k_rbi_mark<<<(unsigned)((ne + 255) / 256), 256, 0, h->stream>>>(s1, h->d_pos_index, ne, T, last_cnt, s2);
  if (last_cnt >= T && true_last == 0) {  // poly-T is "crowded" through the wrap although the genome has none: a lone marker
    const uint32_t kv[2] = {0xFFFFFFFFu, PM_RBI_MARK};
    CKB(cudaMemcpyAsync(s1 + ne, &kv[0], 4, cudaMemcpyHostToDevice, h->stream));
    CKB(cudaMemcpyAsync(s2 + ne, &kv[1], 4, cudaMemcpyHostToDevice, h->stream));
    CKB(cudaStreamSynchronize(h->stream));
    ne++;
  }
  CKB(cudaMalloc(&keyg, n1 * 4));
  CKB(cudaMalloc(&o2, n1 * 4));
  const unsigned nb24 = (1u << 24) + 1u;
  CKB(cudaMalloc(&bstart, ((size_t)nb24 + 1) * 4));
  CKB(cudaMalloc(&units, (size_t)nb24 * 4));
  const unsigned eblk = (unsigned)((ne + 255) / 256);
  h->rbi_bytes = 0;
  for (int g = 0; g < 4; g++) {
    const uint32_t *kk = s1, *vv = s2;  // rotation 0: tag = low byte, the code order is the key order
    if (g > 0) {
      if (ne) pm::k_rbi_keys<<<eblk, 256, 0, h->stream>>>(s1, ne, g, keyg);
      need = tmp_bytes;
      CKB(cub::DeviceRadixSort::SortPairs(tmp, need, keyg, code, s2, o2, (long long)ne, 0, 32, h->stream));
      kk = code;
      vv = o2;
    }
    pm::k_rbi_bucket_starts<<<(nb24 + 255) / 256, 256, 0, h->stream>>>(kk, ne, bstart);
    pm::k_rbi_bucket_units<<<(nb24 + 255) / 256, 256, 0, h->stream>>>(bstart, units);
    CKB(cudaMalloc(&h->d_rbi_dir[g], (size_t)nb24 * 4));
    need = tmp_bytes;
    CKB(cub::DeviceScan::ExclusiveSum(tmp, need, units, h->d_rbi_dir[g], (int)nb24, h->stream));
    uint32_t total_units = 0;
    CKB(cudaMemcpyAsync(&total_units, h->d_rbi_dir[g] + (1u << 24), 4, cudaMemcpyDeviceToHost, h->stream));
    CKB(cudaStreamSynchronize(h->stream));
    const size_t bytes = ((size_t)total_units + 1) * PM_RBI_BLOCK_BYTES;
    CKB(cudaMalloc(&h->d_rbi_data[g], bytes));
    CKB(cudaMemsetAsync(h->d_rbi_data[g], 0xFF, bytes, h->stream));
    if (ne) pm::k_rbi_fill<<<eblk, 256, 0, h->stream>>>(kk, vv, ne, bstart, h->d_rbi_dir[g], h->d_rbi_data[g]);
    h->rbi_bytes += bytes + (size_t)nb24 * 4;
  }
  CKB(cudaStreamSynchronize(h->stream));
  CKB(cudaGetLastError());
#undef CKB
  cleanup();
  if (getenv("PEMAP_VERBOSE"))
    fprintf(stderr, "pemap: rotated bucket index: %llu entries (%llu positions), %.2f GB\n", ne, (unsigned long long)n,
            h->rbi_bytes / 1e9);
  return PEMAP_OK;
}

int finish_init(pemap_ctx* h) {
  fill_dev_params(h);
  if (h->index_only) {
    CK(cudaDeviceSynchronize());
    return PEMAP_OK;
  }
  if (h->seed_legacy) {
    int rc0 = build_filter(h);
    if (rc0) return rc0;
  }
  int rc = upload_border(h);
  if (rc) return rc;
  CK(cudaMalloc(&h->d_counts, (size_t)h->genome_size * 6 * 4 + 64));
  CK(cudaMemset(h->d_counts, 0, (size_t)h->genome_size * 6 * 4 + 64));
  rc = alloc_chunk_buffers(h);
  if (rc) return rc;
  memset(&h->stats, 0, sizeof(h->stats));
  CK(cudaDeviceSynchronize());
  return PEMAP_OK;
}

int upload_cstart(pemap_ctx* h, const uint32_t* cs, int n_contigs) {
  std::vector<uint32_t> pad((size_t)std::max(n_contigs + 1, 16), 0u);
  for (int i = 0; i <= n_contigs; i++) pad[i] = cs[i];
  CK(cudaMalloc(&h->d_cstart, pad.size() * 4));
  CK(cudaMemcpy(h->d_cstart, pad.data(), pad.size() * 4, cudaMemcpyHostToDevice));
  return PEMAP_OK;
}

int check_contigs(pemap_ctx* h, int n) {
  if (n < 1) return fail(h, PEMAP_ERR_ARG, "no contigs");
  // PEMAP_INDEX_ONLY=1 (set by index_genome_gpu): the handle is only used to build and read back the index
  if (n >= 2 && n <= 7 && !(getenv("PEMAP_INDEX_ONLY") && atoi(getenv("PEMAP_INDEX_ONLY"))))
    return fail(h, PEMAP_ERR_UNSUPPORTED,
                "2..7 contigs: the reference's find_chrom (pemapper.c:2168-2186) starts its bisection at index 7 and "
                "reads past contig_starts; its result is undefined there (SURVEY.md section 7-C). Use 1 or >= 8 contigs.");
  return PEMAP_OK;
}

// persistent grid-stride kernels: exactly one wave, as many CTAs per SM as registers and shared memory allow
// (capped by the per-group scratch the context allocated: h->sw_blocks CTAs)
template <class K>
int one_wave_grid(pemap_ctx* h, K kernel, int block, size_t dyn) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return std::min(h->sw_blocks, per_sm * h->sm_count);
}

template <int G, int WD, int MODE>
void launch_sw(pemap_ctx* h, const pm::SwArgs& a) {
  const size_t dyn = MODE == 2 ? pm::trace_band_bytes<G, WD>() : 0;
  static int grids[kMaxDev] = {};  // per instantiation and device (function attributes are per device)
  int& grid = grids[h->device % kMaxDev];
  if (!grid) {
    if (MODE == 2) cudaFuncSetAttribute(pm::k_sw_fp64<G, WD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    grid = one_wave_grid(h, pm::k_sw_fp64<G, WD, MODE>, 128, dyn);
  }
  pm::k_sw_fp64<G, WD, MODE><<<grid, 128, dyn, h->stream>>>(a);
}

template <int MODE>
void dispatch_sw(pemap_ctx* h, const pm::SwArgs& a, int max_len) {
  if (max_len <= 112) launch_sw<16, 7, MODE>(h, a);
  else if (max_len <= 160) launch_sw<16, 10, MODE>(h, a);
  else if (max_len <= 256) launch_sw<32, 8, MODE>(h, a);
  else launch_sw<32, 10, MODE>(h, a);
  h->stats.launches++;
}

// flag scratch of the packed integer traceback: one record per winner pair, sized for a chunk whose read-mates all have gaps
int ensure_flag_scratch(pemap_ctx* h, size_t bytes_per_pair, size_t code_bytes) {
  const size_t pairs = (size_t)h->chunk + 1;  // 2 * chunk read-mates / 2
  const size_t need = pairs * bytes_per_pair, need_codes = pairs * code_bytes;
  if (h->flagq_bytes < need) {
    if (h->d_flagq) cudaFree(h->d_flagq);
    h->d_flagq = nullptr;
    h->flagq_bytes = 0;
    CK(cudaMalloc(&h->d_flagq, need));
    h->flagq_bytes = need;
  }
  if (h->pair_codes_bytes < need_codes) {
    if (h->d_pair_codes) cudaFree(h->d_pair_codes);
    h->d_pair_codes = nullptr;
    h->pair_codes_bytes = 0;
    CK(cudaMalloc(&h->d_pair_codes, need_codes));
    h->pair_codes_bytes = need_codes;
  }
  if (!h->d_walk_meta) CK(cudaMalloc(&h->d_walk_meta, pairs * 4 * sizeof(uint4)));
  return PEMAP_OK;
}

template <int G, int WD>
int launch_trace_int(pemap_ctx* h, pm::TraceIntArgs& a) {
  int rc = ensure_flag_scratch(h, pm::trace_flag_bytes_per_pair<G, WD>(), (size_t)pm::trace_code_bytes<G, WD>());
  if (rc) return rc;
  a.flagq = reinterpret_cast<uint2*>(h->d_flagq);
  a.pair_codes = reinterpret_cast<unsigned char*>(h->d_pair_codes);
  a.walk_meta = reinterpret_cast<uint4*>(h->d_walk_meta);
  const size_t dyn_dp = pm::trace_dp16_smem<G, WD>(), dyn_wk = pm::trace_walk16_smem<G, WD>();
  static int grids_dp[kMaxDev] = {}, grids_wk[kMaxDev] = {};
  int& grid_dp = grids_dp[h->device % kMaxDev];
  int& grid_wk = grids_wk[h->device % kMaxDev];
  if (!grid_dp) {
    cudaFuncSetAttribute(pm::k_trace_dp16<G, WD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_dp);
    cudaFuncSetAttribute(pm::k_trace_walk16<G, WD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_wk);
    grid_dp = one_wave_grid(h, pm::k_trace_dp16<G, WD>, 128, dyn_dp);
    {  // 4 walkers per CTA; the segment scratch (d_pend) holds 16 walkers per sw_blocks entry
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_trace_walk16<G, WD>, 128, dyn_wk) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
      }
      grid_wk = std::min(h->sw_blocks * 4, per_sm * h->sm_count);
    }
    if (getenv("PEMAP_VERBOSE"))
      fprintf(stderr, "pemap: k_trace_dp16<%d,%d> grid %d (%zu B dyn smem), k_trace_walk16 grid %d (%zu B)\n", G, WD, grid_dp,
              dyn_dp, grid_wk, dyn_wk);
  }
  pm::k_trace_dp16<G, WD><<<grid_dp, 128, dyn_dp, h->stream>>>(a);
  pm::k_trace_walk16<G, WD><<<grid_wk, 128, dyn_wk, h->stream>>>(a);
  h->stats.launches += 2;
  return PEMAP_OK;
}

int dispatch_trace_int(pemap_ctx* h, pm::TraceIntArgs& a, int max_len) {
  if (max_len <= 112) return launch_trace_int<16, 7>(h, a);
  if (max_len <= 160) return launch_trace_int<16, 10>(h, a);
  if (max_len <= 256) return launch_trace_int<32, 8>(h, a);
  return launch_trace_int<32, 10>(h, a);
}

template <int G, int WD, int CMM>
struct CmmDispatch {
  static void go(pemap_ctx* h, const pm::SwIntArgs& a, int cmm) {
    if (cmm == CMM) {
      static int grids[kMaxDev] = {};
      int& grid = grids[h->device % kMaxDev];
      if (!grid) grid = one_wave_grid(h, pm::k_sw_i16<G, WD, CMM>, 128, 0);
      pm::k_sw_i16<G, WD, CMM><<<grid, 128, 0, h->stream>>>(a);
    }
    else CmmDispatch<G, WD, CMM - 1>::go(h, a, cmm);
  }
};
template <int G, int WD>
struct CmmDispatch<G, WD, -1> {
  static void go(pemap_ctx* h, const pm::SwIntArgs& a, int) {
    static int grids[kMaxDev] = {};
    int& grid = grids[h->device % kMaxDev];
    if (!grid) grid = one_wave_grid(h, pm::k_sw_i16<G, WD, -1>, 128, 0);
    pm::k_sw_i16<G, WD, -1><<<grid, 128, 0, h->stream>>>(a);
  }
};

// uniform_len > 0: every read of the chunk has that length (lets the kernel fix the last read column at compile time)
void dispatch_sw_int(pemap_ctx* h, pm::SwIntArgs& a, int max_len, int uniform_len) {
  int wd = max_len <= 112 ? 7 : max_len <= 160 ? 10 : max_len <= 256 ? 8 : 10;
  int cmm = uniform_len > 0 ? (uniform_len - 1) % wd : -1;
  a.lane_mm = uniform_len > 0 ? (uniform_len - 1) / wd : -1;
  if (max_len <= 112) CmmDispatch<16, 7, 6>::go(h, a, cmm);
  else if (max_len <= 160) CmmDispatch<16, 10, 9>::go(h, a, cmm);
  else if (max_len <= 256) CmmDispatch<32, 8, 7>::go(h, a, cmm);
  else CmmDispatch<32, 10, 9>::go(h, a, cmm);
  h->stats.launches++;
}

// the round-1 seed kernel over pos_index / mers (PEMAP_SEED=legacy; cross-check of the rotated bucket index)
int launch_seed_legacy(pemap_ctx* h, pm::SeedArgs& sa, int n, const char* d_r1, const int* d_l1, const char* d_r2, const int* d_l2,
                       int stride, bool paired) {
  if (!h->d_pos_index) return fail(h, PEMAP_ERR_ARG, "PEMAP_SEED=legacy needs the file index resident (PEMAP_KEEP_INDEX=1)");
  pm::DevParams keep_p = sa.p;
  sa.pos_index = h->d_pos_index;
  sa.mers = h->d_mers;
  sa.cstart = h->d_cstart;
  sa.reads[0] = d_r1;
  sa.reads[1] = d_r2;
  sa.len[0] = d_l1;
  sa.len[1] = d_l2;
  sa.stride = stride;
  sa.n_reads = n;
  sa.paired = paired ? 1 : 0;
  sa.scratch = h->d_seed_scratch;
  sa.tasks = h->d_tasks;
  sa.task_cursor = h->d_cursors;
  sa.task_cap = h->task_cap;
  sa.cand_base = h->d_cand_base;
  sa.cand_n = h->d_cand_n;
  sa.counters = h->d_counters;
  sa.filter = h->d_filter;
  sa.filter_shift = h->filter_shift;
  sa.filter_k = h->filter_k;
  sa.p = keep_p;
  const int work = paired ? 2 * n : n;
  static int seed_waves[kMaxDev][4] = {};  // one resident wave of the persistent seed kernel, per device and variant
  const int sv = h->d_filter ? std::min(3, std::max(1, h->filter_k)) : 0;  // bits per k-mer of the filter, 0 = none
  int* seed_wave = seed_waves[h->device % kMaxDev];
  if (!seed_wave[sv]) {
    int per_sm = 0;
    cudaError_t oe =
        sv == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_seed_chain<kSeedWarps, 0>, kSeedWarps * 32, 0)
        : sv == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_seed_chain<kSeedWarps, 1>, kSeedWarps * 32, 0)
        : sv == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_seed_chain<kSeedWarps, 2>, kSeedWarps * 32, 0)
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_seed_chain<kSeedWarps, 3>, kSeedWarps * 32, 0);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetLastError();
    // Measured on cfg2 (profiles/README_r01.md, r01e): with the L2-resident filter the kernel is fastest when compiled
    // for and run at 5 CTAs per SM (62.6 ms per 8.4 M read-mates against 64.3 at 6 and 72.2 at 8; capping the grid of
    // the 8-CTA build at 5 gains nothing: it is the registers per warp, i.e. the loads each warp keeps in flight).
    // Without the filter (human-sized genomes) it takes every CTA that fits.
    int want = sv ? PM_SEED_CTAS_FILT : per_sm;
    if (const char* s = getenv("PEMAP_SEED_CTAS")) want = std::max(1, atoi(s));
    per_sm = std::min(per_sm, want);
    seed_wave[sv] = std::min(h->seed_blocks, per_sm * h->sm_count);
    if (getenv("PEMAP_VERBOSE")) fprintf(stderr, "pemap: seed kernel %d CTAs per SM\n", per_sm);
  }
  const int seed_grid = std::min(seed_wave[sv], (work + kSeedWarps - 1) / kSeedWarps);
  if (sv == 0) pm::k_seed_chain<kSeedWarps, 0><<<seed_grid, kSeedWarps * 32, 0, h->stream>>>(sa);
  else if (sv == 1) pm::k_seed_chain<kSeedWarps, 1><<<seed_grid, kSeedWarps * 32, 0, h->stream>>>(sa);
  else if (sv == 2) pm::k_seed_chain<kSeedWarps, 2><<<seed_grid, kSeedWarps * 32, 0, h->stream>>>(sa);
  else pm::k_seed_chain<kSeedWarps, 3><<<seed_grid, kSeedWarps * 32, 0, h->stream>>>(sa);
  return PEMAP_OK;
}

// map one chunk whose reads are already in d_r1/d_r2 (device); results go to d_m1/d_m2/d_type (device)
int run_chunk(pemap_ctx* h, int n, const char* d_r1, const int* d_l1, const char* d_r2, const int* d_l2, int stride,
              int max_len, int uniform_len, uint32_t* d_m1, uint32_t* d_m2, int* d_type, cudaEvent_t* ev) {
  const bool paired = h->params.pair_flag && d_r2;
  const bool exact = h->exact || (h->keep & PEMAP_KEEP_DETAIL) || h->params.match_bonus != 1.0;
  bool forked = false;
  CK(cudaMemsetAsync(h->d_cursors, 0, 16, h->stream));
  CK(cudaMemsetAsync(h->d_cursors + 6, 0, 52, h->stream));  // [6..8] lists, [9..15] work counters of the persistent kernels, [16] DP list, [17]/[18] seed lists of the second / third pass
  CK(cudaEventRecord(ev[0], h->stream));
  pm::SeedArgs sa;
  sa.p = h->dp;
  sa.p.pair_flag = paired ? 1 : 0;
  if (h->seed_legacy) {
    int rc = launch_seed_legacy(h, sa, n, d_r1, d_l1, d_r2, d_l2, stride, paired);
    if (rc) return rc;
  } else {
    pm::SeedRbiArgs ra;
    for (int g = 0; g < 4; g++) {
      ra.ix.data[g] = h->d_rbi_data[g];
      ra.ix.dir[g] = h->d_rbi_dir[g];
    }
    ra.cstart = h->d_cstart;
    ra.reads[0] = d_r1;
    ra.reads[1] = d_r2;
    ra.len[0] = d_l1;
    ra.len[1] = d_l2;
    ra.stride = stride;
    ra.packed[0] = h->cur_packed.rows[0];
    ra.packed[1] = h->cur_packed.rows[1];
    ra.pstride = h->cur_packed.stride;
    ra.pcode_words = h->cur_packed.code_words;
    ra.pmask_words = h->cur_packed.mask_words;
    ra.n_reads = n;
    ra.paired = paired ? 1 : 0;
    ra.tasks = h->d_tasks;
    ra.task_cursor = h->d_cursors;
    ra.task_cap = h->task_cap;
    ra.cand_base = h->d_cand_base;
    ra.cand_n = h->d_cand_n;
    ra.counters = h->d_counters;
    ra.big_scratch = h->d_big_scratch;
    ra.fast_cap = PM_RBI_CAP;
    if (const char* s = getenv("PEMAP_RBI_CAP")) ra.fast_cap = std::min(PM_RBI_CAP, std::max(1, atoi(s)));
    ra.p = sa.p;
    const int work = paired ? 2 * n : n;
    static int rbi_wave[kMaxDev] = {}, rbi_wave2[kMaxDev] = {};
    int& wave = rbi_wave[h->device % kMaxDev];
    int& wave2 = rbi_wave2[h->device % kMaxDev];
    constexpr size_t dyn = pm::seed_rbi_smem<kRbiWarps, PM_RBI_CAP>(), dyn2 = pm::seed_rbi_smem<kRbiMidWarps, PM_RBI_CAP2>(),
                     dyn_big = pm::seed_rbi_smem<kRbiBigWarps, 0>();
    if (!wave) {
      cudaFuncSetAttribute(pm::k_seed_rbi<kRbiWarps, PM_RBI_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      cudaFuncSetAttribute(pm::k_seed_rbi<kRbiMidWarps, PM_RBI_CAP2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2);
      cudaFuncSetAttribute(pm::k_seed_rbi<kRbiBigWarps, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_big);
      int per_sm = 0, per_sm2 = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm::k_seed_rbi<kRbiWarps, PM_RBI_CAP>, kRbiWarps * 32, dyn) != cudaSuccess ||
          per_sm < 1)
        per_sm = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, pm::k_seed_rbi<kRbiMidWarps, PM_RBI_CAP2>, kRbiMidWarps * 32, dyn2) !=
              cudaSuccess || per_sm2 < 1)
        per_sm2 = 1;
      cudaGetLastError();
      if (const char* s = getenv("PEMAP_SEED_CTAS")) per_sm = std::min(per_sm, std::max(1, atoi(s)));
      wave = per_sm * h->sm_count;
      wave2 = per_sm2 * h->sm_count;
      if (getenv("PEMAP_VERBOSE"))
        fprintf(stderr, "pemap: k_seed_rbi %d CTAs per SM (%zu B of shared memory each), second pass %d CTAs per SM (%zu B)\n", per_sm,
                dyn, per_sm2, dyn2);
    }
    // pass 1: every read-mate, 512 entries per strand in shared memory; pass 2: what did not fit, 2048 entries (reads in
    // repeats); pass 3: the rest in the per-warp global scratch (any strand fits)
    ra.work_list = nullptr;
    ra.work_n = nullptr;
    ra.next_list = h->d_big_list;
    ra.next_cursor = h->d_cursors + 17;
    const int grid = std::min(wave, (work + kRbiWarps - 1) / kRbiWarps);
    if (grid > 0) pm::k_seed_rbi<kRbiWarps, PM_RBI_CAP><<<grid, kRbiWarps * 32, dyn, h->stream>>>(ra);
    ra.work_list = h->d_big_list;
    ra.work_n = h->d_cursors + 17;
    ra.next_list = h->d_big_list2;
    ra.next_cursor = h->d_cursors + 18;
    ra.fast_cap = ra.fast_cap < PM_RBI_CAP ? 4 * ra.fast_cap : PM_RBI_CAP2;
    pm::k_seed_rbi<kRbiMidWarps, PM_RBI_CAP2><<<wave2, kRbiMidWarps * 32, dyn2, h->stream>>>(ra);
    ra.work_list = h->d_big_list2;
    ra.work_n = h->d_cursors + 18;
    ra.next_list = nullptr;
    ra.next_cursor = nullptr;
    pm::k_seed_rbi<kRbiBigWarps, 0><<<h->big_grid, kRbiBigWarps * 32, dyn_big, h->stream>>>(ra);
    h->stats.launches += 2;
    h->stats.launches++;
  }
  h->stats.launches++;
  CK(cudaEventRecord(ev[1], h->stream));

  pm::SwArgs wa;
  wa.tasks = h->d_tasks;
  wa.results = h->d_results;
  wa.winners = h->d_winners;
  wa.list_mode = 0;
  wa.n_items = h->d_cursors;
  wa.work = h->d_cursors + 9;
  wa.reads[0] = d_r1;
  wa.reads[1] = d_r2;
  wa.len[0] = d_l1;
  wa.len[1] = d_l2;
  wa.stride = stride;
  wa.genome = h->d_genome;
  wa.border = h->d_border;
  wa.dirs = h->d_dirs;
  wa.sink.counts = h->d_counts;
  wa.sink.pend = h->d_pend;
  wa.sink.ins_buf = h->d_ins;
  wa.sink.ins_cursor = h->d_ins_cursor;
  wa.sink.ins_cap = h->ins_cap;
  wa.oob_winners = h->d_oob_winners;
  wa.oob_cursor = h->d_cursors + 8;
  wa.counters = h->d_counters;
  wa.band_half = h->band_half;
  wa.p = sa.p;

  pm::SelectArgs se;
  se.tasks = h->d_tasks;
  se.results = h->d_results;
  se.cand_base = h->d_cand_base;
  se.cand_n = h->d_cand_n;
  se.len[0] = d_l1;
  se.len[1] = d_l2;
  se.n_reads = n;
  se.read_list = nullptr;
  se.n_list = nullptr;
  se.m1 = d_m1;
  se.m2 = d_m2;
  se.mapping_type = d_type;
  const bool keep_det = (h->keep & PEMAP_KEEP_DETAIL) != 0;
  se.det_best = keep_det ? h->d_det_best : nullptr;
  se.det_orient = keep_det ? h->d_det_orient : nullptr;
  se.det_score = keep_det ? h->d_det_score : nullptr;
  se.winners = h->d_winners;
  se.winner_cursor = h->d_cursors + 1;
  se.p = sa.p;

  if (exact) {  // the reference's arithmetic for every candidate
    dispatch_sw<0>(h, wa, max_len);
    CK(cudaEventRecord(ev[2], h->stream));
    pm::k_select<<<(n + 127) / 128, 128, 0, h->stream>>>(se);
    h->stats.launches++;
  } else {
    // integer DPX scoring of every candidate, integer selection, then fp64 replay of the reads with rational ties
    pm::SwIntArgs ia;
    ia.tasks = h->d_tasks;
    ia.results = h->d_ires;
    ia.n_items = h->d_cursors;
    ia.work = h->d_cursors + 10;
    ia.reads[0] = d_r1;
    ia.reads[1] = d_r2;
    ia.len[0] = d_l1;
    ia.len[1] = d_l2;
    ia.stride = stride;
    ia.genome = h->d_genome;
    ia.lane_mm = -1;
    ia.p = sa.p;
    ia.list = nullptr;
    if (h->certify) {  // candidates decided by their ungapped diagonals alone skip the DP
      pm::CertifyArgs ca;
      ca.tasks = h->d_tasks;
      ca.results = h->d_ires;
      ca.n_items = h->d_cursors;
      ca.list = h->d_sw_list;
      ca.list_cursor = h->d_cursors + 16;
      ca.reads[0] = d_r1;
      ca.reads[1] = d_r2;
      ca.len[0] = d_l1;
      ca.len[1] = d_l2;
      ca.stride = stride;
      ca.genome = h->d_genome;
      ca.cells_certified = &h->d_counters->sw_cells_certified;
      ca.p = sa.p;
      if (sa.p.is_bisulfite) pm::k_diag_certify<4, true><<<h->sm_count * 16, 128, 0, h->stream>>>(ca);
      else pm::k_diag_certify<4, false><<<h->sm_count * 16, 128, 0, h->stream>>>(ca);
      h->stats.launches++;
      ia.list = h->d_sw_list;
      ia.n_items = h->d_cursors + 16;
    }
    dispatch_sw_int(h, ia, max_len, uniform_len);
    CK(cudaEventRecord(ev[2], h->stream));
    pm::SelectIntArgs si;
    si.tasks = h->d_tasks;
    si.ires = h->d_ires;
    si.results64 = h->d_results;
    si.cand_base = h->d_cand_base;
    si.cand_n = h->d_cand_n;
    si.len[0] = d_l1;
    si.len[1] = d_l2;
    si.n_reads = n;
    si.m1 = d_m1;
    si.m2 = d_m2;
    si.mapping_type = d_type;
    si.winners = h->d_winners;
    si.winner_cursor = h->d_cursors + 1;
    si.diag_winners = h->d_diag_winners;
    si.diag_cursor = h->d_cursors + 6;
    si.replay_reads = h->d_replay_reads;
    si.replay_read_cursor = h->d_cursors + 2;
    si.replay_tasks = h->d_replay_tasks;
    si.replay_task_cursor = h->d_cursors + 3;
    si.counters = h->d_counters;
    si.reads[0] = d_r1;
    si.reads[1] = d_r2;
    si.stride = stride;
    si.genome = h->d_genome;
    si.border = h->d_border;
    si.p = sa.p;
    pm::k_select_int<<<(n + 127) / 128, 128, 0, h->stream>>>(si);
    // exact replay: fp64 scores of every candidate of the flagged reads, then the fp64 selection rules on them
    pm::SwArgs ra = wa;
    ra.winners = h->d_replay_tasks;
    ra.list_mode = 1;
    ra.n_items = h->d_cursors + 3;
    ra.work = h->d_cursors + 11;
    dispatch_sw<0>(h, ra, max_len);
    se.read_list = h->d_replay_reads;
    se.n_list = h->d_cursors + 2;
    pm::k_select<<<(n + 127) / 128, 128, 0, h->stream>>>(se);
    h->stats.launches += 2;
  }
  CK(cudaEventRecord(ev[3], h->stream));

  if (!exact) {  // winners whose walk is a pure diagonal need no DP recompute
    pm::DiagArgs da;
    da.tasks = h->d_tasks;
    da.ires = h->d_ires;
    da.winners = h->d_diag_winners;
    da.n_items = h->d_cursors + 6;
    da.reads[0] = d_r1;
    da.reads[1] = d_r2;
    da.len[0] = d_l1;
    da.len[1] = d_l2;
    da.stride = stride;
    da.counts = h->d_counts;
    da.counters = h->d_counters;
    // k_apply_diag is bound by L2 atomics and leaves the ALUs idle; the integer traceback is the opposite.  Two CTAs
    // per SM on a side stream leave room for the traceback kernels' resident waves; joined before ev[4].
    forked = h->diag_overlap != 0;
    if (forked) {
      CK(cudaEventRecord(h->ev_fork, h->stream));
      CK(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
      pm::k_apply_diag<<<h->sm_count * 2, 256, 0, h->s_aux>>>(da);
      CK(cudaEventRecord(h->ev_join, h->s_aux));
    } else {
      pm::k_apply_diag<<<h->sm_count * 8, 256, 0, h->stream>>>(da);
    }
    h->stats.launches++;
    CK(cudaEventRecord(ev[5], h->stream));
    // the other winners: integer traceback; the ones with a rational tie on their path fall through to fp64
    pm::TraceIntArgs ta;
    ta.tasks = h->d_tasks;
    ta.results = h->d_results;
    ta.winners = h->d_winners;
    ta.n_items = h->d_cursors + 1;
    ta.work = h->d_cursors + 12;
    ta.exact_winners = h->d_exact_winners;
    ta.exact_cursor = h->d_cursors + 7;
    ta.reads[0] = d_r1;
    ta.reads[1] = d_r2;
    ta.len[0] = d_l1;
    ta.len[1] = d_l2;
    ta.stride = stride;
    ta.genome = h->d_genome;
    ta.sink = wa.sink;
    ta.band_half = h->band_half;
    ta.counters = h->d_counters;
    ta.p = sa.p;
    ta.work_walk = h->d_cursors + 15;
    ta.flagq = nullptr;
    ta.walk_meta = nullptr;
    ta.pair_codes = nullptr;
    {
      const int trc = dispatch_trace_int(h, ta, max_len);
      if (trc) return trc;
    }
    CK(cudaEventRecord(ev[6], h->stream));
    wa.winners = h->d_exact_winners;
    wa.n_items = h->d_cursors + 7;
    wa.work = h->d_cursors + 13;
  } else {
    wa.n_items = h->d_cursors + 1;
    wa.work = h->d_cursors + 13;
    CK(cudaEventRecord(ev[5], h->stream));
    CK(cudaEventRecord(ev[6], h->stream));
  }
  dispatch_sw<2>(h, wa, max_len);  // exact traceback, decision band in shared memory
  wa.winners = h->d_oob_winners;   // walks that left the band (long indels): every lane's decisions in global memory
  wa.n_items = h->d_cursors + 8;
  wa.work = h->d_cursors + 14;
  dispatch_sw<1>(h, wa, max_len);
  if (forked) CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  CK(cudaEventRecord(ev[4], h->stream));
  CK(cudaGetLastError());
  return PEMAP_OK;
}

int account_chunk(pemap_ctx* h, int n, bool paired, cudaEvent_t* ev) {
  float ms;
  CK(cudaEventSynchronize(ev[4]));
  CK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
  h->stats.ms_seed += ms;
  CK(cudaEventElapsedTime(&ms, ev[1], ev[2]));
  h->stats.ms_sw += ms;
  CK(cudaEventElapsedTime(&ms, ev[2], ev[3]));
  h->stats.ms_select += ms;
  CK(cudaEventElapsedTime(&ms, ev[3], ev[4]));
  h->stats.ms_traceback += ms;
  CK(cudaEventElapsedTime(&ms, ev[3], ev[5]));
  h->stats.ms_tb_diag += ms;
  CK(cudaEventElapsedTime(&ms, ev[5], ev[6]));
  h->stats.ms_tb_int += ms;
  CK(cudaEventElapsedTime(&ms, ev[6], ev[4]));
  h->stats.ms_tb_fp64 += ms;
  CK(cudaEventElapsedTime(&ms, ev[0], ev[4]));
  h->stats.ms_total += ms;
  h->stats.reads += (uint64_t)n * (paired ? 2 : 1);
  return PEMAP_OK;
}

int retain_chunk(pemap_ctx* h, int n, int first, bool paired) {
  if (h->keep & PEMAP_KEEP_DETAIL) {
    std::vector<int32_t> best(2 * (size_t)n), orient(2 * (size_t)n);
    std::vector<double> score(2 * (size_t)n);
    std::vector<uint32_t> cn(2 * (size_t)n);
    CK(cudaMemcpyAsync(best.data(), h->d_det_best, best.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(orient.data(), h->d_det_orient, orient.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(score.data(), h->d_det_score, score.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(cn.data(), h->d_cand_n, cn.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; i++) {
      pemap_detail& d = h->detail[(size_t)first + i];
      d.hits1 = (int32_t)cn[2 * i];
      d.hits2 = paired ? (int32_t)cn[2 * i + 1] : 0;
      d.best1 = best[2 * i];
      d.best2 = best[2 * i + 1];
      d.orient1 = orient[2 * i];
      d.orient2 = orient[2 * i + 1];
      d.score1 = score[2 * i];
      d.score2 = score[2 * i + 1];
    }
  }
  if (h->keep & PEMAP_KEEP_CANDIDATES) {
    uint32_t cur[2];
    CK(cudaMemcpyAsync(cur, h->d_cursors, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<pm::Task> tasks(cur[0]);
    std::vector<uint32_t> cb(2 * (size_t)n), cn(2 * (size_t)n);
    if (cur[0]) CK(cudaMemcpy(tasks.data(), h->d_tasks, tasks.size() * sizeof(pm::Task), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cb.data(), h->d_cand_base, cb.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cn.data(), h->d_cand_n, cn.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
      for (int m = 0; m < (paired ? 2 : 1); m++) {
        const size_t rm = 2 * ((size_t)first + i) + m;
        h->cand_base_h[rm] = (uint32_t)h->cand_spot_h.size();
        h->cand_n_h[rm] = cn[2 * i + m];
        for (uint32_t q = 0; q < cn[2 * i + m]; q++) {
          const pm::Task& t = tasks[cb[2 * i + m] + q];
          h->cand_spot_h.push_back(t.spot);
          h->cand_orient_h.push_back((int8_t)(t.rm >> 31));
        }
      }
  }
  return PEMAP_OK;
}

int fetch_counters(pemap_ctx* h) {
  pm::SeedCounters c;
  CK(cudaMemcpy(&c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
  h->stats.lookups = c.lookups;
  h->stats.mer_positions = c.mer_positions;
  h->stats.candidates = c.candidates;
  h->stats.sw_cells = c.sw_cells;
  h->stats.tb_cells = c.tb_cells;
  h->stats.replayed = c.replayed;
  h->stats.diag_traced = c.diag_traced;
  h->stats.exact_traced = c.exact_traced;
  h->stats.tb_cells_int = c.tb_cells_int;
  h->stats.sw_cells_certified = c.sw_cells_certified;
  return PEMAP_OK;
}

void begin_batch(pemap_ctx* h, int n) {
  if (h->keep & PEMAP_KEEP_DETAIL) h->detail.assign((size_t)n, pemap_detail{});
  if (h->keep & PEMAP_KEEP_CANDIDATES) {
    h->cand_base_h.assign(2 * (size_t)n, 0);
    h->cand_n_h.assign(2 * (size_t)n, 0);
    h->cand_spot_h.clear();
    h->cand_orient_h.clear();
  }
}

bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// Move what the traceback kernels appended so far to the host and rewind the device cursor (the reference mallocs
// every insertion string, pemapper.c:1871-1904; here the append buffer is bounded and spills to host memory).
int drain_insertions(pemap_ctx* h) {
  {
    int rc = sync_lanes(h);
    if (rc) return rc;
  }
  unsigned long long used = 0;
  CK(cudaMemcpyAsync(&used, h->d_ins_cursor, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (used > h->ins_cap)
    return fail(h, PEMAP_ERR_NOMEM, "insertion buffer overflow inside one chunk of reads (raise PEMAP_INS_MB or lower PEMAP_CHUNK)");
  if (used) {
    const size_t at = h->ins_raw.size();
    h->ins_raw.resize(at + used);
    CK(cudaMemcpyAsync(h->ins_raw.data() + at, h->d_ins, used, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemsetAsync(h->d_ins_cursor, 0, 8, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  h->want_drain = false;
  return PEMAP_OK;
}

// after a chunk: fail at once when the append buffer overflowed, ask for a drain past the high-water mark
int check_insertion_fill(pemap_ctx* h, unsigned long long used) {
  if (used > h->ins_cap)
    return fail(h, PEMAP_ERR_NOMEM, "insertion buffer overflow inside one chunk of reads (raise PEMAP_INS_MB or lower PEMAP_CHUNK)");
  if (used > h->ins_cap / 2) h->want_drain = true;
  return PEMAP_OK;
}

// scope guard of map_host: whatever path leaves the function, no slot stays marked pending (a stale slot would copy
// an old chunk's results into the next caller's arrays) and nothing of this call is still running
struct SlotGuard {
  pemap_ctx* h;
  ~SlotGuard() {
    bool any = false;
    for (auto& sl : h->slots) any = any || sl.pending;
    if (any) {
      cudaStreamSynchronize(h->s_h2d);
      sync_lanes(h);
      cudaStreamSynchronize(h->s_d2h);
      for (auto& sl : h->slots) sl.pending = false;
    }
    use_lane(h, 0);
  }
};

// wait for a chunk's results, hand them to the caller and book its stage times
int finish_slot(pemap_ctx* h, pemap_ctx::Slot& sl, uint32_t* m1, uint32_t* m2, int* mapping_type) {
  CK(cudaEventSynchronize(sl.ev_d2h));
  {
    int rc = check_insertion_fill(h, h->h_ins_used[&sl - h->slots]);
    if (rc) return rc;
  }
  if (!sl.direct) {
    memcpy(m1 + sl.first, sl.h_m1, (size_t)sl.n * 4);
    memcpy(m2 + sl.first, sl.h_m2, (size_t)sl.n * 4);
    memcpy(mapping_type + sl.first, sl.h_type, (size_t)sl.n * 4);
  }
  sl.pending = false;
  return account_chunk(h, sl.n, sl.paired, sl.ev);
}

// rows: reads as (n x stride) matrices on the host, or, when ptrs != nullptr, as arrays of pointers
// packed_max_len > 0: rows1 / rows2 are 2-bit packed rows (pemap_pack_read) laid out for that batch maximum
int map_host(pemap_ctx* h, int n, const char* rows1, const char* const* ptr1, const int* len1, const char* rows2,
             const char* const* ptr2, const int* len2, int stride, uint32_t* m1, uint32_t* m2, int* mapping_type,
             int packed_max_len = 0) {
  if (!h) return PEMAP_ERR_ARG;
  if (h->index_only) return fail(h, PEMAP_ERR_ARG, "handle was opened with PEMAP_INDEX_ONLY=1");
  if (n < 0 || (!rows1 && !ptr1) || !len1 || !m1 || !m2 || !mapping_type) return fail(h, PEMAP_ERR_ARG, "NULL argument");
  const bool paired = h->params.pair_flag != 0;
  if (paired && ((!rows2 && !ptr2) || !len2)) return fail(h, PEMAP_ERR_ARG, "pair_flag set but read2/len2 is NULL");
  CK(cudaSetDevice(h->device));
  // every length is checked before the first chunk is submitted: no early return with work in flight
  for (int m = 0; m < (paired ? 2 : 1); m++) {
    const int* len = m ? len2 : len1;
    for (int i = 0; i < n; i++)
      if (len[i] < 0 || len[i] > PM_DP_MAX - 22)
        return fail(h, PEMAP_ERR_ARG, "read longer than 298 bases (reference DP buffers are 300x300)");
  }
  if (rows1 && !ptr1 && stride > h->stride_cap && is_pinned(rows1))
    return fail(h, PEMAP_ERR_ARG, "row stride larger than 320 bytes");
  const bool packed = packed_max_len > 0;
  if (packed) {
    if (packed_max_len > PM_DP_MAX - 22 || stride != (int)pemap_packed_stride(packed_max_len))
      return fail(h, PEMAP_ERR_ARG, "packed rows: stride must be pemap_packed_stride(max_len)");
    if (h->seed_legacy) return fail(h, PEMAP_ERR_UNSUPPORTED, "PEMAP_SEED=legacy reads ASCII rows only");
    const size_t need = (size_t)h->chunk * (size_t)pemap_packed_stride(PM_DP_MAX - 22);
    if (h->packed_cap < need) {
      for (auto& sl : h->d_packed)
        for (auto& p : sl) {
          if (p) cudaFree(p);
          p = nullptr;
          CK(cudaMalloc(&p, need));
        }
      h->packed_cap = need;
    }
  }
  begin_batch(h, n);
  SlotGuard guard{h};
  {
    int rc = fork_lanes(h);
    if (rc) return rc;
  }
  const bool direct = rows1 && is_pinned(rows1) && is_pinned(len1) && (!paired || (is_pinned(rows2) && is_pinned(len2))) &&
                      is_pinned(m1) && is_pinned(m2) && is_pinned(mapping_type);
  const bool pipelined = h->keep == 0;  // the inspection modes read shared scratch after every chunk
  int chunk_no = 0;
  for (int first = 0; first < n; first += h->chunk, chunk_no++) {
    const int cn = std::min(h->chunk, n - first);
    pemap_ctx::Slot& sl = h->slots[chunk_no & 1];
    if (sl.pending) {  // the chunk that used this slot (and this lane) two iterations ago
      int rc = finish_slot(h, sl, m1, m2, mapping_type);
      if (rc) return rc;
    }
    use_lane(h, pipelined && h->lanes_->n > 1 ? (chunk_no & 1) : 0);
    if (h->want_drain) {  // the append buffer passed its high-water mark: spill it before more is appended
      int rc = drain_insertions(h);
      if (rc) return rc;
    }
    int max_len = 0, min_len = 1 << 30;
    for (int m = 0; m < (paired ? 2 : 1); m++) {
      const int* len = (m ? len2 : len1) + first;
      for (int i = 0; i < cn; i++) {
        max_len = std::max(max_len, len[i]);
        min_len = std::min(min_len, len[i]);
      }
    }
    int dstride = stride;
    const char* src_rows[2] = {rows1 ? rows1 + (size_t)first * stride : nullptr, rows2 ? rows2 + (size_t)first * stride : nullptr};
    if (packed) {
      // packed rows: copied as they are (staged through the pinned buffers when the caller's are pageable), then
      // unpacked on the device into the ASCII rows the DP kernels read; the seed kernel reads the packed words
      if (max_len > packed_max_len) return fail(h, PEMAP_ERR_ARG, "a read is longer than the packed rows' max_len");
      const int code_words = (packed_max_len + 15) / 16, mask_words = (packed_max_len + 31) / 32;
      dstride = (packed_max_len + 15) & ~15;
      for (int m = 0; m < (paired ? 2 : 1); m++) {
        const int* len = (m ? len2 : len1) + first;
        const void* src = src_rows[m];
        const int* lsrc = len;
        if (!direct) {
          memcpy(sl.h_reads[m], src_rows[m], (size_t)cn * stride);
          memcpy(sl.h_len[m], len, (size_t)cn * sizeof(int));
          src = sl.h_reads[m];
          lsrc = sl.h_len[m];
        }
        unsigned char* dp = h->d_packed[chunk_no & 1][m];
        CK(cudaMemcpyAsync(dp, src, (size_t)cn * stride, cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaMemcpyAsync(sl.d_len[m], lsrc, (size_t)cn * sizeof(int), cudaMemcpyHostToDevice, h->s_h2d));
        const long long threads = (long long)cn * (dstride / 16);
        pm::k_unpack_reads<<<(unsigned)((threads + 255) / 256), 256, 0, h->s_h2d>>>(dp, stride, code_words, sl.d_len[m], cn,
                                                                                  sl.d_reads[m], dstride);
        h->stats.launches++;
        h->cur_packed.rows[m] = dp;
      }
      if (!paired) h->cur_packed.rows[1] = nullptr;
      h->cur_packed.stride = stride;
      h->cur_packed.code_words = code_words;
      h->cur_packed.mask_words = mask_words;
    } else if (!direct || ptr1) {  // stage through the library's pinned buffers, packed at a 16-byte multiple
      dstride = (max_len + 15) & ~15;
      if (dstride < 16) dstride = 16;
      for (int m = 0; m < (paired ? 2 : 1); m++) {
        const int* len = (m ? len2 : len1) + first;
        char* dst = sl.h_reads[m];
        if (m ? ptr2 != nullptr : ptr1 != nullptr) {
          const char* const* pp = (m ? ptr2 : ptr1) + first;
          for (int i = 0; i < cn; i++) memcpy(dst + (size_t)i * dstride, pp[i], (size_t)len[i]);
        } else {
          const char* rows = src_rows[m];
          for (int i = 0; i < cn; i++) memcpy(dst + (size_t)i * dstride, rows + (size_t)i * stride, (size_t)len[i]);
        }
        memcpy(sl.h_len[m], len, (size_t)cn * sizeof(int));
        src_rows[m] = dst;
      }
    }
    if (dstride > h->stride_cap) return fail(h, PEMAP_ERR_ARG, "row stride larger than 320 bytes");
    for (int m = 0; m < (paired ? 2 : 1) && !packed; m++) {
      const int* lsrc = (!direct || ptr1) ? sl.h_len[m] : (m ? len2 : len1) + first;
      CK(cudaMemcpyAsync(sl.d_reads[m], src_rows[m], (size_t)cn * dstride, cudaMemcpyHostToDevice, h->s_h2d));
      CK(cudaMemcpyAsync(sl.d_len[m], lsrc, (size_t)cn * sizeof(int), cudaMemcpyHostToDevice, h->s_h2d));
    }
    CK(cudaEventRecord(sl.ev_h2d, h->s_h2d));
    CK(cudaStreamWaitEvent(h->stream, sl.ev_h2d, 0));
    int rc = run_chunk(h, cn, sl.d_reads[0], sl.d_len[0], paired ? sl.d_reads[1] : nullptr, paired ? sl.d_len[1] : nullptr,
                       dstride, max_len, min_len == max_len ? max_len : 0, sl.d_m1, sl.d_m2, sl.d_type, sl.ev);
    h->cur_packed.rows[0] = h->cur_packed.rows[1] = nullptr;
    if (rc) return rc;
    CK(cudaStreamWaitEvent(h->s_d2h, sl.ev[4], 0));
    uint32_t* o1 = direct ? m1 + first : sl.h_m1;
    uint32_t* o2 = direct ? m2 + first : sl.h_m2;
    int* ot = direct ? mapping_type + first : sl.h_type;
    CK(cudaMemcpyAsync(o1, sl.d_m1, (size_t)cn * 4, cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaMemcpyAsync(o2, sl.d_m2, (size_t)cn * 4, cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaMemcpyAsync(ot, sl.d_type, (size_t)cn * 4, cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaMemcpyAsync(h->h_ins_used + (chunk_no & 1), h->d_ins_cursor, 8, cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaEventRecord(sl.ev_d2h, h->s_d2h));
    sl.pending = true;
    sl.n = cn;
    sl.first = first;
    sl.paired = paired;
    sl.direct = direct;
    if (!pipelined) {
      rc = finish_slot(h, sl, m1, m2, mapping_type);
      if (rc) return rc;
      rc = retain_chunk(h, cn, first, paired);
      if (rc) return rc;
    }
  }
  for (int k = 0; k < 2; k++) {  // drain in submission order
    pemap_ctx::Slot& sl = h->slots[(chunk_no + k) & 1];
    if (sl.pending) {
      int rc = finish_slot(h, sl, m1, m2, mapping_type);
      if (rc) return rc;
    }
  }
  return join_lanes(h);
}

}  // namespace

namespace {
// BASELINE configs[3]: k_sw_i16 with windows of up to PEMAP_SW_MAX_WINDOW rows; uniform read lengths get the
// instantiation whose last read column is a compile-time constant
template <int G, int WD, int CMM>
struct LongDispatch {
  static void go(pemap_ctx* h, const pm::SwIntArgs& a, int cmm) {
    if (cmm == CMM) {
      static int grids[kMaxDev] = {};
      int& grid = grids[h->device % kMaxDev];
      if (!grid) grid = one_wave_grid(h, pm::k_sw_i16<G, WD, CMM, PEMAP_SW_MAX_WINDOW>, 128, 0);
      pm::k_sw_i16<G, WD, CMM, PEMAP_SW_MAX_WINDOW><<<grid, 128, 0, h->stream>>>(a);
    } else {
      LongDispatch<G, WD, CMM - 1>::go(h, a, cmm);
    }
  }
};
template <int G, int WD>
struct LongDispatch<G, WD, -2> {
  static void go(pemap_ctx*, const pm::SwIntArgs&, int) {}
};
}  // namespace

// ------------------------------------------------------------------------------------------------- C-ABI

extern "C" {

const char* pemap_version(void) { return PEMAP_VERSION; }

void pemap_default_params(pemap_params* p) {
  if (!p) return;
  p->idepth = 16;
  p->max_hits = 200;
  p->too_many_spots = 100;
  p->min_align = 0.9;
  p->match_bonus = 1.0;
  p->is_bisulfite = 0;
  p->pair_flag = 0;
  p->min_dist = 0;
  p->max_dist = 500;
  p->misalign_slop = 10;
}

const char* pemap_last_error(pemap_t* h) { return h ? h->err.c_str() : "NULL handle"; }

int pemap_init(pemap_t** out, const pemap_index* ix, const pemap_params* p, int device) {
  return pemap_init_streamed(out, ix, nullptr, nullptr, p, device);
}

int pemap_init_streamed(pemap_t** out, const pemap_index* ix, pemap_fill_cb next_idx_bytes, void* cb_ctx, const pemap_params* p,
                        int device) {
  if (!out) return PEMAP_ERR_ARG;
  pemap_ctx* h = new pemap_ctx();
  *out = h;
  int rc = check_params(h, p);
  if (rc) return rc;
  if (!ix || (!ix->pos_index && !next_idx_bytes) || !ix->mers || !ix->genome || !ix->contig_starts)
    return fail(h, PEMAP_ERR_ARG, "index has NULL members");
  rc = check_contigs(h, ix->no_contigs);
  if (rc) return rc;
  h->params = *p;
  rc = open_device(h, device);
  if (rc) return rc;
  h->n_contigs = ix->no_contigs;
  h->genome_size = ix->genome_size;
  h->n_mers = ix->n_mers;
  const size_t idx_words = ((size_t)1 << 32) + 1;
  CK(cudaMalloc(&h->d_pos_index, idx_words * 4));
  if (next_idx_bytes) {
    // init_index_buffer (2129-2149) without the 16 GiB host table: the caller inflates the .idx stream chunk by chunk
    // straight into page-locked staging, every chunk is on its way to the GPU while the next one is inflated
    const size_t chunk = (size_t)64 << 20, total = idx_words * 4;
    char* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; k++) {
      CK(cudaHostAlloc(&stage[k], chunk, cudaHostAllocDefault));
      CK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    }
    int cb_rc = 0;
    size_t at = 0;
    for (int k = 0; at < total && !cb_rc; k ^= 1) {
      const size_t nb = std::min(chunk, total - at);
      CK(cudaEventSynchronize(done[k]));  // the copy that used this buffer two chunks ago
      cb_rc = next_idx_bytes(cb_ctx, stage[k], nb);
      if (cb_rc) break;
      CK(cudaMemcpyAsync(reinterpret_cast<char*>(h->d_pos_index) + at, stage[k], nb, cudaMemcpyHostToDevice, h->s_h2d));
      CK(cudaEventRecord(done[k], h->s_h2d));
      at += nb;
    }
    CK(cudaStreamSynchronize(h->s_h2d));
    for (int k = 0; k < 2; k++) {
      cudaFreeHost(stage[k]);
      cudaEventDestroy(done[k]);
    }
    if (cb_rc) return fail(h, PEMAP_ERR_ARG, "pemap_init_streamed: the .idx callback failed (short file?)");
  } else {
    CK(cudaMemcpy(h->d_pos_index, ix->pos_index, idx_words * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&h->d_mers, (size_t)(h->n_mers + 4) * 4));
  CK(cudaMemcpy(h->d_mers, ix->mers, (size_t)h->n_mers * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_genome, h->genome_size + 64));
  CK(cudaMemset(h->d_genome, 'N', h->genome_size + 64));
  CK(cudaMemcpy(h->d_genome, ix->genome, h->genome_size, cudaMemcpyHostToDevice));
  rc = upload_cstart(h, ix->contig_starts, ix->no_contigs);
  if (rc) return rc;
  read_index_env(h);
  if (!h->index_only && !h->seed_legacy) {
    uint32_t* code_of = nullptr;
    CK(cudaMalloc(&code_of, (size_t)(h->n_mers + 1) * 4));
    pm::k_rbi_expand_codes<<<1u << 24, 256, 0, h->stream>>>(h->d_pos_index, code_of, h->n_mers);
    rc = build_rbi(h, code_of, h->d_mers, h->n_mers);
    cudaFree(code_of);
    if (rc) return rc;
    drop_file_index_if_large(h);
  }
  return finish_init(h);
}

int pemap_init_from_genome(pemap_t** out, const char* genome, const int64_t* contig_len, int n_contigs,
                           const pemap_params* p, int device) {
  if (!out) return PEMAP_ERR_ARG;
  pemap_ctx* h = new pemap_ctx();
  *out = h;
  int rc = check_params(h, p);
  if (rc) return rc;
  if (!genome || !contig_len) return fail(h, PEMAP_ERR_ARG, "NULL genome");
  rc = check_contigs(h, n_contigs);
  if (rc) return rc;
  h->params = *p;
  rc = open_device(h, device);
  if (rc) return rc;
  h->n_contigs = n_contigs;
  std::vector<uint64_t> real_start((size_t)n_contigs + 1, 0);
  std::vector<uint32_t> cs((size_t)n_contigs + 1, 0);
  for (int i = 0; i < n_contigs; i++) {
    if (contig_len[i] < 16) return fail(h, PEMAP_ERR_ARG, "contig shorter than 16 bases");
    real_start[i + 1] = real_start[i] + (uint64_t)contig_len[i];
    cs[i + 1] = cs[i] + (uint32_t)(contig_len[i] - 15);  // .sdx stores len-15 (index_genome_whole.c:213-216, 316)
  }
  h->genome_size = real_start[n_contigs];
  if (h->genome_size >= 0xFFFFFFFFull) return fail(h, PEMAP_ERR_UNSUPPORTED, "genome does not fit 32-bit coordinates");
  const uint64_t gs = h->genome_size;
  CK(cudaMalloc(&h->d_genome, gs + 64));
  CK(cudaMemset(h->d_genome, 'N', gs + 64));
  CK(cudaMemcpy(h->d_genome, genome, gs, cudaMemcpyHostToDevice));
  rc = upload_cstart(h, cs.data(), n_contigs);
  if (rc) return rc;

  // ---- k-mers of every start position, compaction of the valid ones, stable sort by k-mer
  uint64_t* d_real = nullptr;
  uint32_t *d_key = nullptr, *d_val = nullptr, *d_key2 = nullptr, *d_val2 = nullptr;
  unsigned char* d_flag = nullptr;
  unsigned long long* d_nsel = nullptr;
  CK(cudaMalloc(&d_real, real_start.size() * 8));
  CK(cudaMemcpy(d_real, real_start.data(), real_start.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_key, gs * 4));
  CK(cudaMalloc(&d_val, gs * 4));
  CK(cudaMalloc(&d_key2, gs * 4));
  CK(cudaMalloc(&d_val2, gs * 4));
  CK(cudaMalloc(&d_flag, gs));
  CK(cudaMalloc(&d_nsel, 8));
  pm::k_index_kmers<<<(unsigned)((gs + 255) / 256), 256, 0, h->stream>>>(h->d_genome, gs, d_real, n_contigs,
                                                                         p->is_bisulfite, d_key, d_val, d_flag);
  void* tmp = nullptr;
  size_t tmp_bytes = 0, need = 0;
  const long long n_items = (long long)gs;  // CUB takes 64-bit item counts: genomes up to the 2^32 coordinate limit
  cub::DeviceSelect::Flagged(nullptr, need, d_key, d_flag, d_key2, d_nsel, n_items, h->stream);
  tmp_bytes = need;
  cub::DeviceRadixSort::SortPairs(nullptr, need, d_key2, d_key, d_val2, d_val, n_items, 0, 32, h->stream);
  tmp_bytes = std::max(tmp_bytes, need);
  CK(cudaMalloc(&tmp, tmp_bytes));
  need = tmp_bytes;
  CK(cub::DeviceSelect::Flagged(tmp, need, d_key, d_flag, d_key2, d_nsel, n_items, h->stream));
  need = tmp_bytes;
  CK(cub::DeviceSelect::Flagged(tmp, need, d_val, d_flag, d_val2, d_nsel, n_items, h->stream));
  unsigned long long nsel = 0;
  CK(cudaMemcpyAsync(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n_mers = nsel;
  need = tmp_bytes;
  CK(cub::DeviceRadixSort::SortPairs(tmp, need, d_key2, d_key, d_val2, d_val, (long long)nsel, 0, 32, h->stream));
  // d_key = sorted k-mers, d_val = positions grouped by k-mer (= .mdx)
  const size_t idx_words = ((size_t)1 << 32) + 1;
  CK(cudaMalloc(&h->d_pos_index, idx_words * 4));
  const uint64_t step = 1ull << 30;
  for (uint64_t first = 0; first < idx_words; first += step) {
    const uint64_t cnt = std::min<uint64_t>(step, idx_words - first);
    pm::k_index_prefix<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(d_key, nsel, h->d_pos_index, first, cnt);
  }
  CK(cudaMalloc(&h->d_mers, (size_t)(nsel + 4) * 4));
  CK(cudaMemcpyAsync(h->d_mers, d_val, (size_t)nsel * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  cudaFree(tmp);
  cudaFree(d_real);
  cudaFree(d_val);
  cudaFree(d_key2);
  cudaFree(d_val2);
  cudaFree(d_flag);
  cudaFree(d_nsel);
  read_index_env(h);
  if (!h->index_only && !h->seed_legacy) {
    rc = build_rbi(h, d_key, h->d_mers, nsel);
    cudaFree(d_key);
    if (rc) return rc;
    drop_file_index_if_large(h);
  } else {
    cudaFree(d_key);
  }
  return finish_init(h);
}

int pemap_set_params(pemap_t* h, const pemap_params* p) {
  if (!h) return PEMAP_ERR_ARG;
  int rc = check_params(h, p);
  if (rc) return rc;
  if (p->is_bisulfite != h->params.is_bisulfite)
    return fail(h, PEMAP_ERR_UNSUPPORTED, "is_bisulfite is baked into the index and cannot change");
  const bool new_scores = p->match_bonus != h->params.match_bonus;
  h->params = *p;
  fill_dev_params(h);
  if (new_scores) return upload_border(h);
  return PEMAP_OK;
}

int pemap_keep(pemap_t* h, int flags) {
  if (!h) return PEMAP_ERR_ARG;
  h->keep = flags;
  return PEMAP_OK;
}

int pemap_map_batch(pemap_t* h, int n, const char* const* read1, const int* len1, const char* const* read2,
                    const int* len2, uint32_t* m1, uint32_t* m2, int* mapping_type) {
  return map_host(h, n, nullptr, read1, len1, nullptr, read2, len2, 0, m1, m2, mapping_type);
}

int pemap_map_batch_rows(pemap_t* h, int n, const char* reads1, const int* len1, const char* reads2, const int* len2,
                         int stride, uint32_t* m1, uint32_t* m2, int* mapping_type) {
  return map_host(h, n, reads1, nullptr, len1, reads2, nullptr, len2, stride, m1, m2, mapping_type);
}

size_t pemap_packed_stride(int max_len) {
  if (max_len < 1) max_len = 1;
  const size_t bytes = 4u * (size_t)((max_len + 15) / 16) + 4u * (size_t)((max_len + 31) / 32);
  return (bytes + 15u) & ~(size_t)15u;
}

int pemap_pack_read(const char* read, int len, int max_len, void* dst) {
  if (!read || !dst || len < 0 || len > max_len) return PEMAP_ERR_ARG;
  uint32_t* w = static_cast<uint32_t*>(dst);
  const int code_words = (max_len + 15) / 16;
  const size_t words = pemap_packed_stride(max_len) / 4;
  for (size_t i = 0; i < words; i++) w[i] = 0u;
  for (int i = 0; i < len; i++) {
    uint32_t c;
    switch (read[i]) {
      case 'A': c = 0; break;
      case 'C': c = 1; break;
      case 'G': c = 2; break;
      case 'T': c = 3; break;
      case 'N': c = 0; w[code_words + (i >> 5)] |= 1u << (i & 31); break;
      default: return PEMAP_ERR_UNSUPPORTED;  // lower case, IUPAC codes ...: such a read goes through the ASCII entry points
    }
    w[i >> 4] |= c << (30 - 2 * (i & 15));
  }
  return PEMAP_OK;
}

int pemap_map_batch_packed(pemap_t* h, int n, const void* packed1, const int* len1, const void* packed2, const int* len2,
                           int max_len, uint32_t* m1, uint32_t* m2, int* mapping_type) {
  return map_host(h, n, static_cast<const char*>(packed1), nullptr, len1, static_cast<const char*>(packed2), nullptr, len2,
                  (int)pemap_packed_stride(max_len), m1, m2, mapping_type, max_len);
}

int pemap_map_batch_device(pemap_t* h, int n, const char* d_reads1, const int* d_len1, const char* d_reads2,
                           const int* d_len2, int stride, int max_len, uint32_t* d_m1, uint32_t* d_m2,
                           int* d_mapping_type) {
  if (!h) return PEMAP_ERR_ARG;
  if (h->index_only) return fail(h, PEMAP_ERR_ARG, "handle was opened with PEMAP_INDEX_ONLY=1");
  if (n < 0 || !d_reads1 || !d_len1 || !d_m1 || !d_m2 || !d_mapping_type) return fail(h, PEMAP_ERR_ARG, "NULL argument");
  if (max_len < 16 || max_len > PM_DP_MAX - 22) return fail(h, PEMAP_ERR_ARG, "max_len out of range");
  const bool paired = h->params.pair_flag != 0;
  if (paired && (!d_reads2 || !d_len2)) return fail(h, PEMAP_ERR_ARG, "pair_flag set but read2/len2 is NULL");
  CK(cudaSetDevice(h->device));
  begin_batch(h, n);
  // one reduction over the lengths: uniform-length batches take the specialised integer kernel
  int uniform_len = 0;
  {
    unsigned mm[2] = {0xFFFFFFFFu, 0u};
    CK(cudaMemcpyAsync(h->d_cursors + 4, mm, 8, cudaMemcpyHostToDevice, h->stream));
    if (n > 0) {
      k_len_range<<<h->sm_count * 4, 256, 0, h->stream>>>(d_len1, paired ? d_len2 : nullptr, n, h->d_cursors + 4);
      h->stats.launches++;
    }
    CK(cudaMemcpyAsync(mm, h->d_cursors + 4, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n > 0 && mm[0] == mm[1]) uniform_len = (int)mm[0];
    if (n > 0 && (int)mm[1] > max_len) return fail(h, PEMAP_ERR_ARG, "a read is longer than max_len");
  }
  // chunks alternate between the two lanes; a lane's previous chunk is booked (stage times, insertion fill) before the
  // lane is reused.  The inspection modes read a chunk's scratch right after it: they run one chunk at a time.
  const bool two = h->lanes_->n > 1 && h->keep == 0;
  {
    int rc = fork_lanes(h);
    if (rc) return rc;
  }
  struct Pending {
    bool on = false;
    int cn = 0;
  } pend[2];
  auto book = [&](int lane) -> int {
    if (!pend[lane].on) return PEMAP_OK;
    pend[lane].on = false;
    int rc = account_chunk(h, pend[lane].cn, paired, h->slots[lane].ev);
    if (rc) return rc;
    return PEMAP_OK;
  };
  int chunk_no = 0;
  for (int first = 0; first < n; first += h->chunk, chunk_no++) {
    const int cn = std::min(h->chunk, n - first);
    const int lane = two ? (chunk_no & 1) : 0;
    int rc = book(lane);
    if (rc) return rc;
    if (h->want_drain) {
      rc = book(lane ^ 1);
      if (rc) return rc;
      rc = drain_insertions(h);
      if (rc) return rc;
    }
    use_lane(h, lane);
    rc = run_chunk(h, cn, d_reads1 + (size_t)first * stride, d_len1 + first,
                   paired ? d_reads2 + (size_t)first * stride : nullptr, paired ? d_len2 + first : nullptr, stride,
                   max_len, uniform_len, d_m1 + first, d_m2 + first, d_mapping_type + first, h->slots[lane].ev);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->h_ins_used + lane, h->d_ins_cursor, 8, cudaMemcpyDeviceToHost, h->stream));
    pend[lane].on = true;
    pend[lane].cn = cn;
    if (!two) {
      rc = book(lane);
      if (rc) return rc;
      CK(cudaStreamSynchronize(h->stream));
      rc = check_insertion_fill(h, h->h_ins_used[lane]);
      if (rc) return rc;
      rc = retain_chunk(h, cn, first, paired);
      if (rc) return rc;
    } else if (chunk_no >= 1) {  // the other lane's chunk was submitted one iteration ago: its fill level is (nearly) known
      CK(cudaEventSynchronize(h->slots[lane ^ 1].ev[4]));
      rc = check_insertion_fill(h, h->h_ins_used[lane ^ 1]);
      if (rc) return rc;
    }
  }
  for (int k = 0; k < 2; k++) {
    int rc = book((chunk_no + k) & 1);
    if (rc) return rc;
  }
  {
    int rc = join_lanes(h);
    if (rc) return rc;
    rc = sync_lanes(h);
    if (rc) return rc;
    rc = check_insertion_fill(h, std::max(h->h_ins_used[0], h->h_ins_used[1]));
    if (rc) return rc;
  }
  return PEMAP_OK;
}

int pemap_get_detail(pemap_t* h, pemap_detail* out, int n) {
  if (!h || !out) return PEMAP_ERR_ARG;
  if (!(h->keep & PEMAP_KEEP_DETAIL)) return fail(h, PEMAP_ERR_ARG, "PEMAP_KEEP_DETAIL not enabled");
  if ((size_t)n > h->detail.size()) return fail(h, PEMAP_ERR_ARG, "n exceeds the last batch");
  memcpy(out, h->detail.data(), (size_t)n * sizeof(pemap_detail));
  return PEMAP_OK;
}

int pemap_get_candidates(pemap_t* h, int i, int mate, uint32_t* spots, int8_t* orients, int cap) {
  if (!h || !spots || !orients) return PEMAP_ERR_ARG;
  if (!(h->keep & PEMAP_KEEP_CANDIDATES)) return fail(h, PEMAP_ERR_ARG, "PEMAP_KEEP_CANDIDATES not enabled");
  const size_t rm = 2 * (size_t)i + (size_t)mate;
  if (i < 0 || rm >= h->cand_n_h.size()) return fail(h, PEMAP_ERR_ARG, "read index out of range");
  const int n = (int)h->cand_n_h[rm];
  for (int q = 0; q < n && q < cap; q++) {
    spots[q] = h->cand_spot_h[h->cand_base_h[rm] + q];
    orients[q] = h->cand_orient_h[h->cand_base_h[rm] + q];
  }
  return n;
}

int pemap_finish_stream(pemap_t* h, pemap_site_cb cb, void* ctx, uint64_t* n_records) {
  if (!h) return PEMAP_ERR_ARG;
  return pemap_finish_stream_range(h, 0, h->genome_size, cb, ctx, n_records);
}

int pemap_finish_stream_range(pemap_t* h, uint64_t site_first, uint64_t site_end, pemap_site_cb cb, void* ctx,
                              uint64_t* n_records) {
  if (!h || !cb) return PEMAP_ERR_ARG;
  if (h->index_only) return fail(h, PEMAP_ERR_ARG, "handle was opened with PEMAP_INDEX_ONLY=1");
  if (site_end > h->genome_size) site_end = h->genome_size;
  if (site_first > site_end) return fail(h, PEMAP_ERR_ARG, "empty or reversed site range");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const uint64_t gs = site_end - site_first;
  const uint64_t tile = (uint64_t)PM_COMPACT_BLOCK * PM_COMPACT_ITEMS;
  if (!h->fin_sites) {  // staging: two device and two pinned host buffers of one window each
    uint64_t w = 1ull << 24;  // 16 M sites = 256 MB of records per buffer
    if (const char* s = getenv("PEMAP_FINISH_SITES")) w = std::max<uint64_t>(tile, strtoull(s, nullptr, 10));
    w = std::min(w, std::max<uint64_t>(h->genome_size, 1));
    w = (w + tile - 1) / tile * tile;
    const size_t n_tiles = (size_t)(w / tile);
    for (int k = 0; k < 2; k++) {
      CK(cudaMalloc(&h->d_fin_rec[k], (size_t)w * sizeof(pm::PileRecord)));
      CK(cudaHostAlloc(&h->h_fin_rec[k], (size_t)w * sizeof(pemap_record), cudaHostAllocDefault));
      CK(cudaEventCreateWithFlags(&h->ev_fin[k], cudaEventDisableTiming));
    }
    CK(cudaMalloc(&h->d_fin_cnt, (n_tiles + 1) * 8));
    CK(cudaMalloc(&h->d_fin_off, 2 * (n_tiles + 1) * 8));
    CK(cudaHostAlloc(&h->h_fin_total, 2 * 8, cudaHostAllocDefault));
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, h->d_fin_cnt, h->d_fin_off, (int)n_tiles + 1, h->stream);
    CK(cudaMalloc(&h->d_fin_tmp, need));
    h->fin_tmp_bytes = need;
    h->fin_sites = w;
  }
  static_assert(sizeof(pm::PileRecord) == sizeof(pemap_record) && sizeof(pemap_record) == 16, "record layout");
  const uint64_t W = h->fin_sites;
  const uint64_t n_win = (gs + W - 1) / W;
  uint64_t total_all = 0;
  uint64_t win_total[2] = {0, 0};
  int cb_rc = 0;
  // window w is compacted on the compute stream and copied out on the D2H stream while the caller consumes window w-1
  for (uint64_t w = 0; w <= n_win; w++) {
    const int slot = (int)(w & 1);
    if (w < n_win) {
      const uint64_t s0 = site_first + w * W, ns = std::min(W, site_end - s0);
      const unsigned n_tiles = (unsigned)((ns + tile - 1) / tile);
      unsigned long long* d_off = h->d_fin_off + (size_t)slot * (W / tile + 1);
      CK(cudaMemsetAsync(h->d_fin_cnt + n_tiles, 0, 8, h->stream));
      pm::k_compact_count<<<n_tiles, PM_COMPACT_BLOCK, 0, h->stream>>>(h->d_counts, s0, ns, h->d_fin_cnt);
      size_t need = h->fin_tmp_bytes;
      CK(cub::DeviceScan::ExclusiveSum(h->d_fin_tmp, need, h->d_fin_cnt, d_off, (int)n_tiles + 1, h->stream));
      pm::k_compact_write<<<n_tiles, PM_COMPACT_BLOCK, 0, h->stream>>>(h->d_counts, s0, ns, d_off, h->d_fin_rec[slot]);
      h->stats.launches += 2;
      CK(cudaMemcpyAsync(h->h_fin_total + slot, d_off + n_tiles, 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      win_total[slot] = h->h_fin_total[slot];
      if (win_total[slot])
        CK(cudaMemcpyAsync(h->h_fin_rec[slot], h->d_fin_rec[slot], (size_t)win_total[slot] * sizeof(pemap_record),
                           cudaMemcpyDeviceToHost, h->s_d2h));
      CK(cudaEventRecord(h->ev_fin[slot], h->s_d2h));
    }
    if (w > 0) {  // hand window w-1 to the caller (its copy overlaps the compaction of window w issued above)
      const int ps = slot ^ 1;
      CK(cudaEventSynchronize(h->ev_fin[ps]));
      if (win_total[ps] && !cb_rc) cb_rc = cb(ctx, h->h_fin_rec[ps], win_total[ps]);
      total_all += win_total[ps];
    }
  }
  CK(cudaGetLastError());
  if (n_records) *n_records = total_all;
  if (cb_rc) return fail(h, PEMAP_ERR_ARG, "pemap_finish_stream: the callback returned non-zero");
  return fetch_counters(h);
}

int pemap_get_insertions(pemap_t* h, const pemap_insertion** ins, uint64_t* n_ins) {
  if (!h) return PEMAP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  // insertion strings: {u32 pos, u32 len, chars padded to 4} records appended by the traceback kernels
  int rc = drain_insertions(h);
  if (rc) return rc;
  const std::vector<unsigned char>& raw = h->ins_raw;
  const size_t used = raw.size();
  h->ins.clear();
  h->ins_pool.clear();
  std::vector<size_t> offs;
  for (size_t o = 0; o + 8 <= used;) {
    uint32_t pos, len;
    memcpy(&pos, raw.data() + o, 4);
    memcpy(&len, raw.data() + o + 4, 4);
    pemap_insertion r;
    r.pos = pos;
    r.len = len;
    r.seq = nullptr;
    offs.push_back(h->ins_pool.size());
    h->ins_pool.insert(h->ins_pool.end(), raw.begin() + o + 8, raw.begin() + o + 8 + len);
    h->ins_pool.push_back('\0');
    h->ins.push_back(r);
    o += 8 + ((len + 3) & ~3u);
  }
  for (size_t i = 0; i < h->ins.size(); i++) h->ins[i].seq = h->ins_pool.data() + offs[i];
  // group by site (ascending position; order inside a site is not defined by the reference either)
  std::stable_sort(h->ins.begin(), h->ins.end(), [](const pemap_insertion& x, const pemap_insertion& y) {
    if (x.pos != y.pos) return x.pos < y.pos;
    return strcmp(x.seq, y.seq) < 0;
  });
  if (ins) *ins = h->ins.data();
  if (n_ins) *n_ins = h->ins.size();
  return PEMAP_OK;
}

static int collect_records(void* ctx, const pemap_record* rec, uint64_t n) {
  std::vector<pemap_record>* v = static_cast<std::vector<pemap_record>*>(ctx);
  v->insert(v->end(), rec, rec + n);
  return 0;
}

int pemap_finish(pemap_t* h, const pemap_record** records, uint64_t* n_records, const pemap_insertion** ins,
                 uint64_t* n_ins) {
  if (!h || !records || !n_records) return PEMAP_ERR_ARG;
  h->records.clear();
  uint64_t total = 0;
  int rc = pemap_finish_stream(h, collect_records, &h->records, &total);
  if (rc) return rc;
  *records = h->records.data();
  *n_records = total;
  return pemap_get_insertions(h, ins, n_ins);
}

int pemap_reset_counts(pemap_t* h) {
  if (!h) return PEMAP_ERR_ARG;
  if (h->index_only) return PEMAP_OK;
  CK(cudaSetDevice(h->device));
  // stream-ordered: every work stream of the handle is non-blocking, so a memset on the legacy stream would not be
  // ordered against the next batch's kernels (the aux and copy streams only ever run between events of h->stream)
  CK(cudaMemsetAsync(h->d_counts, 0, (size_t)h->genome_size * 6 * 4, h->stream));
  CK(cudaMemsetAsync(h->d_ins_cursor, 0, 8, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->records.clear();
  h->records.shrink_to_fit();
  h->ins.clear();
  h->ins_pool.clear();
  h->ins_raw.clear();
  h->want_drain = false;
  return PEMAP_OK;
}

int pemap_counts_device(pemap_t* h, void** d_counts, uint64_t* n_words) {
  if (!h || !d_counts || !n_words) return PEMAP_ERR_ARG;
  *d_counts = h->d_counts;
  *n_words = h->genome_size * 6;
  return PEMAP_OK;
}

int pemap_reduce_counts_peer(pemap_t* h, pemap_t* src) {
  if (!h || !src) return PEMAP_ERR_ARG;
  if (h->genome_size != src->genome_size) return fail(h, PEMAP_ERR_ARG, "handles index different genomes");
  CK(cudaSetDevice(src->device));
  CK(cudaStreamSynchronize(src->stream));
  CK(cudaSetDevice(h->device));
  const uint64_t n_words = h->genome_size * 6;
  if (h->device != src->device) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, h->device, src->device));
    if (!can) return fail(h, PEMAP_ERR_UNSUPPORTED, "no peer access between the two devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(h, PEMAP_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    cudaGetLastError();
  }
  pm::k_add_peer_counts<<<h->sm_count * 8, 256, 0, h->stream>>>(h->d_counts, src->d_counts, n_words);
  h->stats.launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return PEMAP_OK;
}

namespace {
// this rank's slice of the genome when the counters are summed slice-wise over n_ranks GPUs: equal shares cut at
// compaction tiles, so that rank r's records follow rank r-1's in ascending coordinate
void slice_of(const pemap_ctx* h, int n_ranks, int rank, uint64_t* s0, uint64_t* s1) {
  const uint64_t tile = (uint64_t)PM_COMPACT_BLOCK * PM_COMPACT_ITEMS;
  const uint64_t tiles = (h->genome_size + tile - 1) / tile, per = (tiles + n_ranks - 1) / n_ranks;
  *s0 = std::min<uint64_t>(h->genome_size, (uint64_t)rank * per * tile);
  *s1 = std::min<uint64_t>(h->genome_size, (uint64_t)(rank + 1) * per * tile);
}

int reduce_slice(pemap_ctx* h, const pm::PeerPtrs& peers, uint64_t s0, uint64_t s1) {
  if (s1 > s0 && peers.n > 0) {
    // word range [6 * s0, 6 * s1): s0 is a multiple of 2048; round the end up to 16 bytes (the array has 64 bytes of slack)
    const uint64_t first = 6 * s0, n_words = (6 * (s1 - s0) + 3) & ~3ull;
    const int grid = h->sm_count * 8;
    switch (peers.n) {
      case 1: pm::k_reduce_slice<1><<<grid, 256, 0, h->stream>>>(h->d_counts, peers, first, n_words); break;
      case 3: pm::k_reduce_slice<3><<<grid, 256, 0, h->stream>>>(h->d_counts, peers, first, n_words); break;
      case 7: pm::k_reduce_slice<7><<<grid, 256, 0, h->stream>>>(h->d_counts, peers, first, n_words); break;
      default: pm::k_reduce_slice<0><<<grid, 256, 0, h->stream>>>(h->d_counts, peers, first, n_words); break;
    }
    h->stats.launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(h->stream));
  return PEMAP_OK;
}
}  // namespace

int pemap_counts_ipc_handle(pemap_t* h, void* handle64) {
  if (!h || !handle64) return PEMAP_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, h->d_counts));
  memcpy(handle64, &hd, 64);
  return PEMAP_OK;
}

int pemap_reduce_scatter_ipc(pemap_t* h, const void* handles, int n_ranks, int rank, uint64_t* site_first, uint64_t* site_end) {
  if (!h || !handles || n_ranks < 1 || n_ranks > 16 || rank < 0 || rank >= n_ranks) return PEMAP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  pm::PeerPtrs peers;
  peers.n = 0;
  for (int r = 0; r < n_ranks; r++) {
    if (r == rank) continue;
    const std::string key((const char*)handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    for (size_t k = 0; k < h->ipc_keys.size(); k++)
      if (h->ipc_keys[k] == key) ptr = h->ipc_open[k];
    if (!ptr) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, key.data(), 64);
      CK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
      h->ipc_keys.push_back(key);
      h->ipc_open.push_back(ptr);
    }
    peers.p[peers.n++] = (const uint32_t*)ptr;
  }
  uint64_t s0, s1;
  slice_of(h, n_ranks, rank, &s0, &s1);
  if (site_first) *site_first = s0;
  if (site_end) *site_end = s1;
  return reduce_slice(h, peers, s0, s1);
}

int pemap_reduce_scatter_local(pemap_t* const* hs, int n, int which, uint64_t* site_first, uint64_t* site_end) {
  if (!hs || n < 1 || n > 16 || which < 0 || which >= n || !hs[which]) return PEMAP_ERR_ARG;
  pemap_ctx* h = hs[which];
  CK(cudaSetDevice(h->device));
  pm::PeerPtrs peers;
  peers.n = 0;
  for (int r = 0; r < n; r++) {
    if (r == which) continue;
    if (!hs[r] || hs[r]->genome_size != h->genome_size) return fail(h, PEMAP_ERR_ARG, "handles index different genomes");
    if (hs[r]->device != h->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, h->device, hs[r]->device));
      if (!can) return fail(h, PEMAP_ERR_UNSUPPORTED, "no peer access between the two devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(hs[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(h, PEMAP_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
    peers.p[peers.n++] = hs[r]->d_counts;
  }
  uint64_t s0, s1;
  slice_of(h, n, which, &s0, &s1);
  if (site_first) *site_first = s0;
  if (site_end) *site_end = s1;
  return reduce_slice(h, peers, s0, s1);
}

int pemap_sw_score_device(pemap_t* h, int n, const char* d_reads, const int* d_len, int stride, int max_len,
                          const uint32_t* d_win_start, const int* d_win_len, int max_window, int32_t* d_score36, int32_t* d_maxi,
                          int32_t* d_maxk, int32_t* d_flags, float* ms) {
  if (!h || n < 0 || !d_reads || !d_len || !d_win_start || !d_win_len || !d_score36 || !d_maxi || !d_maxk)
    return fail(h, PEMAP_ERR_ARG, "NULL argument");
  if (h->index_only) return fail(h, PEMAP_ERR_ARG, "handle was opened with PEMAP_INDEX_ONLY=1");
  if (max_len < 16 || max_len > PM_DP_MAX - 22) return fail(h, PEMAP_ERR_ARG, "max_len out of range");
  if (max_window < 1 || max_window > PEMAP_SW_MAX_WINDOW) return fail(h, PEMAP_ERR_ARG, "window longer than PEMAP_SW_MAX_WINDOW");
  if ((uint32_t)n > h->task_cap) return fail(h, PEMAP_ERR_ARG, "more pairs than the task buffers of a chunk hold");
  CK(cudaSetDevice(h->device));
  const uint32_t nn = (uint32_t)n;
  CK(cudaMemsetAsync(h->d_cursors, 0, 128, h->stream));
  CK(cudaMemcpyAsync(h->d_cursors, &nn, 4, cudaMemcpyHostToDevice, h->stream));
  if (n) pm::k_sw_bench_tasks<<<(n + 255) / 256, 256, 0, h->stream>>>(n, d_win_start, d_win_len, h->d_tasks);
  pm::SwIntArgs ia;
  ia.tasks = h->d_tasks;
  ia.results = h->d_ires;
  ia.n_items = h->d_cursors;
  ia.work = h->d_cursors + 10;
  ia.reads[0] = d_reads;
  ia.reads[1] = d_reads;
  ia.len[0] = d_len;
  ia.len[1] = d_len;
  ia.stride = stride;
  ia.genome = h->d_genome;
  ia.lane_mm = -1;
  ia.p = h->dp;
  ia.list = nullptr;
  int uniform_len = 0;
  {
    unsigned mm[2] = {0xFFFFFFFFu, 0u};
    CK(cudaMemcpyAsync(h->d_cursors + 4, mm, 8, cudaMemcpyHostToDevice, h->stream));
    if (n > 0) k_len_range<<<h->sm_count * 4, 256, 0, h->stream>>>(d_len, nullptr, n, h->d_cursors + 4);
    CK(cudaMemcpyAsync(mm, h->d_cursors + 4, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n > 0 && mm[0] == mm[1]) uniform_len = (int)mm[0];
    if (n > 0 && (int)mm[1] > max_len) return fail(h, PEMAP_ERR_ARG, "a read is longer than max_len");
  }
  const int wd = max_len <= 112 ? 7 : max_len <= 160 ? 10 : max_len <= 256 ? 8 : 10;
  const int cmm = uniform_len > 0 ? (uniform_len - 1) % wd : -1;
  ia.lane_mm = uniform_len > 0 ? (uniform_len - 1) / wd : -1;
  cudaEvent_t e0 = h->slots[0].ev[0], e1 = h->slots[0].ev[1];
  CK(cudaEventRecord(e0, h->stream));
  if (max_len <= 112) LongDispatch<16, 7, 6>::go(h, ia, cmm);
  else if (max_len <= 160) LongDispatch<16, 10, 9>::go(h, ia, cmm);
  else if (max_len <= 256) LongDispatch<32, 8, 7>::go(h, ia, cmm);
  else LongDispatch<32, 10, 9>::go(h, ia, cmm);
  CK(cudaEventRecord(e1, h->stream));
  if (n) pm::k_sw_bench_results<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d_ires, d_score36, d_maxi, d_maxk, d_flags);
  h->stats.launches += 3;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
  return PEMAP_OK;
}

void* pemap_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void pemap_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int pemap_stream(pemap_t* h, void** stream) {
  if (!h || !stream) return PEMAP_ERR_ARG;
  *stream = (void*)h->stream;
  return PEMAP_OK;
}

int pemap_get_stats(pemap_t* h, pemap_stats* out) {
  if (!h || !out) return PEMAP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = fetch_counters(h);
  if (rc) return rc;
  *out = h->stats;
  return PEMAP_OK;
}

int pemap_reset_stats(pemap_t* h) {
  if (!h) return PEMAP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_counters, 0, sizeof(pm::SeedCounters), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memset(&h->stats, 0, sizeof(h->stats));
  return PEMAP_OK;
}

int pemap_index_device(pemap_t* h, const uint32_t** d_pos_index, const uint32_t** d_mers, uint64_t* n_mers) {
  if (!h) return PEMAP_ERR_ARG;
  if (d_pos_index) *d_pos_index = h->d_pos_index;
  if (d_mers) *d_mers = h->d_mers;
  if (n_mers) *n_mers = h->n_mers;
  return PEMAP_OK;
}

int pemap_read_pos_index(pemap_t* h, uint64_t first, uint64_t n, uint32_t* out) {
  if (!h || !out) return PEMAP_ERR_ARG;
  if (first + n > ((uint64_t)1 << 32) + 1) return fail(h, PEMAP_ERR_ARG, "range beyond 2^32+1");
  if (!h->d_pos_index) return fail(h, PEMAP_ERR_ARG, "pos_index is not resident (dropped after the bucket index was built; PEMAP_KEEP_INDEX=1)");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out, h->d_pos_index + first, n * 4, cudaMemcpyDeviceToHost));
  return PEMAP_OK;
}

int pemap_read_mers(pemap_t* h, uint64_t first, uint64_t n, uint32_t* out) {
  if (!h || !out) return PEMAP_ERR_ARG;
  if (first + n > h->n_mers) return fail(h, PEMAP_ERR_ARG, "range beyond n_mers");
  if (!h->d_mers) return fail(h, PEMAP_ERR_ARG, "mers is not resident (dropped after the bucket index was built; PEMAP_KEEP_INDEX=1)");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out, h->d_mers + first, n * 4, cudaMemcpyDeviceToHost));
  return PEMAP_OK;
}

void pemap_destroy(pemap_t* h) {
  if (!h) return;
  if (h->stream) {
    cudaSetDevice(h->device);
    if (h->lanes_) {
      use_lane(h, 0);
      sync_lanes(h);
      pemap_lane& o = h->lanes_->saved[1];  // the side lane's working set (lane 0's is the context's own fields)
      void* dv[] = {o.d_tasks, o.d_results, o.d_cursors, o.d_diag_winners, o.d_exact_winners, o.d_oob_winners, o.d_ires,
                    o.d_replay_reads, o.d_replay_tasks, o.d_sw_list, o.d_flagq, o.d_walk_meta, o.d_pair_codes, o.d_cand_base,
                    o.d_cand_n, o.d_winners, o.d_det_best, o.d_det_orient, o.d_det_score, o.d_seed_scratch, o.d_dirs, o.d_pend,
                    o.d_big_list, o.d_big_list2, o.d_big_scratch};
      for (void* p : dv)
        if (p) cudaFree(p);
      if (o.stream) cudaStreamDestroy(o.stream);
      if (o.s_aux) cudaStreamDestroy(o.s_aux);
      if (o.ev_fork) cudaEventDestroy(o.ev_fork);
      if (o.ev_join) cudaEventDestroy(o.ev_join);
      if (h->lanes_->ev_main) cudaEventDestroy(h->lanes_->ev_main);
      if (h->lanes_->ev_side) cudaEventDestroy(h->lanes_->ev_side);
      delete h->lanes_;
      h->lanes_ = nullptr;
    }
    cudaStreamSynchronize(h->stream);
#ifdef PM_TIE_DEBUG
    {
      unsigned long long why[32];
      if (cudaMemcpyFromSymbol(why, pm::g_tie_why, sizeof(why)) == cudaSuccess) {
        fprintf(stderr, "pemap tie debug:");
        for (int i = 0; i < 32; i++)
          if (why[i]) fprintf(stderr, " [%d]=%llu", i, i >= 16 ? why[i] / 32 : why[i]);
        fprintf(stderr, "\n");
      }
    }
#endif
    if (h->d_filter) cudaCtxResetPersistingL2Cache();  // give the persisting carve-out back
    void* dev[] = {h->d_filter, h->d_pos_index, h->d_mers, h->d_genome, h->d_cstart, h->d_border, h->d_counts, h->d_ins, h->d_ins_cursor,
                   h->d_tasks, h->d_results, h->d_cursors,
                   h->d_cand_base, h->d_cand_n, h->d_winners, h->d_det_best, h->d_det_orient,
                   h->d_det_score, h->d_seed_scratch, h->d_dirs, h->d_pend, h->d_counters, h->d_ires, h->d_replay_reads,
                   h->d_replay_tasks, h->d_diag_winners, h->d_exact_winners, h->d_oob_winners, h->d_flagq, h->d_walk_meta, h->d_pair_codes, h->d_sw_list};
    for (void* p : dev)
      if (p) cudaFree(p);
    for (void* p : h->ipc_open) cudaIpcCloseMemHandle(p);
    for (auto& sl : h->d_packed)
      for (auto& p : sl)
        if (p) cudaFree(p);
    void* rbi[] = {h->d_rbi_data[0], h->d_rbi_data[1], h->d_rbi_data[2], h->d_rbi_data[3], h->d_rbi_dir[0], h->d_rbi_dir[1],
                   h->d_rbi_dir[2], h->d_rbi_dir[3], h->d_big_list, h->d_big_list2, h->d_big_scratch};
    for (void* p : rbi)
      if (p) cudaFree(p);
    void* fin[] = {h->d_fin_rec[0], h->d_fin_rec[1], h->d_fin_cnt, h->d_fin_off, h->d_fin_tmp};
    for (void* p : fin)
      if (p) cudaFree(p);
    void* finh[] = {h->h_fin_rec[0], h->h_fin_rec[1], h->h_fin_total, h->h_ins_used};
    for (void* p : finh)
      if (p) cudaFreeHost(p);
    for (auto& ev : h->ev_fin)
      if (ev) cudaEventDestroy(ev);
    for (auto& sl : h->slots) {
      void* dv[] = {sl.d_reads[0], sl.d_reads[1], sl.d_len[0], sl.d_len[1], sl.d_m1, sl.d_m2, sl.d_type};
      for (void* p : dv)
        if (p) cudaFree(p);
      void* hv[] = {sl.h_reads[0], sl.h_reads[1], sl.h_len[0], sl.h_len[1], sl.h_m1, sl.h_m2, sl.h_type};
      for (void* p : hv)
        if (p) cudaFreeHost(p);
      for (auto& ev : sl.ev)
        if (ev) cudaEventDestroy(ev);
      if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
      if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
    }
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->s_aux) cudaStreamDestroy(h->s_aux);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

}  // extern "C"
