// gather_probe.cu - how many DRAM bytes does one random 8-byte lookup into a 16 GiB table cost on this GPU?
// Times 2^27 random lookups (LCG addresses, 8 independent loads in flight per thread) for several load flavours and
// L2 fetch-granularity limits.  Event-timed; bytes/lookup = time x measured copy bandwidth is only an estimate, the
// exact figure comes from running this under ncu (dram__bytes_read.sum per kernel).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ uint2 load8(const uint32_t* p) {
  uint2 v;
  if (MODE == 0) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 6) asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 7) asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* tab, uint64_t mask, int iters, uint32_t* out) {
  uint64_t s = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    uint2 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      uint64_t idx = ((s >> 20) & mask) & ~1ull;  // even word -> 8-byte aligned pair
      v[u] = load8<MODE>(tab + idx);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) acc += v[u].y - v[u].x;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
float run(const uint32_t* tab, uint64_t mask, uint32_t* out, int blocks, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_gather<MODE><<<blocks, 256>>>(tab, mask, 2, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  k_gather<MODE><<<blocks, 256>>>(tab, mask, iters, out);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main(int argc, char** argv) {
  const uint64_t words = (1ull << 32);
  uint32_t* tab; uint32_t* out;
  CK(cudaMalloc(&tab, words * 4 + 64));
  CK(cudaMemset(tab, 1, words * 4 + 64));
  CK(cudaMalloc(&out, 64));
  const int blocks = 148 * 8, iters = 64;
  const double lookups = (double)blocks * 256 * iters * 8;
  const char* names[8] = {"ld.cg", "ld.nc", "ld.ca", "ld.cs", "ld.nc.L1::no_allocate", "ld.nc.no_alloc.L2::64B", "ld.volatile", "ld.relaxed.sys"};
  size_t lims[4] = {0, 32, 64, 128};
  printf("{\"lookups\": %.0f, \"runs\": [\n", lookups);
  for (int li = 0; li < 4; li++) {
    size_t got = 0;
    cudaError_t e = cudaSuccess;
    if (lims[li]) e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, lims[li]);
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    float ms[8];
    ms[0] = run<0>(tab, words - 1, out, blocks, iters);
    ms[1] = run<1>(tab, words - 1, out, blocks, iters);
    ms[2] = run<2>(tab, words - 1, out, blocks, iters);
    ms[3] = run<3>(tab, words - 1, out, blocks, iters);
    ms[4] = run<4>(tab, words - 1, out, blocks, iters);
    ms[5] = run<5>(tab, words - 1, out, blocks, iters);
    ms[6] = run<6>(tab, words - 1, out, blocks, iters);
    ms[7] = run<7>(tab, words - 1, out, blocks, iters);
    for (int m = 0; m < 8; m++)
      printf("  {\"limit_set\": %zu, \"set_rc\": \"%s\", \"limit_now\": %zu, \"load\": \"%s\", \"ms\": %.4f, \"Glookups_s\": %.2f}%s\n",
             lims[li], cudaGetErrorString(e), got, names[m], ms[m], lookups / ms[m] / 1e6, (li == 3 && m == 7) ? "" : ",");
  }
  printf("],\n \"range_sweep\": [\n");
  // footprint sweep: is the ~42 G lookups/s ceiling a DRAM (row activation) or a translation (TLB) limit, and what does
  // an L2-resident table deliver?
  for (int lg = 24; lg <= 32; lg += 1) {  // 2^lg words = 64 MB .. 16 GiB
    const uint64_t m = (1ull << lg) - 1;
    float a = run<0>(tab, m, out, blocks, iters), b = run<5>(tab, m, out, blocks, iters);
    a = run<0>(tab, m, out, blocks, iters);
    b = run<5>(tab, m, out, blocks, iters);
    printf("  {\"table_MB\": %llu, \"ld.cg_Glookups_s\": %.2f, \"L2::64B_Glookups_s\": %.2f}%s\n",
           (unsigned long long)((4ull << lg) >> 20), lookups / a / 1e6, lookups / b / 1e6, lg == 32 ? "" : ",");
  }
  printf("]}\n");
  return 0;
}
