// chunk_probe.cu - what does HBM deliver when the unit of random access is a contiguous chunk of S bytes instead
// of one 8-byte word?  (tools/gather_probe.cu measured 41.7 G random 8-byte lookups/s whatever the fetch size.)
// This decides the layout of the seed stage's device-private index: 49 isolated look-ups per segment against a few
// contiguous buckets that hold every k-mer sharing 24 of the 32 code bits.
//
// Sub-warps of LANES lanes each read a random 16-byte-aligned chunk of S bytes with uint4 loads, U chunks in flight
// per sub-warp.  A second kernel adds the dependent step of the real access pattern: an 8-byte read of a bucket
// directory (256 MB, four 64 MB tables) gives the chunk address.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tools/chunk_probe tools/chunk_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld16(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// S bytes per chunk, LANES lanes per chunk, U chunks in flight per sub-warp
template <int S, int LANES, int U>
__global__ void __launch_bounds__(256) k_chunks(const uint4* tab, uint64_t mask16, int iters, uint32_t* out) {
  constexpr int PER = S / (16 * LANES) > 0 ? S / (16 * LANES) : 1;   // uint4 loads per lane per chunk
  const int sub = (blockIdx.x * blockDim.x + threadIdx.x) / LANES, l = threadIdx.x % LANES;
  uint64_t s = (uint64_t)sub * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    uint4 v[U][PER];
#pragma unroll
    for (int u = 0; u < U; u++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t at = (s >> 24) & mask16;  // in 16-byte units
#pragma unroll
      for (int k = 0; k < PER; k++)
        if ((k * LANES + l) * 16 < S) v[u][k] = ld16(tab + at + k * LANES + l);
        else v[u][k] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int k = 0; k < PER; k++) acc += v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// dependent form: directory word pair (random, dir_words table) -> chunk at an address derived from it
template <int S, int LANES, int U>
__global__ void __launch_bounds__(256) k_dir_chunks(const uint4* tab, uint64_t mask16, const uint32_t* dir, uint32_t dir_mask,
                                                    int iters, uint32_t* out) {
  constexpr int PER = S / (16 * LANES) > 0 ? S / (16 * LANES) : 1;
  const int sub = (blockIdx.x * blockDim.x + threadIdx.x) / LANES, l = threadIdx.x % LANES;
  uint64_t s = (uint64_t)sub * 0x9E3779B97F4A7C15ull + 999;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    uint2 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const uint32_t w = (uint32_t)(s >> 28) & dir_mask & ~1u;
      d[u] = __ldg(reinterpret_cast<const uint2*>(dir + w));
    }
    uint4 v[U][PER];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t at = (((uint64_t)d[u].x << 7) ^ (uint64_t)d[u].y) & mask16;
#pragma unroll
      for (int k = 0; k < PER; k++)
        if ((k * LANES + l) * 16 < S) v[u][k] = ld16(tab + at + k * LANES + l);
        else v[u][k] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int k = 0; k < PER; k++) acc += v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

__global__ void k_fill(uint32_t* p, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t x = i * 0x9E3779B97F4A7C15ull;
    p[i] = (uint32_t)(x >> 29);
  }
}

template <class F>
float timed(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(2);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  f(32);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

template <int S, int LANES, int U>
void one(const uint4* tab, uint64_t mask16, const uint32_t* dir, uint32_t dir_mask, uint32_t* out, int ctas_per_sm, bool last) {
  const int blocks = 148 * ctas_per_sm;
  const double chunks = (double)blocks * (256 / LANES) * 32 * U;
  float ms = timed([&](int it) { k_chunks<S, LANES, U><<<blocks, 256>>>(tab, mask16, it, out); });
  float ms2 = timed([&](int it) { k_dir_chunks<S, LANES, U><<<blocks, 256>>>(tab, mask16, dir, dir_mask, it, out); });
  printf("  {\"chunk_bytes\": %d, \"lanes\": %d, \"in_flight\": %d, \"ctas_per_sm\": %d, \"Gchunks_s\": %.2f, \"GBs\": %.0f, "
         "\"dir_Gchunks_s\": %.2f, \"dir_GBs\": %.0f}%s\n",
         S, LANES, U, ctas_per_sm, chunks / ms / 1e6, chunks * S / ms / 1e6, chunks / ms2 / 1e6, chunks * S / ms2 / 1e6, last ? "" : ",");
  fflush(stdout);
}

int main(int argc, char** argv) {
  // argv[1] = GB of table the chunk starts are spread over (default 32; the rotated bucket arrays of a 3.1 Gb genome are 63 GB)
  const uint64_t range_gb = argc > 1 ? (uint64_t)atoi(argv[1]) : 32;
  const uint64_t bytes = (range_gb + 1) << 30;
  uint4* tab; uint32_t* out; uint32_t* dir;
  CK(cudaMalloc(&tab, bytes + 65536));
  CK(cudaMemset(tab, 1, bytes + 65536));
  CK(cudaMalloc(&out, 64));
  const uint32_t dir_words = 1u << 26;  // 256 MB directory
  CK(cudaMalloc(&dir, (size_t)dir_words * 4 + 64));
  k_fill<<<148 * 8, 256>>>(dir, dir_words);
  CK(cudaDeviceSynchronize());
  // chunk starts: any 16-byte unit inside the first 32 GiB (mask), chunks may run 64 KB past it
  const uint64_t mask16 = ((range_gb << 30) / 16) - 1;  // (a power of two is expected)
  printf("{\"table_GB\": %llu, \"dir_MB\": 256, \"runs\": [\n", (unsigned long long)range_gb);
  one<64, 4, 8>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<128, 8, 4>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<128, 8, 8>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<256, 8, 4>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<256, 8, 8>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<512, 8, 2>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<512, 8, 4>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 8, 1>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 8, 2>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 8, 4>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 32, 2>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 32, 4>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<1024, 8, 2>(tab, mask16, dir, dir_words - 1, out, 4, false);
  one<2048, 8, 1>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<2048, 8, 2>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<2048, 32, 2>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<4096, 32, 1>(tab, mask16, dir, dir_words - 1, out, 8, false);
  one<4096, 32, 2>(tab, mask16, dir, dir_words - 1, out, 8, true);
  printf("]}\n");
  return 0;
}
