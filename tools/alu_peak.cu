// alu_peak.cu - measured issue rates of the instructions the Smith-Waterman kernels are made of (SURVEY.md 8d:
// "L must be measured on the box with a dependency-free issue-rate microbenchmark").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_peak tools/alu_peak.cu ; prints one JSON object.
// Every kernel runs 8 independent dependency chains per thread, 1024 threads per SM-resident block set, long enough
// (>= 50 ms per launch) for the CUDA-event time to be the kernel and not its launch, and reports ops per clock per SM
// twice: from the clock64() deltas inside the kernel and from the event time x the SM clock nvidia-smi shows DURING
// the launches (sampled every 100 ms, median printed).  The two must agree; round 1's 0.1 ms kernels did not
// (launch overhead in the event time).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHAINS 8
#define ITERS (1 << 20)

template <int OP>
__device__ __forceinline__ unsigned int_op(unsigned a, unsigned b, unsigned c) {
  if (OP == 0) return a + b;                                   // IADD3
  if (OP == 1) return (unsigned)max((int)a, (int)b);           // VIMNMX / IMNMX
  if (OP == 2) return __viaddmax_s16x2(a, b, c);               // VIADDMNMX.S16x2
  if (OP == 3) return __vimax3_s16x2(a, b, c);                 // VIMNMX3.S16x2
  if (OP == 4) return __vmaxs2(a, b);                          // VIMNMX.S16x2
  if (OP == 5) return __vadd2(a, b);                           // VIADD.16x2
  if (OP == 6) return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);  // VIADDMNMX
  if (OP == 7) return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);    // VIMNMX3
  return a;
}

template <int OP>
__global__ void __launch_bounds__(256) k_int(unsigned* out, unsigned seed, long long* cycles) {
  unsigned v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) v[i] = seed * (threadIdx.x + 1) + i;
  const unsigned b = seed | 1u, c = seed ^ 0x00030003u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    const unsigned bi = b + (unsigned)it * 0x00010001u;  // varies per iteration so max() chains are not idempotent
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = int_op<OP>(v[i], bi, c);
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(256) k_f64(double* out, double seed, long long* cycles) {
  double v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) v[i] = seed * (threadIdx.x + 1) + i;
  const double b = seed * 1e-9, c = seed * 3.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (OP == 0) v[i] = __dadd_rn(v[i], b);                  // DADD
      else v[i] = (v[i] > c) ? __dadd_rn(v[i], -b) : __dadd_rn(v[i], b) ;  // DSETP + select + DADD
    }
  }
  long long t1 = clock64();
  double acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int sms, int blocks_per_sm, long long* d_cycles, bool last) {
  const int blocks = sms * blocks_per_sm;
  launch(blocks);  // warm-up
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(blocks);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  cudaMemcpy(h, d_cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; i++) avg += (double)h[i];
  avg /= blocks;
  const double ops_per_block = 256.0 * CHAINS * ITERS;
  const double per_clk_sm = ops_per_block * blocks_per_sm / avg;  // resident blocks share the SM for `avg` cycles
  const double gops = ops_per_block * blocks / (ms * 1e6);
  printf("  \"%s\": {\"ops_per_clk_per_sm\": %.2f, \"gops\": %.1f, \"ms\": %.4f, \"ops_per_clk_per_sm_from_events_at_1965MHz\": %.2f}%s\n",
         name, per_clk_sm, gops, ms, gops * 1e9 / (1965e6 * sms), last ? "" : ",");
  fflush(stdout);
  free(h);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, bps = 4;
  unsigned* d_u;
  double* d_d;
  long long* d_c;
  cudaMalloc(&d_u, sizeof(unsigned) * 256 * sms * bps);
  cudaMalloc(&d_d, sizeof(double) * 256 * sms * bps);
  cudaMalloc(&d_c, sizeof(long long) * sms * bps);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  FILE* smi = popen("nvidia-smi -i 0 --query-gpu=clocks.sm --format=csv,noheader,nounits -lms 100 -c 60 2>/dev/null", "r");
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_rate_khz\": %d, \"chains\": %d, \"threads_per_sm\": %d,\n", p.name, sms, clk_khz,
         CHAINS, 256 * bps);
#define RUN_INT(OP, NAME) run(NAME, [&](int b) { k_int<OP><<<b, 256>>>(d_u, 12345u, d_c); }, sms, bps, d_c, false)
  RUN_INT(0, "IADD");
  RUN_INT(1, "IMNMX_s32");
  RUN_INT(2, "VIADDMNMX_s16x2");
  RUN_INT(3, "VIMNMX3_s16x2");
  RUN_INT(4, "VIMNMX_s16x2");
  RUN_INT(5, "VIADD_16x2");
  RUN_INT(6, "VIADDMNMX_s32");
  RUN_INT(7, "VIMNMX3_s32");
  run("DADD", [&](int b) { k_f64<0><<<b, 256>>>(d_d, 1.5, d_c); }, sms, bps, d_c, false);
  run("DSETP_SEL_DADD", [&](int b) { k_f64<1><<<b, 256>>>(d_d, 1.5, d_c); }, sms, bps, d_c, false);
  std::vector<int> mhz;
  if (smi) {  // the sampler (60 samples, 6 s) has been printing since before the first launch
    char line[64];
    while (fgets(line, sizeof line, smi)) mhz.push_back(atoi(line));
    pclose(smi);
  }
  std::sort(mhz.begin(), mhz.end());
  printf("  \"sm_clock_mhz_during_run\": {\"samples\": %zu, \"min\": %d, \"median\": %d, \"max\": %d}\n}\n", mhz.size(),
         mhz.empty() ? 0 : mhz.front(), mhz.empty() ? 0 : mhz[mhz.size() / 2], mhz.empty() ? 0 : mhz.back());
  return 0;
}
