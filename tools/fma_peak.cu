// FP32 FMA peak of the GPU this runs on (SURVEY 8d "binding roofline": the cube kernel is bound by the FP32 /
// issue pipes, not by HBM, so its honest denominator is a MEASURED FMA rate, next to MEASURED_PEAKS.json).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fma_peak tools/fma_peak.cu
//   tools/fma_peak > profiles/fp32_peak.json          (on the B200 box; ~1 s)
//
// Three dependent-chain-free loops, every SM filled (blocks = SMs x resident blocks), timed with CUDA events after
// a warm-up, best of `reps` (bench.py samples the SM clock under load through NVML):
//   ffma    16 independent scalar FFMA chains per thread          -> 2 flop per lane per instruction
//   ffma2   16 independent packed fma.rn.f32x2 chains per thread  -> 4 flop per lane per instruction
//   issue   the ffma loop counted in warp instructions: the issue-slot ceiling of 4 schedulers per SM
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                         \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

constexpr int CHAINS = 16;
constexpr int INNER = 64;   // unrolled FMAs per chain and loop trip

__global__ void __launch_bounds__(256) ffma_kernel(float *out, float a, float b, int trips, long long *cycles) {
  float acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (float)(threadIdx.x + c);
  const long long t0 = clock64();
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < INNER; ++i)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) acc[c] = fmaf(acc[c], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float *out, float a, float b, int trips, long long *cycles) {
  unsigned long long acc[CHAINS];
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    const float v = (float)(threadIdx.x + c);
    asm("mov.b64 %0, {%1, %1};" : "=l"(acc[c]) : "f"(v));
  }
  const long long t0 = clock64();
  for (int t = 0; t < trips; ++t) {
#pragma unroll
    for (int i = 0; i < INNER; ++i)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[c]) : "l"(av), "l"(bv));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[c]));
    s += lo + hi;
  }
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main(int argc, char **argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  const int threads = 256, per_sm = 8;   // 64 warps per SM: every scheduler always has a ready warp
  const int blocks = sms * per_sm;
  const int trips = argc > 1 ? atoi(argv[1]) : 400;
  const int reps = 10;
  float *out;
  long long *cyc;
  CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<long long> h(blocks);
  struct Res { double ms, mhz; } res[2];
  for (int which = 0; which < 2; ++which) {
    double best = 1e30, best_mhz = 0;
    for (int r = 0; r < reps + 3; ++r) {
      CK(cudaEventRecord(e0));
      if (which == 0) ffma_kernel<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, trips, cyc);
      else ffma2_kernel<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, trips, cyc);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r < 3) continue;   // warm-up
      CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
      // blocks run in one wave (8 resident x 148): a block's cycle count over the launch time is the SM clock
      std::sort(h.begin(), h.end());
      const double mhz = (double)h[blocks / 2] / (ms * 1e-3) / 1e6;
      if (ms < best) { best = ms; best_mhz = mhz; }
    }
    res[which].ms = best;
    res[which].mhz = best_mhz;
  }
  const double lanes = (double)blocks * threads;
  const double fma_per_lane = (double)trips * INNER * CHAINS;
  const double ffma_tflops = 2.0 * lanes * fma_per_lane / (res[0].ms * 1e-3) / 1e12;
  const double ffma2_tflops = 4.0 * lanes * fma_per_lane / (res[1].ms * 1e-3) / 1e12;
  const double warp_inst_per_s = lanes / 32.0 * fma_per_lane / (res[0].ms * 1e-3);
  const double warp_inst2_per_s = lanes / 32.0 * fma_per_lane / (res[1].ms * 1e-3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"blocks\": %d, \"threads\": %d, \"chains\": %d,\n"
         " \"ffma_tflops\": %.2f, \"ffma_ms\": %.4f,\n"
         " \"ffma2_tflops\": %.2f, \"ffma2_ms\": %.4f,\n"
         " \"fp32_tflops\": %.2f,\n"
         " \"ffma_warp_inst_per_s\": %.4e, \"ffma2_warp_inst_per_s\": %.4e,\n"
         " \"nominal_tflops_at_max_clock\": %.2f, \"max_clock_mhz\": %.0f,\n"
         " \"ffma_warp_inst_per_clk_per_sm_at_max_clock\": %.3f,\n"
         " \"how\": \"tools/fma_peak.cu: %d independent FMA chains per thread, %d x %d threads, best of %d after 3 warm-ups, "
         "CUDA events\"}\n",
         prop.name, sms, blocks, threads, CHAINS, ffma_tflops, res[0].ms, ffma2_tflops, res[1].ms,
         std::max(ffma_tflops, ffma2_tflops), warp_inst_per_s, warp_inst2_per_s,
         2.0 * sms * 128 * prop.clockRate * 1e3 / 1e12, prop.clockRate / 1e3,
         warp_inst_per_s / (sms * prop.clockRate * 1e3), CHAINS, blocks, threads, reps);
  return 0;
}
