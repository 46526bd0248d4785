// FP64-pipe microbenchmark for the roofline denominator of the MHD kernels:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
// Measures sustained DFMA throughput (thread-instructions/s) versus warps per SM and per-thread ILP, the
// dependent-issue latency of DFMA, and the throughput of the MUFU.RCP64H + cubic-Newton reciprocal.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA mixed 1:1 with integer IMAD work (does the scheduler co-issue around the half-rate FP64 pipe?)
template <int ILP>
__global__ void k_mix(double *out, int iters, double a, double b, int m) {
  double x[ILP];
  int y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) { x[i] = fma(x[i], a, b); y[i] = y[i] * m + 12345; }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ double frcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}
template <int ILP, bool IEEE>
__global__ void k_rcp(double *out, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = 1.5 + threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = IEEE ? 1.0 / x[i] : frcp(x[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk / 1000);
  double *out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  const int iters = 2000;
  printf("# DFMA: warps/SM, ILP -> T thread-DFMA/s (x2 = FLOP/s), DFMA per clk per SM at max clock\n");
  for (int wps : {4, 8, 16, 32, 64}) {
    const int threads = 128, blocks = 148 * wps / 4;
#define RUN(ILP) { float ms = time_ms([&] { k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }); \
      double n = (double)blocks * threads * iters * 16.0 * ILP; \
      printf("  warps/SM %2d ILP %d : %7.2f T/s  (%5.1f /clk/SM)\n", wps, ILP, n / ms * 1e-9, n / ms * 1e-3 / 148 / (clk * 1e3) * 1e0); }
    RUN(1) RUN(2) RUN(4) RUN(8)
  }
  printf("# dependent DFMA latency: 1 warp per SM sub-partition, ILP 1\n");
  { float ms = time_ms([&] { k_dfma<1><<<148, 128>>>(out, iters, 1.0000001, 1e-9); });
    printf("  %.1f cycles per dependent DFMA (at max clock)\n", ms * 1e-3 * clk * 1e3 / (iters * 16.0)); }
  printf("# DFMA + IMAD 1:1 (16 warps/SM)\n");
  { const int threads = 128, blocks = 148 * 4;
    float ms = time_ms([&] { k_mix<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 3); });
    double n = (double)blocks * threads * iters * 16.0 * 4;
    printf("  ILP 4: %7.2f T DFMA/s alongside the same number of IMADs\n", n / ms * 1e-9);
    ms = time_ms([&] { k_mix<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 3); });
    n = (double)blocks * threads * iters * 16.0 * 8;
    printf("  ILP 8: %7.2f T DFMA/s alongside the same number of IMADs\n", n / ms * 1e-9); }
  printf("# reciprocal throughput, 16 warps/SM, ILP 4: T rcp/s\n");
  { const int threads = 128, blocks = 148 * 4;
    float ms = time_ms([&] { k_rcp<4, true><<<blocks, threads>>>(out, iters); });
    double n = (double)blocks * threads * iters * 4.0 * 4;
    printf("  IEEE 1.0/x          : %6.3f T/s\n", n / ms * 1e-9);
    ms = time_ms([&] { k_rcp<4, false><<<blocks, threads>>>(out, iters); });
    printf("  MUFU + cubic Newton : %6.3f T/s\n", n / ms * 1e-9); }
  return 0;
}
