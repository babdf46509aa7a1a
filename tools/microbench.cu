// Micro-benchmarks that size the register-resident Jacobi kernel: vector FP64 FMA rate, 64-bit warp-shuffle
// rate, __syncthreads cost, double rsqrt latency.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(int iters, double* out) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_shfl(int iters, double* out) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_sync(int iters, double* out) {
  __shared__ double sh[256];
  double a = threadIdx.x;
  for (int it = 0; it < iters; ++it) { sh[threadIdx.x] = a; __syncthreads(); a += sh[(threadIdx.x + 32) & 255]; }
  if (a == 1.2345) out[0] = a;
}
__global__ void k_rsqrt(int iters, double* out) {
  double a = 1.0 + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) a = rsqrt(a) + 1.0;
  if (a == 1.2345) out[0] = a;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  double* d; cudaMalloc(&d, 64);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e3;
  printf("%s sms=%d clock=%.0f MHz\n", p.name, sms, clk / 1e6);
  for (int warps : {4, 8, 16, 32}) {
    const int it = 20000;
    float ms = timeit([&] { k_dfma<<<sms, warps * 32>>>(it, d); });
    double fma = (double)sms * warps * 32 * it * 8;
    printf("dfma  warps/SM=%2d: %.2f TFLOP/s, %.1f lane-FMA/clk/SM (nominal clock)\n", warps, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / clk / sms);
    ms = timeit([&] { k_shfl<<<sms, warps * 32>>>(it, d); });
    double sh = (double)sms * warps * it * 8;
    printf("shfl64 warps/SM=%2d: %.2f warp-shfl64/clk/SM\n", warps, sh / (ms * 1e-3) / clk / sms);
  }
  float ms = timeit([&] { k_sync<<<sms, 256>>>(100000, d); });
  printf("syncthreads+smem roundtrip (256 thr): %.0f clk/iter\n", ms * 1e-3 * clk / 100000);
  ms = timeit([&] { k_rsqrt<<<sms, 32>>>(100000, d); });
  printf("double rsqrt+add dependent chain: %.0f clk/iter\n", ms * 1e-3 * clk / 100000);
  return 0;
}
