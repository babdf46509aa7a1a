// DMMA (mma.sync m8n8k4 f64) issue rate against the number of warps per SM and the number of independent accumulator
// tiles per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(int iters, double* out) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
// the GEMM inner step: 4 A fragments x 4 B fragments (distinct registers, re-loaded from shared memory every step) -> 16 DMMAs
template <bool LDS>
__global__ void k44(int iters, double* out) {
  __shared__ double sh[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 1e-3 * i;
  __syncthreads();
  double c[4][4][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  double a[4], b[4];
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < 4; ++i) { a[i] = sh[lane + 32 * i]; b[i] = sh[256 + lane + 32 * i]; }
  for (int it = 0; it < iters; ++it) {
    double an[4], bn[4];
    if (LDS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { an[i] = sh[((it * 64 + lane + 32 * i) & 1023)]; bn[i] = sh[1024 + ((it * 64 + lane + 32 * i) & 1023)]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    if (LDS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = an[i]; b[i] = bn[i]; }
    }
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
  if (s == 1.2345) out[0] = s;
}
template <bool LDS>
void run44(int sms, double clk, double* d) {
  for (int warps : {4, 8, 12, 16}) {
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k44<LDS><<<sms, warps * 32>>>(64, d); cudaDeviceSynchronize();
    cudaEventRecord(e0); k44<LDS><<<sms, warps * 32>>>(iters, d); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)sms * warps * iters * 16;
    printf("4x4 fragments %s  warps/SM %2d: %.2f TFLOP/s, %.1f cycles per DMMA per scheduler\n", LDS ? "(LDS per step)" : "(registers)   ", warps,
           n * 512 / ms / 1e9, ms * 1e-3 * clk / (n / sms / 4));
  }
}
template <int NACC>
void run(int sms, double clk, double* d) {
  for (int warps : {4, 8, 12, 16}) {
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC><<<sms, warps * 32>>>(64, d); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<NACC><<<sms, warps * 32>>>(iters, d); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)sms * warps * iters * NACC;
    printf("acc tiles/warp %2d  warps/SM %2d: %.2f TFLOP/s, %.1f cycles per DMMA per scheduler (nominal clock)\n", NACC, warps,
           n * 512 / ms / 1e9, ms * 1e-3 * clk / (n / sms / 4));
  }
}
int main() {
  double* d; cudaMalloc(&d, 64);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double clk = p.clockRate * 1e3;
  printf("%s sms=%d clock=%.0f MHz\n", p.name, p.multiProcessorCount, clk / 1e6);
  run<16>(p.multiProcessorCount, clk, d);
  run44<false>(p.multiProcessorCount, clk, d);
  run44<true>(p.multiProcessorCount, clk, d);
  return 0;
}
