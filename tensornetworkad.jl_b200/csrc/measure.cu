// Measurement helpers behind bench.py: pinned host buffers, an event stopwatch on the context's own
// stream, the per-kernel-family device-time table, and the FP64 DMMA issue-rate microbenchmark that
// supplies the `tensor` roofline denominator (MEASURED_PEAKS.json carries no FP64 number).
#include "drivers.h"

using namespace tnad;

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Every warp issues `iters` x 16 independent DMMAs on register operands (no memory traffic).
__global__ void __launch_bounds__(256) k_dmma_peak(int iters, double* sink) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;   // keep the loop alive
}

}  // namespace

#define API_BEGIN(ctx)            \
  if (!(ctx)) return TNAD_ERR_ARG; \
  try {                           \
    TNAD_CUDA(cudaSetDevice((ctx)->device));
#define API_END(ctx)                                         \
    return TNAD_OK;                                          \
  } catch (const tnad::Error& e) {                           \
    (ctx)->err = e.msg;                                      \
    cudaGetLastError();                                      \
    return e.code;                                           \
  } catch (...) {                                            \
    (ctx)->err = "unknown internal error";                   \
    return TNAD_ERR_INTERNAL;                                \
  }

extern "C" {

int tnad_host_alloc(tnad_ctx* c, int64_t n, double** hptr) {
  API_BEGIN(c)
  TNAD_REQUIRE(hptr && n >= 0, "tnad_host_alloc: bad arguments");
  TNAD_CUDA(cudaMallocHost((void**)hptr, (size_t)(n ? n : 1) * sizeof(double)));
  API_END(c)
}

int tnad_host_free(tnad_ctx* c, double* hptr) {
  API_BEGIN(c)
  TNAD_CUDA(cudaFreeHost(hptr));
  API_END(c)
}

int tnad_timer_start(tnad_ctx* c) {
  API_BEGIN(c)
  if (!c->tstart) {
    TNAD_CUDA(cudaEventCreate(&c->tstart));
    TNAD_CUDA(cudaEventCreate(&c->tstop));
  }
  TNAD_CUDA(cudaStreamSynchronize(c->stream));
  TNAD_CUDA(cudaEventRecord(c->tstart, c->stream));
  API_END(c)
}

int tnad_timer_stop(tnad_ctx* c, double* ms) {
  API_BEGIN(c)
  TNAD_REQUIRE(ms && c->tstart, "tnad_timer_stop: timer was not started");
  TNAD_CUDA(cudaEventRecord(c->tstop, c->stream));
  TNAD_CUDA(cudaEventSynchronize(c->tstop));
  float f = 0.f;
  TNAD_CUDA(cudaEventElapsedTime(&f, c->tstart, c->tstop));
  *ms = f;
  API_END(c)
}

int tnad_set_kernel_timing(tnad_ctx* c, int enable) {
  API_BEGIN(c)
  sync(c);
  for (auto& k : c->kspans) {
    c->event_pool.push_back(k.a);
    c->event_pool.push_back(k.b);
  }
  c->kspans.clear();
  c->ktiming = enable != 0;
  c->gemm_flops = c->gemm_tma_flops = 0.0;
  c->gemm_tma_n = c->gemm_fallback_n = 0;
  TNAD_CUDA(cudaMemsetAsync(c->scal + 20, 0, 2 * sizeof(double), c->stream));
  sync(c);
  API_END(c)
}

int tnad_kernel_timing(tnad_ctx* c, double* ms, int64_t* count) {
  API_BEGIN(c)
  TNAD_REQUIRE(ms && count, "tnad_kernel_timing: null output");
  sync(c);
  for (int i = 0; i < KF_NFAM; ++i) {
    ms[i] = 0.0;
    count[i] = 0;
  }
  for (auto& k : c->kspans) {
    float f = 0.f;
    if (cudaEventElapsedTime(&f, k.a, k.b) == cudaSuccess && k.fam >= 0 && k.fam < KF_NFAM) {
      ms[k.fam] += f;
      count[k.fam] += 1;
    }
  }
  // executed work units of the DMMA update kernels: [5] 64x64 blocks of k_sym_update_m, [6] 128x64 slabs of k_jacobi_update
  unsigned long long wc[2] = {0, 0};
  TNAD_CUDA(cudaMemcpyAsync(wc, c->scal + 20, sizeof(wc), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  count[5] = (int64_t)wc[0];
  count[6] = (int64_t)wc[1];
  // GEMM family: algorithmic flops and the split between the TMA kernel and the cp.async kernel (unused family slots)
  ms[13] = c->gemm_flops;
  ms[14] = c->gemm_tma_flops;
  count[13] = c->gemm_tma_n;
  count[14] = c->gemm_fallback_n;
  API_END(c)
}

int tnad_dmma_peak(tnad_ctx* c, double* tflops) {
  API_BEGIN(c)
  TNAD_REQUIRE(tflops, "tnad_dmma_peak: null output");
  Tens sink = t_alloc(c, {4}, true);
  const int iters = 4096, blocks = c->num_sms * 4, warps = 8;
  for (int wu = 0; wu < 4; ++wu) k_dmma_peak<<<blocks, 256, 0, c->stream>>>(iters, sink.p);   // warm-up: 2 ms of full-rate DMMA (the probe was bimodal, 29.7 / 37.0 TFLOP/s, with a 64-iteration warm-up)
  TNAD_CUDA(cudaGetLastError());
  cudaEvent_t a = get_event(c), b = get_event(c);
  double best = 0.0;
  for (int rep = 0; rep < 10; ++rep) {
    TNAD_CUDA(cudaEventRecord(a, c->stream));
    k_dmma_peak<<<blocks, 256, 0, c->stream>>>(iters, sink.p);
    TNAD_CUDA(cudaEventRecord(b, c->stream));
    TNAD_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    TNAD_CUDA(cudaEventElapsedTime(&ms, a, b));
    const double flops = (double)blocks * warps * iters * 16.0 * 512.0;   // m8n8k4 = 256 FMA
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  c->event_pool.push_back(a);
  c->event_pool.push_back(b);
  c->launches += 14;
  *tflops = best;
  API_END(c)
}

int tnad_gemm_bench(tnad_ctx* c, int m, int n, int k, int reps, double* tflops) {
  API_BEGIN(c)
  TNAD_REQUIRE(tflops && m > 0 && n > 0 && k > 0 && reps > 0, "tnad_gemm_bench: bad arguments");
  Tens A = t_alloc(c, {m, k}), B = t_alloc(c, {k, n}), C = t_alloc(c, {m, n});
  fill(c, A.p, A.numel(), 1.0 / 3.0);
  fill(c, B.p, B.numel(), 0.5);
  contract(c, "ik,kj->ij", A, B, C);   // warm-up
  cudaEvent_t a = get_event(c), b = get_event(c);
  TNAD_CUDA(cudaEventRecord(a, c->stream));
  for (int r = 0; r < reps; ++r) contract(c, "ik,kj->ij", A, B, C);
  TNAD_CUDA(cudaEventRecord(b, c->stream));
  TNAD_CUDA(cudaEventSynchronize(b));
  float ms = 0.f;
  TNAD_CUDA(cudaEventElapsedTime(&ms, a, b));
  c->event_pool.push_back(a);
  c->event_pool.push_back(b);
  *tflops = 2.0 * m * (double)n * k * reps / (ms * 1e-3) / 1e12;
  API_END(c)
}

}  // extern "C"
