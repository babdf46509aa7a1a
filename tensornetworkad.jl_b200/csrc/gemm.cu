// FP64 tensor-core GEMM (DMMA m8n8k4) with multi-level operand strides.
//
// This is the single contraction kernel behind every `ein"..."` call site of the reference
// (trg.jl:25, ctmrg.jl:130-140, variationalipeps.jl:52-54 and their adjoints).  Each of the
// M, N, K and batch index groups of a pairwise contraction is described by up to MAXL
// (extent, stride) levels per operand, so the index permutations the reference performs with
// `permutedims` inside OMEinsum are folded into the tile loads ("permute-on-load") and the tile
// stores; no transpose kernel runs.
//
// Structure: CTA tile BM x BN x 16, 3-stage cp.async (LDGSTS) ring in shared memory, one warp tile
// WM x WN of m8n8k4 DMMA fragments per warp, accumulators in registers (sm_100 has no f64 kind for
// tcgen05/TMEM, so FP64 tensor math is warp-level mma.sync -> DMMA.8x8x4 in SASS).
// Operand tiles are kept in shared memory in the orientation they have in global memory
// (k-fast or row-fast) with a +4 double pad, which makes every fragment load bank-conflict free.
#include "common.h"

namespace tnad {

namespace {

constexpr int BK = 16;
constexpr int STAGES = 3;

__device__ __forceinline__ long long lvl_off(const LvlSet& L, int i) {
  long long o = 0;
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    if (l < L.nl) {
      if (l == L.nl - 1) {
        o += (long long)i * L.s[l];
      } else {
        int q = i / L.n[l];
        o += (long long)(i - q * L.n[l]) * L.s[l];
        i = q;
      }
    }
  }
  return o;
}

__device__ __forceinline__ void cp_async16(double* s, const double* g, bool pred) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(double* s, const double* g, bool pred) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Load one BR x BK operand tile (R = the M side of A or the N side of B) into shared memory.
//   KF == true : shared layout s[r * (BK+4) + k]   (global memory is contiguous along k)
//   KF == false: shared layout s[k * (BR+4) + r]   (global memory is contiguous along r)
// rowoff[r] holds the global offset of row r0+r, or -1 when the row is out of range.
template <int BR, int NT, bool KF>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, const long long* rowoff,
                                          const LvlSet& Lk, int k0, int K, bool vec, int tid) {
  if (KF) {
    constexpr int LD = BK + 4;
    if (vec) {
      constexpr int CPR = BK / 2;                 // 16-byte chunks per row
      constexpr int ITER = BR * CPR / NT;
      const int k2 = tid % CPR;
      const int k = k0 + 2 * k2;
      const bool kok = k < K;
      const long long ko = kok ? lvl_off(Lk, k) : 0;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int r = tid / CPR + i * (NT / CPR);
        const long long ro = rowoff[r];
        const bool ok = kok && ro >= 0;
        cp_async16(s + r * LD + 2 * k2, ok ? g + ro + ko : g, ok);
      }
    } else {
      constexpr int ITER = BR * BK / NT;
      const int kl = tid % BK;
      const int k = k0 + kl;
      const bool kok = k < K;
      const long long ko = kok ? lvl_off(Lk, k) : 0;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int r = tid / BK + i * (NT / BK);
        const long long ro = rowoff[r];
        const bool ok = kok && ro >= 0;
        cp_async8(s + r * LD + kl, ok ? g + ro + ko : g, ok);
      }
    }
  } else {
    constexpr int LD = BR + 4;
    if (vec) {
      constexpr int CPK = BR / 2;                 // 16-byte chunks per k column
      constexpr int ITER = BK * CPK / NT;
      static_assert(NT % CPK == 0, "tile/thread mismatch");
      const int r2 = tid % CPK;
      const long long ro = rowoff[2 * r2];
      const bool rok = ro >= 0;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int kl = tid / CPK + i * (NT / CPK);
        const int k = k0 + kl;
        const bool ok = rok && k < K;
        const long long ko = ok ? lvl_off(Lk, k) : 0;
        cp_async16(s + kl * LD + 2 * r2, ok ? g + ro + ko : g, ok);
      }
    } else {
      constexpr int ITER = BK * BR / NT;
      static_assert(NT % BR == 0, "tile/thread mismatch");
      const int r = tid % BR;
      const long long ro = rowoff[r];
      const bool rok = ro >= 0;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int kl = tid / BR + i * (NT / BR);
        const int k = k0 + kl;
        const bool ok = rok && k < K;
        const long long ko = ok ? lvl_off(Lk, k) : 0;
        cp_async8(s + kl * LD + r, ok ? g + ro + ko : g, ok);
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, bool AKF, bool BKF>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
    gemm_dmma_kernel(const __grid_constant__ GemmDesc d) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LDA = AKF ? (BK + 4) : (BM + 4);
  constexpr int LDB = BKF ? (BK + 4) : (BN + 4);
  constexpr int A_ELEMS = AKF ? BM * (BK + 4) : BK * (BM + 4);
  constexpr int B_ELEMS = BKF ? BN * (BK + 4) : BK * (BN + 4);
  constexpr int TM = WM / 8, TN = WN / 8;

  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_ELEMS;
  long long* rowA = reinterpret_cast<long long*>(Bs + STAGES * B_ELEMS);
  long long* rowB = rowA + BM;

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int bz = blockIdx.z;

  const double* gA = d.A + lvl_off(d.ab, bz);
  const double* gB = d.B + lvl_off(d.bb, bz);
  double* gC = d.C + lvl_off(d.cb, bz);

  for (int r = tid; r < BM; r += NT) rowA[r] = (m0 + r < d.M) ? lvl_off(d.am, m0 + r) : -1;
  for (int r = tid; r < BN; r += NT) rowB[r] = (n0 + r < d.N) ? lvl_off(d.bn, n0 + r) : -1;
  __syncthreads();

  const int K = d.K;
  const int nkt = (K + BK - 1) / BK;
  const bool avec = d.a_vec != 0, bvec = d.b_vec != 0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) {
      load_tile<BM, NT, AKF>(As + s * A_ELEMS, gA, rowA, d.ak, s * BK, K, avec, tid);
      load_tile<BN, NT, BKF>(Bs + s * B_ELEMS, gB, rowB, d.bk, s * BK, K, bvec, tid);
    }
    cp_async_commit();
  }

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % (BM / WM)) * WM;
  const int wn0 = (warp / (BM / WM)) * WN;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < nkt) {
        const int s = nk % STAGES;
        load_tile<BM, NT, AKF>(As + s * A_ELEMS, gA, rowA, d.ak, nk * BK, K, avec, tid);
        load_tile<BN, NT, BKF>(Bs + s * B_ELEMS, gB, rowB, d.bk, nk * BK, K, bvec, tid);
      }
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * A_ELEMS;
    const double* bs = Bs + (kt % STAGES) * B_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[TM], bf[TN];
      const int k = kk * 4 + t;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = wm0 + i * 8 + g;
        af[i] = AKF ? as[m * LDA + k] : as[k * LDA + m];
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = wn0 + j * 8 + g;
        bf[j] = BKF ? bs[n * LDB + k] : bs[k * LDB + n];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread holds C[m = wm0+8i+g][n = wn0+8j+2t+{0,1}]
  const double alpha = d.alpha, beta = d.beta;
  long long coff[TN][2];
  bool cok[TN][2];
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n0 + wn0 + j * 8 + 2 * t + e;
      cok[j][e] = n < d.N;
      coff[j][e] = cok[j][e] ? lvl_off(d.cn, n) : 0;
    }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + wm0 + i * 8 + g;
    if (m < d.M) {
      const long long ro = lvl_off(d.cm, m);
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (cok[j][e]) {
            double* pc = gC + ro + coff[j][e];
            double v = alpha * acc[i][j][e];
            if (beta != 0.0) v += beta * *pc;
            *pc = v;
          }
    }
  }
}

template <int BM, int BN, int WM, int WN, bool AKF, bool BKF>
void launch_cfg(tnad_ctx* c, const GemmDesc& d) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int A_ELEMS = AKF ? BM * (BK + 4) : BK * (BM + 4);
  constexpr int B_ELEMS = BKF ? BN * (BK + 4) : BK * (BN + 4);
  const size_t smem = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(double) + (BM + BN) * sizeof(long long);
  static bool attr_set = false;
  auto kern = gemm_dmma_kernel<BM, BN, WM, WN, AKF, BKF>;
  if (!attr_set) {
    TNAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((d.M + BM - 1) / BM, (d.N + BN - 1) / BN, d.batch);
  KTimer kt(c, KF_GEMM);
  kern<<<grid, NT, smem, c->stream>>>(d);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
}

template <int BM, int BN, int WM, int WN>
void launch_layout(tnad_ctx* c, const GemmDesc& d) {
  if (d.a_kfast) {
    if (d.b_kfast) launch_cfg<BM, BN, WM, WN, true, true>(c, d);
    else launch_cfg<BM, BN, WM, WN, true, false>(c, d);
  } else {
    if (d.b_kfast) launch_cfg<BM, BN, WM, WN, false, true>(c, d);
    else launch_cfg<BM, BN, WM, WN, false, false>(c, d);
  }
}

}  // namespace

void gemm_run(tnad_ctx* c, const GemmDesc& d) {
  if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return;
  TNAD_REQUIRE(d.batch <= 65535 && (d.N + 63) / 64 <= 65535, "gemm: grid too large");
  const long long big_tiles = (long long)((d.M + 127) / 128) * ((d.N + 127) / 128) * d.batch;
  if (big_tiles >= c->num_sms / 2 && d.M >= 96 && d.N >= 96)
    launch_layout<128, 128, 64, 32>(c, d);
  else
    launch_layout<64, 64, 32, 32>(c, d);
}

}  // namespace tnad
