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
#include <algorithm>

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

// Out-of-line copy for set-up / epilogue code: keeps the integer divisions out of the unrolled main loop
// (the first version inlined them into every tile load and stalled on instruction fetch -- ncu: 35-60 % of
// issue slots lost to `no_instruction`).
__device__ __noinline__ long long lvl_off_ni(const LvlSet& L, int i) { return lvl_off(L, i); }

// Load one BR x BK operand tile (R = the M side of A or the N side of B) into shared memory.
//   KF == true : shared layout s[r * (BK+4) + k]   (global memory is contiguous along k)
//   KF == false: shared layout s[k * (BR+4) + r]   (global memory is contiguous along r)
// rowoff[r]: global offset of row r0+r (-1: out of range); kof[kl]: offset of k index kl of this tile (-1: past K).
template <int BR, int NT, bool KF>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, const long long* rowoff,
                                          const long long* kof, bool vec, int tid) {
  if (KF) {
    constexpr int LD = BK + 4;
    if (vec) {
      constexpr int CPR = BK / 2;                 // 16-byte chunks per row
      constexpr int ITER = BR * CPR / NT;
      const int k2 = tid % CPR;
      const long long ko = kof[2 * k2];
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int r = tid / CPR + i * (NT / CPR);
        const long long ro = rowoff[r];
        const bool ok = ko >= 0 && ro >= 0;
        cp_async16(s + r * LD + 2 * k2, ok ? g + ro + ko : g, ok);
      }
    } else {
      constexpr int ITER = BR * BK / NT;
      const int kl = tid % BK;
      const long long ko = kof[kl];
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int r = tid / BK + i * (NT / BK);
        const long long ro = rowoff[r];
        const bool ok = ko >= 0 && ro >= 0;
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
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int kl = tid / CPK + i * (NT / CPK);
        const long long ko = kof[kl];
        const bool ok = ro >= 0 && ko >= 0;
        cp_async16(s + kl * LD + 2 * r2, ok ? g + ro + ko : g, ok);
      }
    } else {
      constexpr int ITER = BK * BR / NT;
      static_assert(NT % BR == 0, "tile/thread mismatch");
      const int r = tid % BR;
      const long long ro = rowoff[r];
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int kl = tid / BR + i * (NT / BR);
        const long long ko = kof[kl];
        const bool ok = ro >= 0 && ko >= 0;
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
  long long* kofA = rowB + BN;                 // [STAGES][BK]
  long long* kofB = kofA + STAGES * BK;        // [STAGES][BK]

  const int tid = threadIdx.x;
  // The level sets are read through out-of-line calls (code size); a reference into the kernel parameters turns every
  // access into a generic load with hundreds of cycles of latency (ncu: long-scoreboard stalls, about 0.5 us per call,
  // 12 calls per thread in the epilogue -- most of the 13 us a tiny product took).  A copy in shared memory costs 30 cycles.
  __shared__ LvlSet slv[9];                      // am, ak, bk, bn, cm, cn, ab, bb, cb (their order in GemmDesc)
  if (tid < 9) slv[tid] = (&d.am)[tid];
  __syncthreads();
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int S = d.splitk > 1 ? d.splitk : 1;
  const int bz = blockIdx.z / S, sp = blockIdx.z - bz * S;

  const double* gA = d.A + lvl_off_ni(slv[6], bz);
  const double* gB = d.B + lvl_off_ni(slv[7], bz);

  // this CTA's K range (whole k-tiles)
  const int K = d.K;
  const int nkt_all = (K + BK - 1) / BK;
  const int per = (nkt_all + S - 1) / S;
  const int kt0 = sp * per;
  const int nkt = max(0, min(per, nkt_all - kt0));
  const int kbase = kt0 * BK;

  for (int r = tid; r < BM; r += NT) rowA[r] = (m0 + r < d.M) ? lvl_off_ni(slv[0], m0 + r) : -1;
  for (int r = tid; r < BN; r += NT) rowB[r] = (n0 + r < d.N) ? lvl_off_ni(slv[3], n0 + r) : -1;
  // k-offset tables of the first STAGES tiles
  for (int q = tid; q < 2 * STAGES * BK; q += NT) {
    const int which = q / (STAGES * BK), rem = q % (STAGES * BK);
    const int k = kbase + rem;                                   // tile (rem / BK) sits in slot (rem / BK)
    const long long o = (rem / BK < nkt && k < K) ? lvl_off_ni(slv[which ? 2 : 1], k) : -1;
    (which ? kofB : kofA)[rem] = o;
  }
  __syncthreads();

  const bool avec = d.a_vec != 0, bvec = d.b_vec != 0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) {
      load_tile<BM, NT, AKF>(As + s * A_ELEMS, gA, rowA, kofA + s * BK, avec, tid);
      load_tile<BN, NT, BKF>(Bs + s * B_ELEMS, gB, rowB, kofB + s * BK, bvec, tid);
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
        load_tile<BM, NT, AKF>(As + s * A_ELEMS, gA, rowA, kofA + s * BK, avec, tid);
        load_tile<BN, NT, BKF>(Bs + s * B_ELEMS, gB, rowB, kofB + s * BK, bvec, tid);
      }
      cp_async_commit();
      // k-offsets of tile kt+STAGES go into the slot tile kt used (its loads were issued two iterations ago);
      // they are read after the next iteration's barrier
      const int fk = kt + STAGES;
      if (tid < 2 * BK && fk < nkt) {
        const int which = tid / BK, kl = tid % BK;
        const int k = kbase + fk * BK + kl;
        (which ? kofB : kofA)[(fk % STAGES) * BK + kl] = k < K ? lvl_off_ni(slv[which ? 2 : 1], k) : -1;
      }
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
  if (S > 1) {
    // split-K: raw partial tile to the workspace [batch][split][N][M]; k_splitk_reduce applies alpha/beta
    double* ws = d.ws + ((long long)blockIdx.z) * d.M * (long long)d.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + wm0 + i * 8 + g;
      if (m < d.M) {
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = n0 + wn0 + j * 8 + 2 * t + e;
            if (n < d.N) ws[m + (long long)d.M * n] = acc[i][j][e];
          }
      }
    }
    return;
  }
  double* gC = d.C + lvl_off_ni(slv[8], bz);
  const double alpha = d.alpha, beta = d.beta;
  long long coff[TN][2];
  bool cok[TN][2];
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n0 + wn0 + j * 8 + 2 * t + e;
      cok[j][e] = n < d.N;
      coff[j][e] = cok[j][e] ? lvl_off_ni(slv[5], n) : 0;
    }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + wm0 + i * 8 + g;
    if (m < d.M) {
      const long long ro = lvl_off_ni(slv[4], m);
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

// C = alpha * sum_s ws[b][s] + beta * C   (fixed summation order: deterministic)
__global__ void k_splitk_reduce(const __grid_constant__ GemmDesc d) {
  const long long MN = (long long)d.M * d.N;
  const long long total = MN * d.batch;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / MN);
    const long long r = idx - (long long)b * MN;
    const int m = (int)(r % d.M), n = (int)(r / d.M);
    double v = 0.0;
    for (int s = 0; s < d.splitk; ++s) v += d.ws[((long long)b * d.splitk + s) * MN + r];
    double* pc = d.C + lvl_off(d.cb, b) + lvl_off(d.cm, m) + lvl_off(d.cn, n);
    v *= d.alpha;
    if (d.beta != 0.0) v += d.beta * *pc;
    *pc = v;
  }
}

template <int BM, int BN, int WM, int WN, bool AKF, bool BKF>
void launch_cfg(tnad_ctx* c, const GemmDesc& d) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int A_ELEMS = AKF ? BM * (BK + 4) : BK * (BM + 4);
  constexpr int B_ELEMS = BKF ? BN * (BK + 4) : BK * (BN + 4);
  const size_t smem = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(double) +
                      (BM + BN + 2 * STAGES * BK) * sizeof(long long);
  static std::atomic<unsigned long long> attr_devs{0};   // kernel attributes are per device: one bit per device id (set after the attribute call: a racing thread at worst repeats it)
  const bool attr_set = (attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL;
  auto kern = gemm_dmma_kernel<BM, BN, WM, WN, AKF, BKF>;
  if (!attr_set) {
    TNAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const int S = d.splitk > 1 ? d.splitk : 1;
  dim3 grid((d.M + BM - 1) / BM, (d.N + BN - 1) / BN, d.batch * S);
  KTimer kt(c, KF_GEMM);
  kern<<<grid, NT, smem, c->stream>>>(d);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  if (S > 1) {
    const long long total = (long long)d.M * d.N * d.batch;
    int nb = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    k_splitk_reduce<<<nb < 1 ? 1 : nb, 256, 0, c->stream>>>(d);
    c->launches++;
    TNAD_CUDA(cudaGetLastError());
  }
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

void gemm_run(tnad_ctx* c, const GemmDesc& d0) {
  if (d0.M <= 0 || d0.N <= 0 || d0.batch <= 0) return;
  const double flops = 2.0 * d0.M * (double)d0.N * d0.K * d0.batch;
  if (c->ktiming) c->gemm_flops += flops;
  if (gemm_tma_try(c, d0)) {
    c->gemm_tma_n++;
    if (c->ktiming) c->gemm_tma_flops += flops;
    return;
  }
  c->gemm_fallback_n++;
  if (opt_i(c, "TNAD_GEMM_DEBUG", 0))
    fprintf(stderr, "[tnad gemm] cp.async kernel for M=%d N=%d K=%d batch=%d (akf %d bkf %d)\n", d0.M, d0.N, d0.K, d0.batch, d0.a_kfast, d0.b_kfast);
  GemmDesc d = d0;
  bool large = d.M >= 96 && d.N >= 96;
  {
    // short-K products with few 128x128 tiles (trailing updates of the tridiagonalisation, compact-WY updates): the
    // 64x64 tiling gives 4x the CTAs for the same bytes; A/B knob TNAD_GEMM_SMALLK=<K limit> (0 = off)
    const int smallk = opt_i(c, "TNAD_GEMM_SMALLK", 128);   // measured at n = 2048: trailing updates 3.3 -> 2.3 ms, back-transform 2.5 -> 2.2 ms per decomposition
    const long long t128 = (long long)((d.M + 127) / 128) * ((d.N + 127) / 128) * d.batch;
    if (large && smallk > 0 && d.K <= smallk && t128 < 2LL * c->num_sms) large = false;
  }
  const int bm = large ? 128 : 64;
  const long long tiles = (long long)((d.M + bm - 1) / bm) * ((d.N + bm - 1) / bm) * d.batch;
  // split-K when the output tiles alone cannot fill the machine (small M x N, long K: the projector and
  // svd_back products of the CTMRG step); partial tiles are summed in a fixed order by k_splitk_reduce
  const int nkt = (d.K + BK - 1) / BK;
  int S = 1;
  if (tiles < c->num_sms && nkt >= 8) {
    S = (int)std::min<long long>((2LL * c->num_sms + tiles - 1) / tiles, nkt / 4);
    S = std::max(1, std::min(S, 64));
  }
  TNAD_REQUIRE((long long)d.batch * S <= 65535 && (d.N + 63) / 64 <= 65535, "gemm: grid too large");
  Tens ws;
  d.splitk = S;
  d.ws = nullptr;
  if (S > 1) {
    ws = t_alloc(c, {(int64_t)d.M * d.N * d.batch * S});
    d.ws = ws.p;
  }
  if (large) launch_layout<128, 128, 64, 32>(c, d);
  else launch_layout<64, 64, 32, 32>(c, d);
}

}  // namespace tnad
