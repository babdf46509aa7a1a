// Householder tridiagonalisation of a dense symmetric matrix and the matching back-transformation: the O(n^3)
// direct route to the eigen-decomposition ctmrgstep needs (ctmrg.jl:133-136), next to the block-Jacobi solver of
// symeig.cu (which spends ~20 sweeps x 8 n^3 flops; this path needs ~(4/3 + 2 + 2) n^3).
//
//   A = Q T Q',  Q = H_0 H_1 ... H_{n-3},  H_j = I - tau_j v_j v_j'
//
// k_sytrd_panel is ONE persistent cooperative kernel per panel of nb columns (dlatrd-style blocking, published
// LAPACK algorithm): all CTAs stay resident and meet at a global barrier twice per column.
//   phase B  every warp takes trailing columns c and forms p_c = A22[:, c] . v   (A22 is symmetric and stored in
//            full, so the matrix-vector product is a set of independent, coalesced column dots - no atomics, no
//            cross-CTA partial sums, bit-reproducible); the 2i panel columns V, W are dotted with v in the same pass.
//   phase C  (rows cyclic over warps) w = tau (p - V W'v - W V'v) - 1/2 tau^2 (p'v) v, then the NEXT column of A
//            is brought up to date with the pending rank-2i update and its norm is reduced for the next reflector.
// The rank-2nb trailing update A22 -= V W' + W V' between panels is one DMMA GEMM (contract) with K = 2 nb.
// The dominant cost is phase B: 8 n^3 / 3 bytes of matrix reads (HBM/L2 bound), see DESIGN.md.
#include "drivers.h"
#include "eigdc.h"
#include <algorithm>
#include <numeric>

namespace tnad {

namespace {

constexpr int ST_NT = 512;          // threads per CTA of the panel kernel (one CTA per SM)
constexpr int ST_NW = ST_NT / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All CTAs of a cooperative launch meet here: CTA b publishes `epoch` in its own flag line (release) and every CTA
// polls all flags (one acquire load per thread) - no contended atomic, one L2 round trip after the last arrival.
__device__ __forceinline__ void grid_barrier(unsigned int* flags, unsigned int epoch, int G, int* err, int* s_to) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(flags + 32 * blockIdx.x), "r"(epoch) : "memory");
  const long long t0 = clock64();
  for (;;) {
    int bad = 0;
    for (int q = threadIdx.x; q < G; q += ST_NT) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * q) : "memory");
      bad |= ((int)(v - epoch) < 0) ? 1 : 0;
    }
    if (threadIdx.x == 0 && clock64() - t0 > 6000000000LL) {   // ~3 s: never hang the device on a lost CTA
      *s_to = 1;
      *err = 1;
    }
    if (__syncthreads_count(bad) == 0 || *s_to) break;
  }
}

// block-wide sum, identical result in every thread; red: >= ST_NW doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < ST_NW; ++w) s += red[w];
  return s;
}

// P1 = [V W], P2 = [W V]  (n x 2nb each, leading dimension ldp): A22 -= P1 P2' after the panel.
// Work item of phase B = (column, chunk of ST_CH rows): 16 independent loads per lane in flight; chunk partials
// land in ppart[chunk][column] / spart[chunk][slot] and are summed (fixed order) by their consumers in phase C.
// Loads that do not depend on the other CTAs (trailing-matrix chunks, panel rows) are issued BEFORE the barrier
// they would otherwise wait behind, so the critical path per column is two barriers and two L2 round trips.
constexpr int ST_CH = 512;
constexpr int ST_PJ = 64;   // max chunks (n <= ST_CH * ST_PJ)
__global__ void __launch_bounds__(ST_NT, 1) k_sytrd_panel(double* A, long long lda, int n, int j0, int nbc, int nb, double* P1,
                                                          double* P2, long long ldp, double* Vh, long long ldv, double* tau,
                                                          double* dd, double* ee, double* ppart, double* spart,
                                                          double* pvpart, unsigned int* flags, unsigned int epoch0, int* err, int dbg) {
  extern __shared__ double sm[];
  double* vs = sm;                 // v, indexed by absolute row
  double* sv = sm + n;             // V'v   (nb)
  double* sw = sv + nb;            // W'v   (nb)
  double* vj = sw + nb;            // V[jn, :]
  double* wj = vj + nb;            // W[jn, :]
  double* pjs = wj + nb;           // chunk partials of p[jn]  (ST_PJ)
  double* red = pjs + ST_PJ;       // ST_NW + 8
  __shared__ int s_to;
  const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int gw = b * ST_NW + warp, GW = G * ST_NW;
  unsigned int epoch = epoch0;
  if (t == 0) s_to = 0;

  for (int i = 0; i < nbc; ++i) {
    const int j = j0 + i, jn = j + 1;
    const int m = n - jn, ncol = m + 2 * i;
    const int nch = (m + ST_CH - 1) / ST_CH;
    const long long nitem = (long long)ncol * nch;
    // ---- prefetch the first phase-B item if it is a chunk of the trailing matrix (independent of the barrier) -----
    double x[16];
    const bool pre = gw < nitem && (gw / nch) < m;
    if (pre) {
      const int cc = gw / nch, q = gw - cc * nch;
      const int rb = jn + q * ST_CH + lane, rend = min(n, jn + (q + 1) * ST_CH);
      const double* col = A + (long long)(jn + cc) * lda;
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? col[rb + 32 * u] : 0.0;
    }
    if (i > 0 && !(dbg & 2)) grid_barrier(flags, ++epoch, G, err, &s_to);   // column j of A is up to date
    // ---- reflector (every CTA, identical arithmetic, norm from the shared-memory copy) ---------------------------
    double ss = 0.0;
    for (int r = jn + t; r < n; r += ST_NT) {
      const double xr = __ldcg(A + r + (long long)j * lda);
      vs[r] = xr;
      if (r > jn) ss += xr * xr;
    }
    const double xn2 = block_sum(ss, red);
    const double alpha = vs[jn];
    double beta, tauj, scale;
    if (xn2 == 0.0) {
      beta = alpha;
      tauj = 0.0;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      tauj = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (int r = jn + t; r < n; r += ST_NT) vs[r] = (r == jn) ? 1.0 : vs[r] * scale;
    __syncthreads();
    for (int r = jn + b * ST_NT + t; r < n; r += G * ST_NT) {
      const double v = vs[r];
      P1[r + (long long)i * ldp] = v;
      P2[r + (long long)(nb + i) * ldp] = v;
      Vh[r + (long long)j * ldv] = v;
    }
    if (b == 0 && t == 0) {
      ee[j] = beta;
      tau[j] = tauj;
      dd[j] = __ldcg(A + j + (long long)j * lda);
    }
    // ---- phase B: chunk partials of p = A22 v and of the panel dots V'v, W'v --------------------------------------
    double pvacc = 0.0;
    for (long long it = gw; it < ((dbg & 1) ? 0 : nitem); it += GW) {
      const int cc = (int)(it / nch), q = (int)(it - (long long)cc * nch);
      const int rb = jn + q * ST_CH + lane, rend = min(n, jn + (q + 1) * ST_CH);
      if (!(pre && it == gw)) {
        if (cc < m) {
          const double* col = A + (long long)(jn + cc) * lda;   // never written inside this kernel
#pragma unroll
          for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? col[rb + 32 * u] : 0.0;
        } else {
          const int kk = cc - m;
          const double* col = P1 + (long long)(kk < i ? kk : nb + (kk - i)) * ldp;   // written by other CTAs: bypass L1
#pragma unroll
          for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? __ldcg(col + rb + 32 * u) : 0.0;
        }
      }
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int u = 0; u < 16; u += 2) {
        s0 += x[u] * ((rb + 32 * u < rend) ? vs[rb + 32 * u] : 0.0);
        s1 += x[u + 1] * ((rb + 32 * (u + 1) < rend) ? vs[rb + 32 * (u + 1)] : 0.0);
      }
      const double s = warp_sum(s0 + s1);
      if (cc < m) {
        if (lane == 0) ppart[(long long)q * n + jn + cc] = s;
        pvacc += s * vs[jn + cc];
      } else if (lane == 0) {
        const int kk = cc - m;
        spart[q * 2 * nb + (kk < i ? kk : nb + (kk - i))] = s;
      }
    }
    __syncthreads();
    if (lane == 0) red[warp] = pvacc;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int w = 0; w < ST_NW; ++w) s += red[w];
      pvpart[b] = s;
    }
    // ---- prefetch for phase C: this warp's first row of the panels and of column jn (written in earlier columns) --
    const bool next = (i + 1 < nbc);
    const int rfirst = jn + gw;
    double pv_r = 0.0, pw_r = 0.0, pa_r = 0.0;
    if (rfirst < n) {
      if (lane < i) {
        pv_r = __ldcg(P1 + rfirst + (long long)lane * ldp);
        pw_r = __ldcg(P1 + rfirst + (long long)(nb + lane) * ldp);
      }
      if (next) pa_r = __ldcg(A + rfirst + (long long)jn * lda);
    }
    if (!(dbg & 2)) grid_barrier(flags, ++epoch, G, err, &s_to);
    // ---- phase C: w, then column jn of A brought up to date -----------------------------------------------------
    if (t < 2 * i) {
      const int slot = t < i ? t : nb + (t - i);
      double s = 0.0;
      for (int q = 0; q < nch; ++q) s += __ldcg(spart + q * 2 * nb + slot);
      if (t < i) {
        sv[t] = s;
        vj[t] = __ldcg(P1 + jn + (long long)t * ldp);
      } else {
        sw[t - i] = s;
        wj[t - i] = __ldcg(P1 + jn + (long long)(nb + t - i) * ldp);
      }
    } else if (t >= 64 && t < 64 + nch) {
      pjs[t - 64] = __ldcg(ppart + (long long)(t - 64) * n + jn);
    }
    double pv = 0.0;
    for (int q = t; q < G; q += ST_NT) pv += __ldcg(pvpart + q);
    pv = block_sum(pv, red);   // (also orders the sv/sw/vj/wj/pjs stores)
    double cross = 0.0, pcj = 0.0, p_jn = 0.0;
    if (lane < i) {
      cross = sv[lane] * sw[lane];
      pcj = vj[lane] * sw[lane] + wj[lane] * sv[lane];
    }
    cross = warp_sum(cross);
    pcj = warp_sum(pcj);
    for (int q = 0; q < nch; ++q) p_jn += pjs[q];
    const double pcv = pv - 2.0 * cross;
    const double c1 = tauj, c2 = 0.5 * tauj * tauj * pcv;
    const double w_jn = c1 * (p_jn - pcj) - c2;
    for (int r = rfirst; r < ((dbg & 4) ? 0 : n); r += GW) {
      double vr = 0.0, wr = 0.0, ar = 0.0, pr = 0.0;
      for (int q = lane; q < nch; q += 32) pr += __ldcg(ppart + (long long)q * n + r);
      if (r == rfirst) {
        vr = pv_r;
        wr = pw_r;
        ar = pa_r;
      } else {
        if (lane < i) {
          vr = __ldcg(P1 + r + (long long)lane * ldp);
          wr = __ldcg(P1 + r + (long long)(nb + lane) * ldp);
        }
        if (next) ar = __ldcg(A + r + (long long)jn * lda);
      }
      double acc = 0.0, acc2 = 0.0;
      if (lane < i) {
        acc = vr * sw[lane] + wr * sv[lane];
        acc2 = vr * wj[lane] + wr * vj[lane];
      }
      acc = warp_sum(acc);
      acc2 = warp_sum(acc2);
      pr = warp_sum(pr);
      const double w = c1 * (pr - acc) - c2 * vs[r];
      if (lane == 0) {
        P1[r + (long long)(nb + i) * ldp] = w;
        P2[r + (long long)i * ldp] = w;
        if (next) A[r + (long long)jn * lda] = ar - acc2 - (vs[r] * w_jn + w);
      }
    }
  }
}

__global__ void k_sytrd_tail(const double* A, long long lda, int n, double* dd, double* ee) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (n >= 2) {
      dd[n - 2] = A[(n - 2) + (long long)(n - 2) * lda];
      ee[n - 2] = A[(n - 1) + (long long)(n - 2) * lda];
    }
    dd[n - 1] = A[(n - 1) + (long long)(n - 1) * lda];
  }
}

// dst = src (+ src') for an n x n strided source; dst column-major with leading dimension ldd
__global__ void k_load_symm(double* dst, long long ldd, const double* __restrict__ A, long long s0, long long s1, long long n,
                            int add_t) {
  const long long total = n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % n, j = idx / n;
    double v;
    if (add_t) v = A[i * s0 + j * s1] + A[j * s0 + i * s1];
    else v = 0.5 * (A[i * s0 + j * s1] + A[j * s0 + i * s1]);
    dst[i + j * ldd] = v;
  }
}

// Compact WY factor of one panel (dlarft, forward / columnwise): T upper triangular, H_0..H_{k-1} = I - V T V'.
// G = V'V (k x k, ldg), one CTA.
__global__ void k_larft(const double* __restrict__ G, int ldg, const double* __restrict__ tau, int k, double* T, int ldt) {
  extern __shared__ double sh[];
  G += (long long)blockIdx.x * ldg * k;   // one CTA per panel: G, T are (k, k, panel), tau is (k, panel)
  T += (long long)blockIdx.x * ldt * k;
  tau += (long long)blockIdx.x * k;
  double* Ts = sh;            // k x k
  double* col = sh + k * k;   // k
  const int t = threadIdx.x;
  for (int idx = t; idx < k * k; idx += blockDim.x) Ts[idx] = 0.0;
  __syncthreads();
  for (int i = 0; i < k; ++i) {
    const double ti = tau[i];
    // col[0:i] = -ti * Ts[0:i, 0:i] * G[0:i, i]
    if (t < i) {
      double s = 0.0;
      for (int q = t; q < i; ++q) s += Ts[t + q * k] * G[q + i * ldg];   // Ts upper triangular: q >= t
      col[t] = -ti * s;
    }
    __syncthreads();
    if (t < i) Ts[t + i * k] = col[t];
    if (t == 0) Ts[i + i * k] = ti;
    __syncthreads();
  }
  for (int idx = t; idx < k * k; idx += blockDim.x) T[(idx % k) + (long long)(idx / k) * ldt] = Ts[idx];
}

__global__ void k_gather_cols(const double* __restrict__ Q, long long ldq, const int* __restrict__ perm,
                              const double* __restrict__ sgn, long long n, double* __restrict__ U, double* __restrict__ V) {
  const long long j = blockIdx.x;
  const double* src = Q + (long long)perm[j] * ldq;
  const double s = sgn[j];
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = src[i];
    U[i + j * n] = v;
    V[i + j * n] = v * s;
  }
}

int env_i(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

Tens view2(double* p, int64_t rows, int64_t cols, int64_t ld) {
  Tens t = t_wrap(p, {rows, cols});
  t.str[0] = 1;
  t.str[1] = ld;
  return t;
}

}  // namespace

// A (n x n, leading dimension lda, full symmetric storage) is overwritten; Vh (n x n, zero-initialised by the
// caller) receives v_j in column j (rows > j, v[j+1] = 1); tau, dd (n), ee (n-1).
void sytrd(tnad_ctx* c, double* A, int64_t lda, int64_t n, double* Vh, int64_t ldv, double* tau, double* dd, double* ee) {
  TNAD_REQUIRE(n >= 1, "sytrd: empty matrix");
  const int nb = 32;
  const int64_t nref = n >= 3 ? n - 2 : 0;
  cudaStream_t st = c->stream;
  if (nref > 0) {
    const size_t smem = (size_t)(n + 4 * nb + ST_PJ + ST_NW + 16) * sizeof(double);
    TNAD_REQUIRE(n <= (int64_t)ST_CH * ST_PJ, "sytrd: n too large");
    TNAD_REQUIRE(smem <= 220 * 1024, "sytrd: matrix too large for the shared-memory reflector (n <= 28000)");
    TNAD_CUDA(cudaFuncSetAttribute(k_sytrd_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TNAD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sytrd_panel, ST_NT, smem));
    TNAD_REQUIRE(per_sm >= 1, "sytrd: panel kernel does not fit on an SM");
    const int G = c->num_sms;
    const int64_t ldp = n;
    Tens P1 = t_alloc(c, {ldp, 2 * nb}, true), P2 = t_alloc(c, {ldp, 2 * nb}, true);
    const int64_t nchmax = (n + ST_CH - 1) / ST_CH;
    Tens ppart = t_alloc(c, {n, nchmax}), spart = t_alloc(c, {2 * nb, nchmax}), pvpart = t_alloc(c, {G});
    Tens ctl = t_alloc(c, {16 * (int64_t)G + 2}, true);   // one 128-byte flag line per CTA, then the error flag
    unsigned int* bar = reinterpret_cast<unsigned int*>(ctl.p);
    int* err = reinterpret_cast<int*>(ctl.p + 16 * (int64_t)G);
    unsigned int bar_base = 0;
    int dbg = env_i("TNAD_SYTRD_DBG", 0);
    for (int64_t j0 = 0; j0 < nref; j0 += nb) {
      int nbc = (int)std::min<int64_t>(nb, nref - j0);
      int ni = (int)n, j0i = (int)j0, nbi = nb;
      long long lda_ = lda, ldp_ = ldp, ldv_ = ldv;
      double *P1p = P1.p, *P2p = P2.p, *pp = ppart.p, *sp = spart.p, *pvp = pvpart.p;
      void* args[] = {&A, &lda_, &ni, &j0i, &nbc, &nbi, &P1p, &P2p, &ldp_, &Vh, &ldv_, &tau, &dd, &ee, &pp, &sp, &pvp, &bar, &bar_base, &err, &dbg};
      {
        KTimer kt(c, KF_EIG);
        TNAD_CUDA(cudaLaunchCooperativeKernel((void*)k_sytrd_panel, dim3(G), dim3(ST_NT), args, smem, st));
      }
      c->launches++;
      bar_base += (unsigned int)(2 * nbc);   // barrier epochs consumed by this launch
      // trailing update A22 -= [V W] [W V]'
      const int64_t jt = j0 + nbc, nt = n - jt;
      if (nt > 0) {
        Tens A22 = view2(A + jt + jt * lda, nt, nt, lda);
        Tens L = view2(P1.p + jt, nt, 2 * nb, ldp), R = view2(P2.p + jt, nt, 2 * nb, ldp);
        if (nbc < nb) {   // unused panel columns of a short last panel must not contribute
          for (int q = nbc; q < nb; ++q) {
            TNAD_CUDA(cudaMemsetAsync(P1.p + (int64_t)q * ldp, 0, (size_t)n * sizeof(double), st));
            TNAD_CUDA(cudaMemsetAsync(P1.p + (int64_t)(nb + q) * ldp, 0, (size_t)n * sizeof(double), st));
          }
        }
        contract(c, "ik,jk->ij", L, R, A22, -1.0, 1.0);
      }
    }
    int herr = 0;
    TNAD_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    sync(c);
    if (herr) fail(TNAD_ERR_INTERNAL, "sytrd: grid barrier timed out");
  }
  k_sytrd_tail<<<1, 32, 0, st>>>(A, lda, (int)n, dd, ee);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
}

// Number of columns the reflector store Vh (and tau) must have: panels of the back-transformation are read as a
// batch of equal width, so the column count is rounded up (extra columns stay zero, tau = 0).
int64_t sytrd_vcols(int64_t n) { return (n + 127) / 128 * 128; }

// Z[0:n, 0:ncols] <- Q Z with Q = H_0 ... H_{n-3} from sytrd (panels applied last to first, compact WY):
// Gram matrices, T factors and V T of ALL panels come from three batched launches; each panel then costs two GEMMs.
void apply_q(tnad_ctx* c, const double* Vh, int64_t ldv, const double* tau, int64_t n, double* Z, int64_t ldz, int64_t ncols) {
  const int64_t nref = n >= 3 ? n - 2 : 0;
  if (nref == 0) return;
  const int kb = env_i("TNAD_APPLYQ_NB", n >= 4096 ? 128 : 64) >= 128 ? 128 : 64;
  const int64_t npan = (nref + kb - 1) / kb;
  TNAD_REQUIRE(npan * kb <= sytrd_vcols(n), "apply_q: reflector store too narrow");
  TNAD_CUDA(cudaFuncSetAttribute(k_larft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kb * kb + kb) * sizeof(double))));
  double* V = const_cast<double*>(Vh);
  Tens Vall;   // (row, panel, column in panel)
  Vall.p = V;
  Vall.rank = 3;
  Vall.dim[0] = n; Vall.dim[1] = npan; Vall.dim[2] = kb;
  Vall.str[0] = 1; Vall.str[1] = (int64_t)kb * ldv; Vall.str[2] = ldv;
  Tens Gall = t_alloc(c, {kb, kb, npan}), Tall = t_alloc(c, {kb, kb, npan}), VT = t_alloc(c, {n, kb, npan});
  contract(c, "rpi,rpj->ijp", Vall, Vall, Gall);
  k_larft<<<(int)npan, 128, (kb * kb + kb) * sizeof(double), c->stream>>>(Gall.p, kb, tau, kb, Tall.p, kb);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  contract(c, "rpi,ijp->rjp", Vall, Tall, VT);
  Tens Y = t_alloc(c, {kb, ncols});
  for (int64_t pi = npan - 1; pi >= 0; --pi) {
    const int64_t j0 = pi * kb, jr = j0 + 1, rows = n - jr;
    Tens Vp = view2(V + jr + j0 * ldv, rows, kb, ldv);
    Tens VTp = view2(VT.p + jr + pi * (int64_t)kb * n, rows, kb, n);
    Tens Zr = view2(Z + jr, rows, ncols, ldz);
    contract(c, "ki,kj->ij", Vp, Zr, Y);
    contract(c, "ik,kj->ij", VTp, Y, Zr, -1.0, 1.0);
  }
}

// Eigen-decomposition route: tridiagonalise, divide and conquer, back-transform.
SvdResult svd_symmetric_dc(tnad_ctx* c, const Tens& A, bool sym_add_transpose) {
  TNAD_REQUIRE(A.rank == 2 && A.dim[0] == A.dim[1], "svd_symmetric_dc: need a square matrix");
  const int64_t n = A.dim[0];
  cudaStream_t st = c->stream;
  Tens Aw = t_alloc(c, {n, n});
  const int nbk = (int)std::max<long long>(1, std::min<long long>((n * n + 1023) / 1024, 148 * 8));
  k_load_symm<<<nbk, 256, 0, st>>>(Aw.p, n, A.p, A.str[0], A.str[1], n, sym_add_transpose ? 1 : 0);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  Tens Vh = t_alloc(c, {n, sytrd_vcols(n)}, true), tau = t_alloc(c, {sytrd_vcols(n)}, true), dd = t_alloc(c, {n}), ee = t_alloc(c, {n}, true);
  const bool debug = env_i("TNAD_DC_DEBUG", 0) != 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (debug)
    for (auto& e : ev) e = get_event(c);
  if (debug) TNAD_CUDA(cudaEventRecord(ev[0], st));
  sytrd(c, Aw.p, n, n, Vh.p, n, tau.p, dd.p, ee.p);
  if (debug) TNAD_CUDA(cudaEventRecord(ev[1], st));
  Tens lam, Z;
  int64_t N = 0;
  stedc(c, dd.p, ee.p, n, lam, Z, N);   // Z: N x N (ld N), lam: N, pads carry eigenvalues above the spectrum
  if (debug) TNAD_CUDA(cudaEventRecord(ev[2], st));
  apply_q(c, Vh.p, n, tau.p, n, Z.p, N, N);
  if (debug) {
    TNAD_CUDA(cudaEventRecord(ev[3], st));
    TNAD_CUDA(cudaEventSynchronize(ev[3]));
    float a = 0, b = 0, d3 = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    cudaEventElapsedTime(&d3, ev[2], ev[3]);
    fprintf(stderr, "[tnad dc] n=%lld N=%lld sytrd %.2f ms  stedc %.2f ms  apply_q %.2f ms\n", (long long)n, (long long)N, a, b, d3);
    for (auto& e : ev) c->event_pool.push_back(e);
  }
  std::vector<double> lh((size_t)N);
  d2h(c, lh.data(), lam.p, (size_t)N);
  // the N - n pad eigenvalues are the largest ones (stedc puts them above 3 |T|)
  std::vector<int> idx((size_t)N);
  for (int64_t i = 0; i < N; ++i) idx[(size_t)i] = (int)i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lh[x] < lh[y]; });
  std::vector<int> perm(idx.begin(), idx.begin() + n);
  std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return std::fabs(lh[x]) > std::fabs(lh[y]); });
  std::vector<double> sval((size_t)n), sgn((size_t)n);
  double fro2 = 0.0;
  for (int64_t r = 0; r < n; ++r) {
    sval[r] = std::fabs(lh[perm[r]]);
    sgn[r] = lh[perm[r]] < 0.0 ? -1.0 : 1.0;
    fro2 += sval[r] * sval[r];
  }
  Tens meta = t_alloc(c, {3 * n + 4});
  int* dperm = reinterpret_cast<int*>(meta.p);
  double* dsgn = meta.p + (n + 1) / 2 + 1;
  double* dsval = dsgn + n;
  TNAD_CUDA(cudaMemcpyAsync(dperm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  TNAD_CUDA(cudaMemcpyAsync(dsgn, sgn.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
  TNAD_CUDA(cudaMemcpyAsync(dsval, sval.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
  SvdResult res;
  res.U = t_alloc(c, {n, n});
  res.V = t_alloc(c, {n, n});
  res.S = t_alloc(c, {n});
  k_gather_cols<<<(int)n, 128, 0, st>>>(Z.p, N, dperm, dsgn, n, res.U.p, res.V.p);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  TNAD_CUDA(cudaMemcpyAsync(res.S.p, dsval, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  sync(c);
  res.s_host = sval;
  res.null_thr = 16.0 * 2.220446049250313e-16 * std::sqrt(fro2);   // same absolute level as the Jacobi solver
  res.sweeps = 0;
  return res;
}

SvdResult svd_symmetric_auto(tnad_ctx* c, const Tens& A, bool sym_add_transpose, const Tens* Q0) {
  const int mode = env_i("TNAD_SYMEIG", -1);
  const int64_t n = A.dim[0];
  const bool dc = mode == 2 || (mode < 0 && n >= env_i("TNAD_DC_MIN", 256));
  return dc ? svd_symmetric_dc(c, A, sym_add_transpose) : svd_symmetric(c, A, sym_add_transpose, Q0);
}

}  // namespace tnad
