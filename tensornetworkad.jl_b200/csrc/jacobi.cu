// GPU-resident SVD: one-sided block Jacobi (Hestenes) with FP64 tensor-core panel updates.
//
// Replaces `LinearAlgebra.svd` (LAPACK dgesdd) at trg.jl:36 and ctmrg.jl:136.  The working matrix
// G (m x N, N = n padded to a multiple of 64) starts as A (or A + A^T, ctmrg.jl:135, fused into the
// load) and V as the identity.  Columns are grouped in blocks of 32; a sweep visits every block pair
// once in round-robin order, all N/64 pairs of a round in parallel:
//   1. k_jacobi_gram   H = X^T X for X = [G_I G_J] (64 columns), rows split over CTAs, DMMA.
//   2. k_pivot_eig<GRAM> (symeig.cu) sums the partial Grams and diagonalises the 64x64 H with the
//                      register-resident implicit Jacobi kernel (32 disjoint rotations per step) -> W;
//                      a pair whose largest cosine is already below tol is skipped.
//   3. k_jacobi_update [G_I G_J] <- [G_I G_J] W and [V_I V_J] <- [V_I V_J] W, in place, DMMA.
// At convergence G = A V has orthogonal columns: S = column norms, U = G / S.  Exactly-null columns
// get an orthonormal completion (V's own column for the symmetric case, projected random vectors
// otherwise) so that the full-size U, V needed by svd_back (trg.jl:72-105) are orthonormal.
// Rows are padded to a multiple of 128 and columns to a multiple of 64 with zeros, so no kernel has
// a tail; zero columns are never rotated (their Gram entries are exactly 0).
#include "common.h"
#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace tnad {

void launch_gram_pivot_eig(tnad_ctx* c, const double* Hpart, int nsplit, int npairs, int p, int round, int nreal,
                           const double* fro2, double tol_rel, double nullfac, int max_inner, int cross, double* Wbuf,
                           int* skip, unsigned long long* offbits, cudaStream_t st);   // symeig.cu

namespace {

constexpr int JB = 32;        // block width
constexpr int JP = 2 * JB;    // columns per pair
constexpr int RC = 128;       // rows per chunk
constexpr int XLD = RC + 4;   // smem leading dimension of a slab column (bank-conflict-free fragments)
constexpr int WLD = JP + 4;
constexpr int HLD = JP + 1;

#define LAUNCH_CHECK(c)            \
  do {                             \
    (c)->launches++;               \
    TNAD_CUDA(cudaGetLastError()); \
  } while (0)

__host__ __device__ inline void rr_pair(int p, int r, int k, int& a, int& b) {
  // round-robin tournament: p even players, round r in [0, p-1), table k in [0, p/2)
  if (k == 0) {
    a = p - 1;
    b = r;
  } else {
    a = (r + k) % (p - 1);
    b = (r - k + (p - 1)) % (p - 1);
  }
  if (a > b) {
    int t = a;
    a = b;
    b = t;
  }
}

__device__ __forceinline__ void cp_async16(double* s, const double* g) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// slab: RC rows x 64 columns (block I then block J) -> Xs[col * XLD + row]
__device__ __forceinline__ void load_slab(double* Xs, const double* __restrict__ P, long long ld, int cI, int cJ,
                                          int tid) {
#pragma unroll
  for (int it = 0; it < (JP * RC / 2) / 256; ++it) {
    const int q = tid + it * 256;
    const int col = q / (RC / 2), r2 = q % (RC / 2);
    const long long gcol = col < JB ? cI + col : cJ + col - JB;
    cp_async16(Xs + col * XLD + 2 * r2, P + gcol * ld + 2 * r2);
  }
}

// ---- 1. partial Gram matrices -------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jacobi_gram(const double* __restrict__ G, long long ldg, int mchunks, int p,
                                                     int round, int nsplit, double* __restrict__ Hpart) {
  extern __shared__ __align__(16) double Xs[];
  const int pair = blockIdx.x, split = blockIdx.y, tid = threadIdx.x;
  int I, J;
  rr_pair(p, round, pair, I, J);
  const int cI = I * JB, cJ = J * JB;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wy = warp & 3, wx = warp >> 2;   // rows 16*wy, cols 32*wx of H
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int chunk = split; chunk < mchunks; chunk += nsplit) {
    __syncthreads();
    load_slab(Xs, G + (long long)chunk * RC, ldg, cI, cJ, tid);
    cp_async_wait_all();
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < RC / 4; ++kk) {
      const int k = kk * 4 + t;
      double af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = Xs[(16 * wy + 8 * i + g) * XLD + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Xs[(32 * wx + 8 * j + g) * XLD + k];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  double* out = Hpart + ((long long)pair * nsplit + split) * (JP * JP);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = 16 * wy + 8 * i + g, n = 32 * wx + 8 * j + 2 * t + e;
        out[m + JP * n] = acc[i][j][e];
      }
}

// ---- 3. rotate the column pair of G and V -------------------------------------------------------
__global__ void __launch_bounds__(256) k_jacobi_update(double* __restrict__ G, long long ldg, int mchunks,
                                                       double* __restrict__ V, long long ldv, int p, int round,
                                                       const double* __restrict__ Wbuf, const int* __restrict__ skip,
                                                       unsigned long long* work_counter) {
  extern __shared__ __align__(16) double sm[];
  double* Xs = sm;
  double* Ws = sm + JP * XLD;
  const int pair = blockIdx.x, tid = threadIdx.x;
  if (skip[pair]) return;
  if (work_counter && tid == 0) atomicAdd(work_counter, 1ULL);   // executed slabs (roofline accounting)
  int I, J;
  rr_pair(p, round, pair, I, J);
  const int cI = I * JB, cJ = J * JB;
  int chunk = blockIdx.y;
  double* P;
  long long ld;
  if (chunk < mchunks) {
    P = G + (long long)chunk * RC;
    ld = ldg;
  } else {
    P = V + (long long)(chunk - mchunks) * RC;
    ld = ldv;
  }
  const double* wsrc = Wbuf + (long long)pair * (JP * JP);
#pragma unroll
  for (int it = 0; it < (JP * JP / 2) / 256; ++it) {
    const int q = tid + it * 256;
    const int n = q / (JP / 2), k2 = q % (JP / 2);
    cp_async16(Ws + n * WLD + 2 * k2, wsrc + n * JP + 2 * k2);
  }
  load_slab(Xs, P, ld, cI, cJ, tid);
  cp_async_wait_all();
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < JP / 4; ++kk) {
    const int k = kk * 4 + t;
    double af[2], bf[8];
#pragma unroll
    for (int i = 0; i < 2; ++i) af[i] = Xs[k * XLD + 16 * warp + 8 * i + g];
#pragma unroll
    for (int j = 0; j < 8; ++j) bf[j] = Ws[(8 * j + g) * WLD + k];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = 16 * warp + 8 * i + g, n = 8 * j + 2 * t + e;
        const long long gcol = n < JB ? cI + n : cJ + n - JB;
        P[gcol * ld + m] = acc[i][j][e];
      }
}

// ---- finalisation ------------------------------------------------------------------------------
__global__ void k_colnorm2(const double* __restrict__ G, long long ldg, long long mrows, double* __restrict__ s2) {
  const long long j = blockIdx.x;
  double v = 0.0;
  for (long long i = threadIdx.x; i < mrows; i += blockDim.x) {
    const double x = G[i + j * ldg];
    v += x * x;
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    s2[j] = t;
  }
}

// U[:, r] = G[:, perm[r]] / s ; V[:, r] = Vw[:, perm[r]] ; null columns: U <- Vw column (sym) or 0
__global__ void k_finalize(const double* __restrict__ G, long long ldg, const double* __restrict__ Vw, long long ldv,
                           const int* __restrict__ perm, const double* __restrict__ sval, const int* __restrict__ isnull,
                           int sym, long long m, long long n, double* __restrict__ U, double* __restrict__ Vout) {
  const long long r = blockIdx.x;
  const long long j = perm[r];
  const double s = sval[r];
  const int nul = isnull[r];
  const double inv = nul ? 0.0 : 1.0 / s;
  for (long long i = threadIdx.x; i < m; i += blockDim.x) {
    double u;
    if (!nul) u = G[i + j * ldg] * inv;
    else u = (sym && i < n) ? Vw[i + j * ldv] : 0.0;
    U[i + r * m] = u;
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) Vout[i + r * n] = Vw[i + j * ldv];
}

// G[i,j] = A(i,j) (+ A(j,i) if sym) for i < m, j < n, where A is a rank-4 strided view read as the
// matrix [(i0,i1),(j0,j1)] (column-major merges); this folds trg.jl:20-21's permutes and
// ctmrg.jl:135's `cpmat += cpmat'` into the load of the SVD working matrix.  G is pre-zeroed.
__global__ void k_load_matrix(double* G, long long ldg, const double* __restrict__ A, long long d0, long long d2,
                              long long s0, long long s1, long long s2, long long s3, long long m, long long n,
                              int sym) {
  const long long total = m * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % m, j = idx / m;
    const long long i0 = i % d0, i1 = i / d0, j0 = j % d2, j1 = j / d2;
    double v = A[i0 * s0 + i1 * s1 + j0 * s2 + j1 * s3];
    if (sym) {
      const long long a0 = j % d0, a1 = j / d0, b0 = i % d2, b1 = i / d2;
      v += A[a0 * s0 + a1 * s1 + b0 * s2 + b1 * s3];
    }
    G[i + j * ldg] = v;
  }
}

__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

// E = 1.5 I - 0.5 T in place (Newton-Schulz polish of a warm-start basis)
__global__ void k_ns_factor(double* T, long long n) {
  const long long total = n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % n, j = idx / n;
    T[idx] = (i == j ? 1.5 : 0.0) - 0.5 * T[idx];
  }
}

__global__ void k_fill_random(double* p, long long n, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);   // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    p[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}


}  // namespace

static SvdResult svd_jacobi_impl(tnad_ctx* c, const Tens& Ain0, bool sym, int depth, const Tens* V0,
                                 bool complete_null) {
  TNAD_REQUIRE(Ain0.rank == 2 || Ain0.rank == 4, "svd: need a matrix or a rank-4 view [(i0,i1),(j0,j1)]");
  Tens Ain = Ain0;
  if (Ain0.rank == 2) {   // promote to rank 4 with unit middle dims
    Ain.rank = 4;
    Ain.dim[0] = Ain0.dim[0]; Ain.str[0] = Ain0.str[0];
    Ain.dim[1] = 1;           Ain.str[1] = 0;
    Ain.dim[2] = Ain0.dim[1]; Ain.str[2] = Ain0.str[1];
    Ain.dim[3] = 1;           Ain.str[3] = 0;
  }
  const int64_t m0 = Ain.dim[0] * Ain.dim[1], n0 = Ain.dim[2] * Ain.dim[3];
  if (sym) TNAD_REQUIRE(m0 == n0 && Ain.dim[0] == Ain.dim[2], "svd: symmetrised input must be square");
  const bool transposed = m0 < n0;
  Tens A = transposed ? t_perm(Ain, {2, 3, 0, 1}) : Ain;
  const int64_t m = A.dim[0] * A.dim[1], n = A.dim[2] * A.dim[3];
  TNAD_REQUIRE(m >= 1 && n >= 1, "svd: empty matrix");
  const int64_t mpad = (m + RC - 1) / RC * RC;
  const int64_t N = (n + JP - 1) / JP * JP;
  const int64_t ldv = (N + RC - 1) / RC * RC;
  const int p = (int)(N / JB);
  const int npairs = p / 2;
  const int mchunks = (int)(mpad / RC), vchunks = (int)(ldv / RC);

  static std::atomic<unsigned long long> attr_devs{0};   // kernel attributes are per device: one bit per device id (set after the attribute call: a racing thread at worst repeats it)
  const bool attr_set = (attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL;
  const size_t smem_gram = (size_t)JP * XLD * sizeof(double);
  const size_t smem_upd = (size_t)(JP * XLD + JP * WLD) * sizeof(double);
  if (!attr_set) {
    TNAD_CUDA(cudaFuncSetAttribute(k_jacobi_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gram));
    TNAD_CUDA(cudaFuncSetAttribute(k_jacobi_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }

  Tens G = t_alloc(c, {mpad, N}, true);
  Tens Vw = t_alloc(c, {ldv, N}, false);
  {
    const long long total = m * n;
    int nb = (int)std::min<long long>((total + 1023) / 1024, 148 * 8);
    if (nb < 1) nb = 1;
    k_load_matrix<<<nb, 256, 0, c->stream>>>(G.p, mpad, A.p, A.dim[0], A.dim[2], A.str[0], A.str[1], A.str[2],
                                             A.str[3], m, n, sym ? 1 : 0);
    LAUNCH_CHECK(c);
  }
  set_identity(c, Vw.p, ldv, N);
  if (V0 && !transposed && m == n && V0->rank == 2 && V0->dim[0] == n && V0->dim[1] == n) {
    // Warm start (CTMRG: successive cpmat are close): V <- polish(V0), G <- A V.  Any orthogonal V
    // is a valid start for one-sided Jacobi; a good one leaves only a few sweeps of work.
    Tens Ms = t_alloc(c, {m, n});
    {
      Tens Gv = t_wrap(G.p, {m, n});
      Gv.str[1] = mpad;
      tcopy(c, Gv, Ms, 1.0, 0.0);
    }
    Tens T = contract_new(c, "ki,kj->ij", *V0, *V0);
    {
      const long long total = n * n;
      int nb = (int)std::min<long long>((total + 1023) / 1024, 148 * 8);
      k_ns_factor<<<nb < 1 ? 1 : nb, 256, 0, c->stream>>>(T.p, n);
      LAUNCH_CHECK(c);
    }
    Tens Vv = t_wrap(Vw.p, {n, n});
    Vv.str[1] = ldv;
    contract(c, "ik,kj->ij", *V0, T, Vv, 1.0, 0.0);
    Tens Gv = t_wrap(G.p, {m, n});
    Gv.str[1] = mpad;
    contract(c, "ik,kj->ij", Ms, Vv, Gv, 1.0, 0.0);
  }

  int nsplit = (2 * c->num_sms + npairs - 1) / npairs;
  if (nsplit > mchunks) nsplit = mchunks;
  if (nsplit < 1) nsplit = 1;
  Tens Hpart = t_alloc(c, {(int64_t)JP * JP, (int64_t)npairs * nsplit});
  Tens Wbuf = t_alloc(c, {(int64_t)JP * JP, (int64_t)npairs});
  Tens skipbuf = t_alloc(c, {(int64_t)npairs + 2});
  int* skip = reinterpret_cast<int*>(skipbuf.p);
  unsigned long long* offbits = reinterpret_cast<unsigned long long*>(c->scal + 16);

  const double eps = 2.220446049250313e-16;
  const double tol = std::max(8.0, 2.0 * std::sqrt((double)m)) * eps;
  // numerically-null threshold on squared column norms: (4 eps)^2 max(m,n) |A|_F^2
  double nullfac = 16.0 * eps * eps * (double)std::max(m, n);
  if (const char* nr = opt_s(c, "TNAD_NULL_REL")) {   // experiment: freeze columns below nr * |A|_F
    const double v = atof(nr);
    if (v > 0.0) nullfac = v * v;
  }
  double* fro2 = c->scal + 18;
  reduce(c, RED_SUMSQ, G, nullptr, fro2);
  const bool debug = opt_i(c, "TNAD_JACOBI_DEBUG", 0) != 0;
  const bool cross_mode = opt_i(c, "TNAD_JACOBI_CROSS", 0) != 0;
  const int max_inner = opt_i(c, "TNAD_JACOBI_INNER", 1);
  const int max_sweeps = opt_i(c, "TNAD_JACOBI_SWEEPS", 60);

  SvdResult res;
  double prev_off = 1e300;
  bool converged = false;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    k_set_u64<<<1, 1, 0, c->stream>>>(offbits, 0ULL);
    LAUNCH_CHECK(c);
    for (int r = 0; r < p - 1; ++r) {
      {
      KTimer kt(c, KF_GRAM);
      k_jacobi_gram<<<dim3(npairs, nsplit), 256, smem_gram, c->stream>>>(G.p, mpad, mchunks, p, r, nsplit, Hpart.p);
      }
      LAUNCH_CHECK(c);
      {
      KTimer kt(c, KF_EIG);
      // register-resident pivot kernel (symeig.cu, Gram mode)
        launch_gram_pivot_eig(c, Hpart.p, nsplit, npairs, p, r, (int)n, fro2, tol, nullfac, max_inner, (cross_mode && r > 0) ? 1 : 0,
                              Wbuf.p, skip, offbits, nullptr);
      }
      LAUNCH_CHECK(c);
      {
      KTimer kt(c, KF_UPDATE);
      k_jacobi_update<<<dim3(npairs, mchunks + vchunks), 256, smem_upd, c->stream>>>(G.p, mpad, mchunks, Vw.p, ldv, p,
                                                                                     r, Wbuf.p, skip, c->ktiming ? reinterpret_cast<unsigned long long*>(c->scal + 21) : nullptr);
      }
      LAUNCH_CHECK(c);
    }
    double off;
    d2h(c, &off, c->scal + 16, 1);
    if (debug) fprintf(stderr, "[tnad jacobi] m=%lld n=%lld sweep %d off %.3e (tol %.1e)\n", (long long)m, (long long)n, sweep, off, tol);
    if (off <= 4.0 * tol) {   // rotations are applied down to tol; 4 tol is the rounding floor of the cosines
      converged = true;
      ++sweep;
      break;
    }
    if (sweep >= 8 && off < 1e-11 && off > 0.5 * prev_off) {   // rounding floor reached
      converged = true;
      ++sweep;
      break;
    }
    prev_off = off;
  }
  if (!converged) fail(TNAD_ERR_NOCONV, "svd: block Jacobi did not converge in " + std::to_string(max_sweeps) + " sweeps");
  res.sweeps = sweep;

  // singular values, ordering, null columns
  Tens s2 = t_alloc(c, {N});
  k_colnorm2<<<(int)N, 128, 0, c->stream>>>(G.p, mpad, mpad, s2.p);
  LAUNCH_CHECK(c);
  std::vector<double> s2h((size_t)N);
  d2h(c, s2h.data(), s2.p, (size_t)N);
  double fro2h;
  d2h(c, &fro2h, fro2, 1);
  const double thr2 = nullfac * fro2h;
  std::vector<int> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return s2h[a] > s2h[b]; });
  std::vector<double> sval((size_t)n);
  std::vector<int> isnull((size_t)n);
  std::vector<int64_t> nullcols;
  for (int64_t r = 0; r < n; ++r) {
    const double v = s2h[perm[r]];
    sval[r] = std::sqrt(v);
    isnull[r] = (v <= thr2) ? 1 : 0;
    if (isnull[r]) nullcols.push_back(r);
  }
  // upload perm / sval / isnull
  Tens meta = t_alloc(c, {3 * n + 4});
  int* dperm = reinterpret_cast<int*>(meta.p);
  int* dnull = reinterpret_cast<int*>(meta.p + (n + 1) / 2 + 1);
  double* dsval = meta.p + 2 * ((n + 1) / 2 + 1);
  TNAD_CUDA(cudaMemcpyAsync(dperm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  TNAD_CUDA(cudaMemcpyAsync(dnull, isnull.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  TNAD_CUDA(cudaMemcpyAsync(dsval, sval.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));

  Tens U = t_alloc(c, {m, n});
  Tens Vo = t_alloc(c, {n, n});
  Tens S = t_alloc(c, {n});
  k_finalize<<<(int)n, 128, 0, c->stream>>>(G.p, mpad, Vw.p, ldv, dperm, dsval, dnull, sym ? 1 : 0, m, n, U.p, Vo.p);
  LAUNCH_CHECK(c);
  TNAD_CUDA(cudaMemcpyAsync(S.p, dsval, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  sync(c);   // perm/sval/isnull host vectors must outlive the async uploads

  if (!sym && !nullcols.empty() && !complete_null && !transposed) {
    // the caller (TRG) handles the left null space through the projector I - U_r U_r' in svd_back
    res.rank_left = n - (int64_t)nullcols.size();
  } else if (!sym && !nullcols.empty()) {
    // Orthonormal completion of the left null space (needed because svd_back uses the full U):
    // project a random block out of span(U_range) twice, then orthonormalise it with the same
    // Jacobi kernels (left singular vectors of a full-column-rank block).
    const int64_t q = (int64_t)nullcols.size(), r = n - q;   // nulls sort last
    TNAD_REQUIRE(depth < 3, "svd: null-space completion did not terminate");
    Tens Wr = t_alloc(c, {m, q});
    {
      const long long total = m * q;
      int nb = (int)std::min<long long>((total + 1023) / 1024, 148 * 8);
      k_fill_random<<<nb < 1 ? 1 : nb, 256, 0, c->stream>>>(Wr.p, total, 0x9E3779B97F4A7C15ULL + (unsigned long long)depth);
      LAUNCH_CHECK(c);
    }
    if (r > 0) {
      Tens Ur = t_slice_last(U, 0, r);
      for (int pass = 0; pass < 2; ++pass) {
        Tens Cm = contract_new(c, "mr,mq->rq", Ur, Wr);
        contract(c, "mr,rq->mq", Ur, Cm, Wr, -1.0, 1.0);
      }
    }
    SvdResult wn = svd_jacobi_impl(c, Wr, false, depth + 1, nullptr, true);
    TNAD_CUDA(cudaMemcpyAsync(U.p + r * m, wn.U.p, (size_t)(m * q) * sizeof(double), cudaMemcpyDeviceToDevice,
                              c->stream));
  }

  res.s_host = sval;
  res.null_thr = std::sqrt(thr2);
  if (transposed) {
    res.U = Vo;
    res.V = U;
  } else {
    res.U = U;
    res.V = Vo;
  }
  res.S = S;
  return res;
}

// X[:, (I,J)] <- X[:, (I,J)] * W_pair for every pair of the round (X has ld rows in nchunks 128-row chunks)
void jacobi_rotate_columns(tnad_ctx* c, double* X, int64_t ld, int nchunks, int p, int round, const double* Wbuf,
                           const int* skip, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_devs{0};   // kernel attributes are per device: one bit per device id (set after the attribute call: a racing thread at worst repeats it)
  const bool attr_set = (attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL;
  const size_t smem_upd = (size_t)(JP * XLD + JP * WLD) * sizeof(double);
  if (!attr_set) {
    TNAD_CUDA(cudaFuncSetAttribute(k_jacobi_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  k_jacobi_update<<<dim3(p / 2, nchunks), 256, smem_upd, st ? st : c->stream>>>(X, ld, nchunks, nullptr, 0, p, round, Wbuf, skip,
      c->ktiming ? reinterpret_cast<unsigned long long*>(c->scal + 21) : nullptr);
  LAUNCH_CHECK(c);
}

SvdResult svd_jacobi(tnad_ctx* c, const Tens& A, bool sym_add_transpose, const Tens* V0, bool complete_null) {
  SvdResult r = svd_jacobi_impl(c, A, sym_add_transpose, 0, V0, complete_null);
  const int64_t m = r.U.dim[0], n = r.V.dim[0], k = r.S.dim[0];
  signfix_cols(c, r.U.p, m, m, r.V.p, n, n, k);   // canonical gauge
  return r;
}

}  // namespace tnad
