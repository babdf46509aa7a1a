// HBM-bound elementwise / reduction kernels of the TRG / CTMRG / energy path.
// Each kernel cites the reference lines whose arithmetic it carries.
#include "common.h"

namespace tnad {

namespace {

constexpr int TB = 256;

inline int grid_for(int64_t n, int per_thread = 1, int cap = 148 * 8) {
  int64_t b = (n + (int64_t)TB * per_thread - 1) / ((int64_t)TB * per_thread);
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

#define LAUNCH_CHECK(c)        \
  do {                         \
    (c)->launches++;           \
    TNAD_CUDA(cudaGetLastError()); \
  } while (0)

// ------------------------------------------------------------------------------------------
// strided copy / accumulate:  out = alpha * in + beta * out   (all `permutedims` of the reference
// that cannot be folded into a GEMM operand, symmetrisation adds ctmrg.jl:145-146, ipeps.jl:34-37)
// ------------------------------------------------------------------------------------------
struct CopyDesc {
  int rank;
  long long total;
  int dim[MAXR];
  long long sin[MAXR], sout[MAXR];
};

__global__ void k_tcopy(const double* __restrict__ in, double* __restrict__ out, const __grid_constant__ CopyDesc d,
                        double alpha, double beta) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx, oi = 0, oo = 0;
#pragma unroll 1
    for (int k = 0; k < d.rank; ++k) {
      long long q = r / d.dim[k];
      long long i = r - q * d.dim[k];
      oi += i * d.sin[k];
      oo += i * d.sout[k];
      r = q;
    }
    double v = alpha * in[oi];
    if (beta != 0.0) v += beta * out[oo];
    out[oo] = v;
  }
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// ------------------------------------------------------------------------------------------
// deterministic two-stage reductions (norm: ctmrg.jl:149-150, ipeps.jl:38, variationalipeps.jl:51;
// maximum(abs.(a)): trg.jl:16; dots: variationalipeps.jl:53-54)
// ------------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ double red_elem(double x, double y) {
  if (OP == RED_SUMSQ) return x * x;
  if (OP == RED_DOT) return x * y;
  if (OP == RED_ABSMAX) return fabs(x);
  return x;
}
template <int OP>
__device__ __forceinline__ double red_comb(double a, double b) {
  if (OP == RED_ABSMAX) return fmax(a, b);
  return a + b;
}

template <int OP>
__device__ double block_reduce(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;   // valid in thread 0
}

template <int OP>
__global__ void k_reduce1(const double* __restrict__ x, const double* __restrict__ y, long long n, double* partial) {
  double v = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v = red_comb<OP>(v, red_elem<OP>(x[i], OP == RED_DOT ? y[i] : 0.0));
  v = block_reduce<OP>(v);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
template <int OP>
__global__ void k_reduce2(const double* __restrict__ partial, int np, double* res) {
  double v = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) v = red_comb<OP>(v, partial[i]);
  v = block_reduce<OP>(v);
  if (threadIdx.x == 0) *res = v;
}

__global__ void k_scale_dev(const double* __restrict__ in, double* __restrict__ out, long long n,
                            const double* __restrict__ scalar, int mode) {
  const double s = *scalar;
  const double f = mode == SC_INV ? 1.0 / s : (mode == SC_INVSQRT ? 1.0 / sqrt(s) : s);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i] * f;
}

// norm pullback of x/||x|| (autodiff.jl:23-29 with eps(0f0) ~ 1.4e-45 added to the norm)
__global__ void k_norm_back(const double* __restrict__ ybar, const double* __restrict__ x, long long n,
                            const double* __restrict__ ss, const double* __restrict__ dot, double* __restrict__ xbar) {
  const double nrm = sqrt(*ss);
  const double a = 1.0 / nrm;
  const double b = *dot / (nrm * nrm * (nrm + 1.401298464324817e-45));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    xbar[i] = ybar[i] * a - b * x[i];
}

// trace = sum_ij a[i,j,i,j]  (trg.jl:27)
__global__ void k_trace_ijij(const double* __restrict__ a, int d0, int d1, long long s0, long long s1, long long s2,
                             long long s3, double* res) {
  double v = 0.0;
  for (int q = threadIdx.x; q < d0 * d1; q += blockDim.x) {
    int i = q % d0, j = q / d0;
    v += a[i * (s0 + s2) + j * (s1 + s3)];
  }
  v = block_reduce<RED_SUM>(v);
  if (threadIdx.x == 0) *res = v;
}

__global__ void k_add_diag_ijij(double* a, int d0, int d1, long long s0, long long s1, long long s2, long long s3,
                                double w) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < d0 * d1; q += gridDim.x * blockDim.x) {
    int i = q % d0, j = q / d0;
    a[i * (s0 + s2) + j * (s1 + s3)] += w;
  }
}

// u[:, j] = U[:, j] * sqrt(S[j])   (trg.jl:38-41)
__global__ void k_colscale_sqrt(const double* __restrict__ in, long long ldin, const double* __restrict__ S,
                                double* __restrict__ out, long long ldout, long long m, long long k) {
  const long long total = m * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long j = idx / m, i = idx - j * m;
    out[i + j * ldout] = in[i + j * ldin] * sqrt(S[j]);
  }
}

__global__ void k_colscale_sinv(double* __restrict__ x, long long ld, long long m, long long k,
                                const double* __restrict__ S, double eta) {
  const long long total = m * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long j = idx / m, i = idx - j * m;
    const double s = S[j];
    x[i + j * ld] *= s / (s * s + eta);
  }
}

__global__ void k_set_identity(double* p, long long ld, long long n) {
  const long long total = ld * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long j = idx / ld, i = idx - j * ld;
    p[idx] = (i == j) ? 1.0 : 0.0;
  }
}

// _initializect_square(bulk, Val(:raw), chi)  (ctmrg.jl:74-86); corner/edge are pre-zeroed
__global__ void k_init_raw(const double* __restrict__ bulk, int D, int chi, double* corner, double* edge) {
  const int m = D < chi ? D : chi;
  const long long D2 = (long long)D * D, D3 = D2 * D;
  const int total = m * D * m;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int i = idx % m, j = (idx / m) % D, k = idx / (m * D);
    double v = 0.0;
    for (int l = 0; l < D; ++l) v += bulk[i + (long long)j * D + k * D2 + l * D3];
    edge[i + (long long)j * chi + (long long)k * chi * D] = v;
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m * m; idx += gridDim.x * blockDim.x) {
    int i = idx % m, j = idx / m;
    double v = 0.0;
    for (int k = 0; k < D; ++k)
      for (int l = 0; l < D; ++l) v += bulk[i + (long long)j * D + k * D2 + l * D3];
    corner[i + (long long)j * chi] = v;
  }
}

// ap[(a,i),(b,j),(c,k),(d,l),x,y] = A[a,b,c,d,x] * A[i,j,k,l,y]   (variationalipeps.jl:32-33)
__global__ void k_double_layer(const double* __restrict__ A, int d, int s, double* __restrict__ ap) {
  const long long D = (long long)d * d, D4 = D * D * D * D, total = D4 * s * s;
  const long long d4 = (long long)d * d * d * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    int m0 = r % D; r /= D;
    int m1 = r % D; r /= D;
    int m2 = r % D; r /= D;
    int m3 = r % D; r /= D;
    int x = r % s, y = r / s;
    int a = m0 % d, i = m0 / d, b = m1 % d, j = m1 / d, c = m2 % d, k = m2 / d, dd = m3 % d, l = m3 / d;
    const double ket = A[a + d * (b + d * (c + d * dd)) + d4 * x];
    const double bra = A[i + d * (j + d * (k + d * l)) + d4 * y];
    ap[idx] = ket * bra;
  }
}

// a[q] = sum_x ap[q, x, x]   (variationalipeps.jl:34)
__global__ void k_ptrace(const double* __restrict__ ap, long long D4, int s, double* __restrict__ a) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < D4; q += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int x = 0; x < s; ++x) v += ap[q + D4 * (x + (long long)s * x)];
    a[q] = v;
  }
}

// adjoint of the double layer + partial trace: one block per element of Abar
__global__ void k_double_layer_back(const double* __restrict__ A, int d, int s, const double* __restrict__ apbar,
                                    const double* __restrict__ abar, double* __restrict__ Abar) {
  const long long D = (long long)d * d, D4 = D * D * D * D;
  const int d4 = d * d * d * d;
  const int o = blockIdx.x;               // (a,b,c,dd,x)
  int r = o;
  const int a = r % d; r /= d;
  const int b = r % d; r /= d;
  const int c = r % d; r /= d;
  const int dd = r % d; r /= d;
  const int x = r;
  double v = 0.0;
  for (int q = threadIdx.x; q < d4 * s; q += blockDim.x) {
    int rr = q;
    const int i = rr % d; rr /= d;
    const int j = rr % d; rr /= d;
    const int k = rr % d; rr /= d;
    const int l = rr % d; rr /= d;
    const int y = rr;
    const double Aq = A[q];
    // ket slot: apbar'[(a,i),(b,j),(c,k),(dd,l),x,y]
    const long long q1 = (a + d * i) + D * ((b + d * j) + D * ((c + d * k) + D * (long long)(dd + d * l)));
    double g1 = apbar[q1 + D4 * (x + (long long)s * y)];
    if (x == y) g1 += abar[q1];
    // bra slot: apbar'[(i,a),(j,b),(k,c),(l,dd),y,x]
    const long long q2 = (i + d * a) + D * ((j + d * b) + D * ((k + d * c) + D * (long long)(l + d * dd)));
    double g2 = apbar[q2 + D4 * (y + (long long)s * x)];
    if (x == y) g2 += abar[q2];
    v += (g1 + g2) * Aq;
  }
  v = block_reduce<RED_SUM>(v);
  if (threadIdx.x == 0) Abar[o] = v;
}

// svd_back core (trg.jl:76-93): panels of  R = (J+J')*S + S*(K+K') + diag(dS),
// J = F.*(U'dU), K = F.*(V'dV), F[i,j] = (S_j^2 - S_i^2) / ((S_j^2 - S_i^2)^2 + eta),
// for cotangents that are non-zero only in their first k columns.
__global__ void k_svdback_panels(long long n, long long k, const double* __restrict__ S,
                                 const double* __restrict__ G1, const double* __restrict__ G2,
                                 const double* __restrict__ dS, double eta, double* __restrict__ Rrow,
                                 double* __restrict__ Rcol) {
  const long long total = 2 * n * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    if (idx < n * k) {
      // Rrow[i,j], i < k, j < n  (ld = k)
      const long long i = idx % k, j = idx / k;
      const double si = S[i], sj = S[j];
      const double dlt = sj * sj - si * si;
      const double F = dlt / (dlt * dlt + eta);
      double acc = 0.0;
      if (G1) {
        const double gij = (j < k) ? G1[i + j * n] : 0.0;
        const double gji = G1[j + i * n];
        acc += (gij - gji) * sj;
      }
      if (G2) {
        const double gij = (j < k) ? G2[i + j * n] : 0.0;
        const double gji = G2[j + i * n];
        acc += si * (gij - gji);
      }
      double r = F * acc;
      if (dS && i == j) r += dS[i];
      Rrow[idx] = r;
    } else {
      // Rcol[i,j], i < n, j < k (ld = n); rows i < k are covered by Rrow -> zero here
      const long long q = idx - n * k;
      const long long i = q % n, j = q / n;
      double r = 0.0;
      if (i >= k) {
        const double si = S[i], sj = S[j];
        const double dlt = sj * sj - si * si;
        const double F = dlt / (dlt * dlt + eta);
        double acc = 0.0;
        if (G1) acc += G1[i + j * n] * sj;
        if (G2) acc += si * G2[i + j * n];
        r = F * acc;
      }
      Rcol[q] = r;
    }
  }
}

// reverse of trg.jl:38-41 for one split; one block per kept column j
__global__ void k_trg_factor_back(long long m, long long n, long long k, const double* __restrict__ U, long long ldu,
                                  const double* __restrict__ V, long long ldv, const double* __restrict__ S,
                                  const double* __restrict__ du, const double* __restrict__ dvt,
                                  double* __restrict__ dUk, double* __restrict__ dVk, double* __restrict__ dS) {
  const long long j = blockIdx.x;
  const double sq = sqrt(S[j]);
  double acc = 0.0;
  for (long long i = threadIdx.x; i < m; i += blockDim.x) {
    const double g = du[i + j * m];
    acc += U[i + j * ldu] * g;
    dUk[i + j * m] = g * sq;
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double g = dvt[j + i * k];
    acc += V[i + j * ldv] * g;
    dVk[i + j * n] = g * sq;
  }
  acc = block_reduce<RED_SUM>(acc);
  if (threadIdx.x == 0) dS[j] = sq > 0.0 ? acc / (2.0 * sq) : 0.0;
}

// first index with |a[i]| == maxval (findmax semantics of Zygote's `maximum` adjoint)
__global__ void k_argmax_first(const double* __restrict__ a, long long n, double maxval, unsigned long long* idx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (fabs(a[i]) == maxval) atomicMin(idx, (unsigned long long)i);
}
__global__ void k_maxval_apply(double* da_in, const double* __restrict__ a_in, const unsigned long long* idx,
                               const double* __restrict__ dot, double maxval, double coef) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double dmax = -(*dot) / (maxval * maxval) + coef / maxval;
    const unsigned long long i = *idx;
    if (i != ~0ULL) da_in[i] += dmax * (a_in[i] >= 0.0 ? 1.0 : -1.0);
  }
}
__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

}  // namespace

// ------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------
void tcopy(tnad_ctx* c, const Tens& in, Tens& out, double alpha, double beta) {
  TNAD_REQUIRE(in.rank == out.rank, "tcopy: rank mismatch");
  CopyDesc d;
  memset(&d, 0, sizeof(d));
  long long total = 1;
  // order the loop by the output strides (fastest first), drop unit dims, merge compatible dims
  std::vector<int> order;
  for (int i = 0; i < out.rank; ++i) {
    TNAD_REQUIRE(in.dim[i] == out.dim[i], "tcopy: dim mismatch");
    total *= out.dim[i];
    if (out.dim[i] != 1) order.push_back(i);
  }
  if (total == 0) return;
  for (size_t i = 1; i < order.size(); ++i)
    for (size_t j = i; j > 0 && out.str[order[j]] < out.str[order[j - 1]]; --j) std::swap(order[j], order[j - 1]);
  int r = 0;
  for (int i : order) {
    if (r > 0 && out.str[i] == d.sout[r - 1] * d.dim[r - 1] && in.str[i] == d.sin[r - 1] * d.dim[r - 1] &&
        (long long)d.dim[r - 1] * out.dim[i] < (1LL << 31)) {
      d.dim[r - 1] *= (int)out.dim[i];
      continue;
    }
    d.dim[r] = (int)out.dim[i];
    d.sin[r] = in.str[i];
    d.sout[r] = out.str[i];
    ++r;
  }
  d.rank = r;
  d.total = total;
  k_tcopy<<<grid_for(total, 4), TB, 0, c->stream>>>(in.p, out.p, d, alpha, beta);
  LAUNCH_CHECK(c);
}

__global__ void k_signfix_cols(double* U, long long ldu, long long m, double* V, long long ldv, long long nv) {
  __shared__ double sh[SIGNFIX_SH];
  double* u = U + (long long)blockIdx.x * ldu;
  const double s = block_canonical_sign(u, m, sh);
  if (s < 0.0) {
    for (long long i = threadIdx.x; i < m; i += blockDim.x) u[i] = -u[i];
    if (V) {
      double* v = V + (long long)blockIdx.x * ldv;
      for (long long i = threadIdx.x; i < nv; i += blockDim.x) v[i] = -v[i];
    }
  }
}

void signfix_cols(tnad_ctx* c, double* U, int64_t ldu, int64_t m, double* V, int64_t ldv, int64_t nv, int64_t ncols) {
  if (ncols <= 0 || m <= 0) return;
  k_signfix_cols<<<(int)ncols, 128, 0, c->stream>>>(U, ldu, m, V, ldv, nv);
  LAUNCH_CHECK(c);
}

void fill(tnad_ctx* c, double* p, int64_t n, double v) {
  if (n <= 0) return;
  k_fill<<<grid_for(n, 4), TB, 0, c->stream>>>(p, n, v);
  LAUNCH_CHECK(c);
}

template <int OP>
static void reduce_t(tnad_ctx* c, const double* x, const double* y, long long n, double* res) {
  int nb = grid_for(n, 8, 1024);
  k_reduce1<OP><<<nb, TB, 0, c->stream>>>(x, y, n, c->partial);
  LAUNCH_CHECK(c);
  k_reduce2<OP><<<1, TB, 0, c->stream>>>(c->partial, nb, res);
  LAUNCH_CHECK(c);
}

void reduce(tnad_ctx* c, RedOp op, const Tens& x, const Tens* y, double* res) {
  TNAD_REQUIRE(x.contiguous(), "reduce: x must be contiguous");
  if (op == RED_DOT) TNAD_REQUIRE(y && y->contiguous() && y->numel() == x.numel(), "reduce: bad y");
  const long long n = x.numel();
  switch (op) {
    case RED_SUMSQ: reduce_t<RED_SUMSQ>(c, x.p, nullptr, n, res); break;
    case RED_DOT: reduce_t<RED_DOT>(c, x.p, y->p, n, res); break;
    case RED_ABSMAX: reduce_t<RED_ABSMAX>(c, x.p, nullptr, n, res); break;
    default: reduce_t<RED_SUM>(c, x.p, nullptr, n, res); break;
  }
}

void scale_dev(tnad_ctx* c, const Tens& in, Tens& out, const double* scalar, ScaleMode mode) {
  TNAD_REQUIRE(in.contiguous() && out.contiguous() && in.numel() == out.numel(), "scale_dev: bad tensors");
  const long long n = in.numel();
  if (!n) return;
  k_scale_dev<<<grid_for(n, 4), TB, 0, c->stream>>>(in.p, out.p, n, scalar, (int)mode);
  LAUNCH_CHECK(c);
}

void norm_back(tnad_ctx* c, const Tens& ybar, const Tens& x, const double* ss, const double* dot, Tens& xbar) {
  TNAD_REQUIRE(ybar.contiguous() && x.contiguous() && xbar.contiguous(), "norm_back: views not supported");
  const long long n = x.numel();
  k_norm_back<<<grid_for(n, 4), TB, 0, c->stream>>>(ybar.p, x.p, n, ss, dot, xbar.p);
  LAUNCH_CHECK(c);
}

void trace_ijij(tnad_ctx* c, const Tens& a, double* res) {
  TNAD_REQUIRE(a.rank == 4 && a.dim[0] == a.dim[2] && a.dim[1] == a.dim[3], "trace_ijij: need (i,j,i,j)");
  k_trace_ijij<<<1, TB, 0, c->stream>>>(a.p, (int)a.dim[0], (int)a.dim[1], a.str[0], a.str[1], a.str[2], a.str[3], res);
  LAUNCH_CHECK(c);
}

void add_diag_trace_back(tnad_ctx* c, Tens& abar, double w) {
  k_add_diag_ijij<<<grid_for(abar.dim[0] * abar.dim[1]), TB, 0, c->stream>>>(
      abar.p, (int)abar.dim[0], (int)abar.dim[1], abar.str[0], abar.str[1], abar.str[2], abar.str[3], w);
  LAUNCH_CHECK(c);
}

void colscale_sqrt(tnad_ctx* c, const double* in, int64_t ldin, const double* S, double* out, int64_t ldout,
                   int64_t m, int64_t k) {
  if (m * k == 0) return;
  k_colscale_sqrt<<<grid_for(m * k, 2), TB, 0, c->stream>>>(in, ldin, S, out, ldout, m, k);
  LAUNCH_CHECK(c);
}

void colscale_sinv(tnad_ctx* c, double* x, int64_t ld, int64_t m, int64_t k, const double* S, double eta) {
  if (m * k == 0) return;
  k_colscale_sinv<<<grid_for(m * k, 2), TB, 0, c->stream>>>(x, ld, m, k, S, eta);
  LAUNCH_CHECK(c);
}

void set_identity(tnad_ctx* c, double* p, int64_t ld, int64_t n) {
  k_set_identity<<<grid_for(ld * n, 4), TB, 0, c->stream>>>(p, ld, n);
  LAUNCH_CHECK(c);
}

void init_raw(tnad_ctx* c, const Tens& bulk, Tens& corner, Tens& edge) {
  const int D = (int)bulk.dim[0], chi = (int)corner.dim[0];
  TNAD_REQUIRE(bulk.contiguous() && corner.contiguous() && edge.contiguous(), "init_raw: views not supported");
  t_zero(c, corner);
  t_zero(c, edge);
  k_init_raw<<<grid_for((int64_t)D * D * D), TB, 0, c->stream>>>(bulk.p, D, chi, corner.p, edge.p);
  LAUNCH_CHECK(c);
}

// _initializect_square(bulk, Val(:random), chi) (ctmrg.jl:66-72) on the device: standard normals from a counter-based
// generator (splitmix64 of (seed, element index) -> two uniforms -> Box-Muller), then corner += corner', edge +=
// permutedims(edge, (3,2,1)).  The reference draws from Julia's global RNG, so the stream cannot be matched anyway; what
// the caller gets is the same distribution, reproducible from the seed and independent of the launch geometry.
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_randn(double* __restrict__ out, long long n, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long a = splitmix64(seed ^ splitmix64(2ull * (unsigned long long)i));
    const unsigned long long b = splitmix64(seed ^ splitmix64(2ull * (unsigned long long)i + 1ull));
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);      // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);              // [0, 1)
    out[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
}

void init_random(tnad_ctx* c, int64_t D, int64_t chi, unsigned long long seed, Tens& corner, Tens& edge) {
  Tens c0 = t_alloc(c, {chi, chi}), e0 = t_alloc(c, {chi, D, chi});
  k_randn<<<grid_for(c0.numel()), TB, 0, c->stream>>>(c0.p, c0.numel(), splitmix64(seed));
  LAUNCH_CHECK(c);
  k_randn<<<grid_for(e0.numel()), TB, 0, c->stream>>>(e0.p, e0.numel(), splitmix64(seed + 0x5851F42D4C957F2Dull));
  LAUNCH_CHECK(c);
  corner = t_clone(c, c0);
  edge = t_clone(c, e0);
  tcopy(c, t_perm(c0, {1, 0}), corner, 1.0, 1.0);
  tcopy(c, t_perm(e0, {2, 1, 0}), edge, 1.0, 1.0);
}

// indexperm_symmetrize (ipeps.jl:32-39): four permute-adds, then x / norm(x)
static const int SYM_PERMS[4][5] = {{0, 3, 2, 1, 4}, {2, 1, 0, 3, 4}, {1, 0, 3, 2, 4}, {3, 2, 1, 0, 4}};

void ipeps_symmetrize(tnad_ctx* c, const Tens& A, Tens& xsum, Tens& out, double* ss) {
  Tens x = t_clone(c, A);
  for (int p = 0; p < 4; ++p) {
    Tens y = t_clone(c, x);
    Tens xv = t_perm(x, {SYM_PERMS[p][0], SYM_PERMS[p][1], SYM_PERMS[p][2], SYM_PERMS[p][3], SYM_PERMS[p][4]});
    tcopy(c, xv, y, 1.0, 1.0);
    x = y;
  }
  xsum = x;
  reduce(c, RED_SUMSQ, xsum, nullptr, ss);
  out = t_alloc_v(c, std::vector<int64_t>(A.dim, A.dim + A.rank));
  scale_dev(c, xsum, out, ss, SC_INVSQRT);
}

void ipeps_symmetrize_back(tnad_ctx* c, const Tens& ybar, const Tens& xsum, const double* ss, Tens& Abar) {
  double* dot = c->scal + 8;
  reduce(c, RED_DOT, ybar, &xsum, dot);
  Tens x = t_alloc_v(c, std::vector<int64_t>(xsum.dim, xsum.dim + xsum.rank));
  norm_back(c, ybar, xsum, ss, dot, x);
  for (int p = 3; p >= 0; --p) {
    Tens y = t_clone(c, x);
    Tens xv = t_perm(x, {SYM_PERMS[p][0], SYM_PERMS[p][1], SYM_PERMS[p][2], SYM_PERMS[p][3], SYM_PERMS[p][4]});
    tcopy(c, xv, y, 1.0, 1.0);
    x = y;
  }
  Abar = x;
}

void double_layer(tnad_ctx* c, const Tens& A, Tens& ap, Tens& a) {
  const int d = (int)A.dim[0], s = (int)A.dim[4];
  const int64_t D = (int64_t)d * d, D4 = D * D * D * D;
  TNAD_REQUIRE(A.contiguous(), "double_layer: A must be contiguous");
  ap = t_alloc(c, {D, D, D, D, s, s});
  a = t_alloc(c, {D, D, D, D});
  k_double_layer<<<grid_for(D4 * s * s, 2), TB, 0, c->stream>>>(A.p, d, s, ap.p);
  LAUNCH_CHECK(c);
  k_ptrace<<<grid_for(D4), TB, 0, c->stream>>>(ap.p, D4, s, a.p);
  LAUNCH_CHECK(c);
}

void double_layer_back(tnad_ctx* c, const Tens& A, const Tens& apbar, const Tens& abar, Tens& Abar) {
  const int d = (int)A.dim[0], s = (int)A.dim[4];
  Abar = t_alloc_v(c, std::vector<int64_t>(A.dim, A.dim + A.rank));
  k_double_layer_back<<<(int)A.numel(), 128, 0, c->stream>>>(A.p, d, s, apbar.p, abar.p, Abar.p);
  LAUNCH_CHECK(c);
}

void svdback_panels(tnad_ctx* c, int64_t n, int64_t k, const double* S, const double* G1, const double* G2,
                    const double* dS, double eta, double* Rrow, double* Rcol) {
  k_svdback_panels<<<grid_for(2 * n * k, 2), TB, 0, c->stream>>>(n, k, S, G1, G2, dS, eta, Rrow, Rcol);
  LAUNCH_CHECK(c);
}

void trg_factor_back(tnad_ctx* c, int64_t m, int64_t n, int64_t k, const double* U, int64_t ldu, const double* V,
                     int64_t ldv, const double* S, const double* du, const double* dvt, double* dUk, double* dVk,
                     double* dS) {
  k_trg_factor_back<<<(int)k, TB, 0, c->stream>>>(m, n, k, U, ldu, V, ldv, S, du, dvt, dUk, dVk, dS);
  LAUNCH_CHECK(c);
}

void trg_maxval_back(tnad_ctx* c, const Tens& da, const Tens& a_in, double maxval, double coef, Tens& da_in) {
  // a = a_in / maxval ; lnZ += 2^(1-n) log(maxval) ; maxval = maximum(abs.(a_in))   (trg.jl:16-18)
  double* dot = c->scal + 9;
  unsigned long long* idx = reinterpret_cast<unsigned long long*>(c->scal + 10);
  reduce(c, RED_DOT, da, &a_in, dot);
  Tens src = da;
  tcopy(c, src, da_in, 1.0 / maxval, 0.0);
  k_set_u64<<<1, 1, 0, c->stream>>>(idx, ~0ULL);
  LAUNCH_CHECK(c);
  k_argmax_first<<<grid_for(a_in.numel(), 4), TB, 0, c->stream>>>(a_in.p, a_in.numel(), maxval, idx);
  LAUNCH_CHECK(c);
  k_maxval_apply<<<1, 32, 0, c->stream>>>(da_in.p, a_in.p, idx, dot, maxval, coef);
  LAUNCH_CHECK(c);
}

}  // namespace tnad
