// Internal header of libtnad_b200: context, device tensors, error plumbing, kernel prototypes.
// Nothing here is part of the C ABI (see include/tnad.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <memory>
#include <string>
#include <vector>
#include <map>
#include <atomic>
#include <initializer_list>
#include "../../include/tnad.h"

namespace tnad {

struct Error {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Error{code, msg}; }

#define TNAD_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::tnad::fail(_e == cudaErrorMemoryAllocation ? TNAD_ERR_NOMEM : TNAD_ERR_CUDA,            \
                   std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                       std::to_string(__LINE__) + ")");                                         \
    }                                                                                           \
  } while (0)

#define TNAD_REQUIRE(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) ::tnad::fail(TNAD_ERR_ARG, std::string(msg));   \
  } while (0)

#ifdef __CUDACC__
// sign (+1 / -1) that makes the largest-magnitude entry of x[0:n] (first index on ties) positive; whole CTA, blockDim.x
// a multiple of 32 and <= 1024; sh: 66 doubles of shared memory
__device__ __forceinline__ double block_canonical_sign(const double* __restrict__ x, long long n, double* sh) {
  double best = -1.0, val = 1.0;
  long long bi = n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i], a = fabs(v);
    if (a > best) { best = a; val = v; bi = i; }     // strided scan: the first index of a tie within a thread wins
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_down_sync(0xffffffffu, best, o), ov = __shfl_down_sync(0xffffffffu, val, o);
    const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; val = ov; bi = oi; }
  }
  long long* shi = reinterpret_cast<long long*>(sh + 33);
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = val; shi[threadIdx.x >> 5] = bi; sh[66 + (threadIdx.x >> 5)] = best; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w)
      if (sh[66 + w] > best || (sh[66 + w] == best && shi[w] < bi)) { best = sh[66 + w]; val = sh[w]; bi = shi[w]; }
    sh[32] = val < 0.0 ? -1.0 : 1.0;
  }
  __syncthreads();
  const double s = sh[32];
  __syncthreads();
  return s;
}
constexpr int SIGNFIX_SH = 100;   // doubles of shared memory block_canonical_sign needs
#endif

}  // namespace tnad

struct tnad_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr, stream3 = nullptr;   // look-ahead streams of the symmetric eigensolver
  cudaEvent_t ev_eig = nullptr, ev_rest = nullptr, ev_v = nullptr;
  // cached workspace + captured one-sweep CUDA graph of the symmetric eigensolver, keyed by matrix size
  struct SymEigCache* symcache = nullptr;
  std::string err;
  int pointer_mode = TNAD_POINTER_HOST;
  int64_t launches = 0;
  double* scal = nullptr;      // device scalar scratch (SCAL_SLOTS doubles)
  double* partial = nullptr;   // reduction partials (PARTIAL_SLOTS doubles)
  double* hpin = nullptr;      // pinned host staging (HPIN_SLOTS doubles)
  int num_sms = 148;
  int coop_launch = 0;         // cudaDevAttrCooperativeLaunch, queried once at tnad_create
  int* gemm_cnt[2] = {nullptr, nullptr};   // split-K tile counters of the TMA GEMM kernel (zero between launches); [1]: products enqueued on stream2
  void* comm = nullptr;        // ncclComm_t of the chi-sharded step (tnad_comm_init; NCCL is loaded at run time)
  int comm_rank = 0, comm_world = 0;
  // host-side time per category (TNAD_HOST_PROF=1; printed by tnad_destroy): where the enqueue time of small problems goes
  bool host_prof = false;
  double hp_ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long hp_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemPool_t side_pool = nullptr;   // allocations made while c->stream == stream2 (side-stream work)
  int gemm_grid_cap = 0;       // > 0: persistent GEMM grids use at most this many CTAs (side-stream work next to a kernel that owns SMs)
  int64_t gemm_tma_n = 0, gemm_fallback_n = 0;   // products on the TMA kernel / on the cp.async kernel
  double gemm_flops = 0.0, gemm_tma_flops = 0.0; // 2 M N K batch of the products launched while kernel timing is on
  // A/B switches: every TNAD_* environment variable is read ONCE at tnad_create into this table; tnad_set_option
  // changes an entry afterwards.  Nothing on the hot path calls getenv.
  std::map<std::string, std::string> opts;
  int live_tapes = 0;          // tapes created on this context and not yet freed (tnad_destroy refuses while > 0)
  double timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // accumulators for the timing breakdown (host wall clock around synchronised phases is not
  // used; phases are bracketed by events and summed at the end of a call)
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
  std::vector<cudaEvent_t> event_pool;
  // stopwatch + optional per-kernel-family timing (bench.py)
  cudaEvent_t tstart = nullptr, tstop = nullptr;
  cudaEvent_t ev_api0 = nullptr, ev_api1 = nullptr;   // bracket of an API call (timing_begin / timing_end)
  bool ktiming = false;
  struct KSpan {
    int fam;
    cudaEvent_t a, b;
  };
  std::vector<KSpan> kspans;
};

namespace tnad {

constexpr int SCAL_SLOTS = 1024;
constexpr int PARTIAL_SLOTS = 4096;
constexpr int HPIN_SLOTS = 1 << 16;

// --------------------------------------------------------------------------------------------
// device buffers and tensors
// --------------------------------------------------------------------------------------------
// RAII host stopwatch (no-op unless tnad_ctx::host_prof): 0 plan, 1 malloc, 2 free, 3 tensor-map encode, 4 GEMM launch
struct HostTimer {
  tnad_ctx* c;
  int slot;
  long long t0;
  HostTimer(tnad_ctx* c_, int slot_);
  ~HostTimer();
};

struct DBuf {
  tnad_ctx* c;
  double* p;
  size_t n;
  DBuf(tnad_ctx* c_, size_t n_) : c(c_), p(nullptr), n(n_) {
    HostTimer ht(c, 1);
    if (c->side_pool && c->stream == c->stream2)
      TNAD_CUDA(cudaMallocFromPoolAsync((void**)&p, (n ? n : 1) * sizeof(double), c->side_pool, c->stream));
    else
      TNAD_CUDA(cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(double), c->stream));
  }
  ~DBuf() {
    HostTimer ht(c, 2);
    if (p) cudaFreeAsync(p, c->stream);
  }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

constexpr int MAXR = 10;

// A strided view of doubles on the device. dim/str follow the Julia index order (dim[0] is the
// first Julia index); a freshly allocated tensor is column-major contiguous.
struct Tens {
  double* p = nullptr;
  int rank = 0;
  int64_t dim[MAXR] = {0};
  int64_t str[MAXR] = {0};
  std::shared_ptr<DBuf> own;

  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dim[i];
    return n;
  }
  bool contiguous() const {
    int64_t s = 1;
    for (int i = 0; i < rank; ++i) {
      if (dim[i] != 1 && str[i] != s) return false;
      s *= dim[i];
    }
    return true;
  }
};

Tens t_alloc(tnad_ctx* c, std::initializer_list<int64_t> dims, bool zero = false);
Tens t_alloc_v(tnad_ctx* c, const std::vector<int64_t>& dims, bool zero = false);
Tens t_wrap(double* p, const std::vector<int64_t>& dims);                       // non-owning, contiguous
Tens t_reshape(const Tens& t, std::initializer_list<int64_t> dims);             // contiguous only
Tens t_perm(const Tens& t, std::initializer_list<int> perm);                    // view, no data movement
Tens t_slice_last(const Tens& t, int64_t start, int64_t count);                 // slice of last index
Tens t_in(tnad_ctx* c, const double* user, const std::vector<int64_t>& dims);   // host->upload / device->wrap
void t_out(tnad_ctx* c, const Tens& t, double* user);                           // contiguous tensor -> user
Tens t_clone(tnad_ctx* c, const Tens& t);                                       // contiguous copy
void t_zero(tnad_ctx* c, Tens& t);
void sync(tnad_ctx* c);
void d2h(tnad_ctx* c, double* host, const double* dev, size_t n);               // synchronous
void h2d(tnad_ctx* c, double* dev, const double* host, size_t n);

// Every C-ABI call that enqueues work is bracketed by its own pair of recorded events (see timing_end in tensor.cu)
struct ApiBracket {
  tnad_ctx* c;
  explicit ApiBracket(tnad_ctx* c);
  ~ApiBracket();
};

// timing spans (CUDA events on the ctx stream)
struct Span {
  tnad_ctx* c;
  int key;
  cudaEvent_t a, b;
  Span(tnad_ctx* c, int key);
  ~Span();
};
// RAII bracket around one kernel launch for the per-family device-time table (no-op unless enabled)
enum KFam { KF_GRAM = 0, KF_EIG = 1, KF_UPDATE = 2, KF_GEMM = 3, KF_OTHER = 4,
            KF_CHASE = 7, KF_Q2 = 8, KF_PANEL = 9, KF_SYMM = 10, KF_RANK64 = 11, KF_STEDC = 12, KF_NFAM = 16 };
struct KTimer {
  tnad_ctx* c;
  int fam;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st = nullptr;
  KTimer(tnad_ctx* c, int fam, cudaStream_t st = nullptr);
  ~KTimer();
};
cudaEvent_t get_event(tnad_ctx* c);
void timing_begin(tnad_ctx* c);
void timing_end(tnad_ctx* c);   // synchronises and sums the spans into c->timing

// --------------------------------------------------------------------------------------------
// GEMM with multi-level strides ("permute-on-load")
// --------------------------------------------------------------------------------------------
constexpr int MAXL = 4;
struct LvlSet {
  int nl;
  int n[MAXL];
  long long s[MAXL];
};
struct GemmDesc {
  int M, N, K, batch;
  LvlSet am, ak, bk, bn, cm, cn, ab, bb, cb;
  const double* A;
  const double* B;
  double* C;
  double alpha, beta;
  int a_kfast, b_kfast, a_vec, b_vec;
  int splitk;      // > 1: K is split over blockIdx.z, partial tiles go to ws
  double* ws;
};
void gemm_run(tnad_ctx* c, const GemmDesc& d);
bool gemm_tma_try(tnad_ctx* c, const GemmDesc& d);   // gemm_tma.cu: false = operands not expressible as tensor maps

// einsum pairwise contraction C[..] = alpha * sum A[..] B[..] + beta C[..]
GemmDesc contract_plan(const char* spec, const Tens& A, const Tens& B, const Tens& C);
void contract(tnad_ctx* c, const char* spec, const Tens& A, const Tens& B, Tens& C, double alpha = 1.0,
              double beta = 0.0);
// allocate the output (column-major in the label order of the spec's right-hand side) and contract
Tens contract_new(tnad_ctx* c, const char* spec, const Tens& A, const Tens& B, double alpha = 1.0);

// --------------------------------------------------------------------------------------------
// elementwise / reduction kernels (elementwise.cu)
// --------------------------------------------------------------------------------------------
// out = alpha * in + beta * out over a common index space given by out.dim; `in` must have the same
// rank and dims (use t_perm to express permutations).
void tcopy(tnad_ctx* c, const Tens& in, Tens& out, double alpha = 1.0, double beta = 0.0);
enum RedOp { RED_SUMSQ = 0, RED_DOT = 1, RED_ABSMAX = 2, RED_SUM = 3 };
// contiguous tensors; result -> device scalar *res
void reduce(tnad_ctx* c, RedOp op, const Tens& x, const Tens* y, double* res);
enum ScaleMode { SC_INV = 0, SC_INVSQRT = 1, SC_MUL = 2 };
// out = in * f(*scalar)   (INV: 1/s, INVSQRT: 1/sqrt(s), MUL: s); contiguous, may alias
void scale_dev(tnad_ctx* c, const Tens& in, Tens& out, const double* scalar, ScaleMode mode);
// xbar = ybar / sqrt(ss) - dot * x / sqrt(ss)^3  (norm pullback, autodiff.jl:23-29); ss = sum x^2
void norm_back(tnad_ctx* c, const Tens& ybar, const Tens& x, const double* ss, const double* dot, Tens& xbar);
void trace_ijij(tnad_ctx* c, const Tens& a, double* res);
// out[:, j] = in[:, j] * sqrt(S[j]) for j < k    (m x k, contiguous columns with given lds)
void colscale_sqrt(tnad_ctx* c, const double* in, int64_t ldin, const double* S, double* out, int64_t ldout,
                   int64_t m, int64_t k);
void set_identity(tnad_ctx* c, double* p, int64_t ld, int64_t n);
void init_raw(tnad_ctx* c, const Tens& bulk, Tens& corner, Tens& edge);
void init_random(tnad_ctx* c, int64_t D, int64_t chi, unsigned long long seed, Tens& corner, Tens& edge);
void ipeps_symmetrize(tnad_ctx* c, const Tens& A, Tens& xsum, Tens& out, double* ss);
void ipeps_symmetrize_back(tnad_ctx* c, const Tens& ybar, const Tens& xsum, const double* ss, Tens& Abar);
void double_layer(tnad_ctx* c, const Tens& A, Tens& ap, Tens& a);
void double_layer_back(tnad_ctx* c, const Tens& A, const Tens& apbar, const Tens& abar, Tens& Abar);
// svd_back core: panels of R = (J+J')S + S(K+K') + diag(dS)  (trg.jl:79-93) restricted to the
// first k rows (Rrow, k x n) and the first k columns below row k (Rcol, n x k, rows < k zeroed).
void svdback_panels(tnad_ctx* c, int64_t n, int64_t k, const double* S, const double* G1, const double* G2,
                    const double* dS, double eta, double* Rrow, double* Rcol);
// x[:, j] *= S[j]/(S[j]^2+eta)
void colscale_sinv(tnad_ctx* c, double* x, int64_t ld, int64_t m, int64_t k, const double* S, double eta);
// TRG: dS[j] = (sum_i U[i,j] du[i,j] + sum_i V[i,j] dvt[j,i]) / (2 sqrt(S[j])), dU = du*sqrt(S), dV = (sqrt(S) dvt)'
void trg_factor_back(tnad_ctx* c, int64_t m, int64_t n, int64_t k, const double* U, int64_t ldu, const double* V,
                     int64_t ldv, const double* S, const double* du, const double* dvt, double* dUk, double* dVk,
                     double* dS);
void add_diag_trace_back(tnad_ctx* c, Tens& abar, double w);   // abar[i,j,i,j] += w
void trg_maxval_back(tnad_ctx* c, const Tens& da, const Tens& a_in, double maxval, double coef, Tens& da_in);
void fill(tnad_ctx* c, double* p, int64_t n, double v);
// option table of the context (see tnad_ctx::opts)
const char* opt_s(const tnad_ctx* c, const char* name);                  // nullptr when unset
int opt_i(const tnad_ctx* c, const char* name, int dflt);
double opt_d(const tnad_ctx* c, const char* name, double dflt);
// Canonical gauge of a decomposition (the "sign-fix" of SURVEY appendix A.10): every column j < ncols of U (m rows) is
// multiplied by the sign that makes its largest-magnitude entry (first one on ties) positive; the same sign is applied
// to column j of V (nv rows; V may be null).  U diag(S) V' is unchanged.
void signfix_cols(tnad_ctx* c, double* U, int64_t ldu, int64_t m, double* V, int64_t ldv, int64_t nv, int64_t ncols);

// --------------------------------------------------------------------------------------------
// Jacobi SVD (jacobi.cu)
// --------------------------------------------------------------------------------------------
struct SvdResult {
  Tens U;   // m x k
  Tens S;   // k
  Tens V;   // n x k
  std::vector<double> s_host;
  double null_thr = 0.0;   // singular values <= null_thr are numerically null (their U columns are a completion)
  int sweeps = 0;
  int64_t rank_left = -1;   // >= 0: only the first rank_left columns of U are valid (null columns left zero)
  int64_t rank_right = -1;  // >= 0: the same for V (Jordan-Wielandt route; the Jacobi route always returns the full V)
};
// A: any rank-2 strided view (m x n). If sym_add_transpose, the matrix decomposed is A + A^T (m == n).
// V0 (optional, n x n orthogonal): warm start for square inputs.
SvdResult svd_jacobi(tnad_ctx* c, const Tens& A, bool sym_add_transpose = false, const Tens* V0 = nullptr,
                     bool complete_null = true);

// Symmetric input (ctmrg.jl:135-136): two-sided block Jacobi eigensolver, M = Q L Q' -> U = Q, S = |L|, V = Q sign(L).
SvdResult svd_symmetric(tnad_ctx* c, const Tens& A, bool sym_add_transpose = false, const Tens* Q0 = nullptr);
void symeig_cache_free(tnad_ctx* c);

}  // namespace tnad
