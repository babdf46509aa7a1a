// Direct symmetric eigensolver: Householder tridiagonalisation (tridiag.cu) + divide and conquer (stedc.cu).
#pragma once
#include "common.h"
#include <functional>

namespace tnad {

// A = Q T Q' (A overwritten; Vh zero-initialised n x n receives the reflectors; dd, ee: the tridiagonal T)
void sytrd(tnad_ctx* c, double* A, int64_t lda, int64_t n, double* Vh, int64_t ldv, double* tau, double* dd, double* ee);
int64_t sytrd_vcols(int64_t n);   // columns Vh / entries tau must provide (zero-initialised)
// Z[0:n, 0:ncols] <- Q Z
// (offset: row distance between a reflector's column index and its unit element: 1 for sytrd, 32 for the band reduction)
void apply_q(tnad_ctx* c, const double* Vh, int64_t ldv, const double* tau, int64_t n, double* Z, int64_t ldz, int64_t ncols,
             int64_t offset = 1);
// All eigenpairs of the symmetric tridiagonal (dd, ee) of order n.  The problem is padded to N = s 2^L >= n
// (s <= 64) with decoupled diagonal entries above the spectrum; lam (N) and Z (N x N, leading dimension N) come
// back unsorted, the pad eigenvalues being the N - n largest, their eigenvectors unit vectors in the pad rows.
void stedc(tnad_ctx* c, const double* dd, const double* ee, int64_t n, Tens& lam, Tens& Z, int64_t& N);
// ---- two-stage route (band.cu) ----
int64_t chase_positions(int64_t n);
void extract_band(tnad_ctx* c, const double* A, int64_t lda, int64_t n, double* AB, int64_t ldab);
void sy2sb(tnad_ctx* c, double* A, int64_t lda, int64_t n, double* Yst, int64_t ldy, double* tau1);
// overlap(ctas): called right after the chase kernel is launched (it occupies `ctas` SMs) and before the host waits for it
void sb2st(tnad_ctx* c, const double* AB, int64_t ldab, int64_t n, double* dd, double* ee, double* V2, int64_t ldv, double* tau2,
           const std::function<void(int)>& overlap = {});
// x_is_identity: X = I (n x n) on entry -- the pass skips the sweeps that only see zeros
void apply_q2(tnad_ctx* c, const double* V2, int64_t ldv, const double* tau2, int64_t n, double* X, int64_t ldx, int64_t ncols,
              bool x_is_identity = false);
// three-phase form of the direct symmetric eigensolver (reduce / back-transform a column block / finish)
struct EigFactor {
  int64_t n = 0, N = 0, ldv2 = 0;
  bool two_stage = false;
  Tens Vh, tau;          // reflectors of the (first) reduction stage
  Tens V2, tau2;         // reflectors of the chase (two-stage route)
  Tens Z;                // eigenvectors of the tridiagonal matrix (N x N, N >= n padded)
  // explicit-Q mode (two-stage route, n <= TNAD_EXPLICITQ_MAX): Q1 is formed on stream2 while the chase runs, Q2 and
  // Qfull = Q1 Q2 while the divide and conquer runs; the back-transformation is then ONE product Qfull Z
  Tens Q1x, Q2x, Qfull;
  cudaEvent_t q_ready = nullptr;   // recorded on stream2 after Qfull
  std::vector<double> lam;
};
void load_symmetric(tnad_ctx* c, const Tens& A, bool sym_add_transpose, Tens& Aw);
// want_q: the caller will back-transform ALL columns on this context (enables the explicit-Q mode)
void symeig_reduce(tnad_ctx* c, Tens& Aw, int64_t n, EigFactor& f, bool want_q = false);
void symeig_backtransform(tnad_ctx* c, const EigFactor& f, double* Zc, int64_t ldz, int64_t ncols);
SvdResult symeig_finish(tnad_ctx* c, const EigFactor& f, const double* Zfull, int64_t ldz = 0);   // ldz = 0: f.N
SvdResult svd_symmetric_dc(tnad_ctx* c, const Tens& A, bool sym_add_transpose);
// general (rank-2 or rank-4 [(d1,d2),(d3,d4)] view) matrix through the Jordan-Wielandt embedding; only the r non-null
// triplets are formed (rank_left = rank_right = r)
SvdResult svd_general_dc(tnad_ctx* c, const Tens& A);
// Solver selection for symmetric inputs: TNAD_SYMEIG = 1 block Jacobi (symeig.cu), 2 tridiagonal divide and conquer;
// default: divide and conquer from n >= TNAD_DC_MIN (48) on (measured with the cluster kernel: 0.88 vs 1.34 ms at n = 64).
SvdResult svd_symmetric_auto(tnad_ctx* c, const Tens& A, bool sym_add_transpose, const Tens* Q0 = nullptr);

}  // namespace tnad
