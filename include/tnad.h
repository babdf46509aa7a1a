/*
 * tnad.h -- C ABI of libtnad_b200.so: the B200 (sm_100a) implementation of the
 * TRG / CTMRG / iPEPS-energy hot path of under-Peter/TensorNetworkAD.jl.
 *
 * The reference is pure Julia and has no FFI boundary of its own; every entry
 * point below replaces the *body* of one reference function (cited as
 * file:line into the reference tree) so that the Julia signatures stay as they
 * are and call in here through `ccall` (see INTEGRATION.md for the stubs).
 *
 * Conventions
 *   - All arrays are caller-owned, column-major (Julia / Fortran order) `double`.
 *     By default they are HOST pointers: the library copies in and out.  With
 *     tnad_set_pointer_mode(ctx, TNAD_POINTER_DEVICE) array arguments are device
 *     pointers on the context's device (scalars and int outputs stay on the host).
 *   - Every function returns an int status (TNAD_OK == 0).  A human-readable
 *     message for the last failure is available from tnad_last_error().  No C++
 *     exception crosses this boundary.
 *   - One tnad_ctx == one device + one stream + its workspace.  Calls on one
 *     context are serialised by the caller; different contexts are independent.
 *     Results are valid on return (synchronous semantics).
 *   - Tapes are opaque, created by a forward call, consumed by any number of
 *     backward calls, released with tnad_tape_free.  A tape belongs to its context:
 *     tnad_destroy fails (TNAD_ERR_ARG) while tapes of the context are alive.
 *   - There is no CPU fallback: without a usable sm_100 device tnad_create fails.
 */
#ifndef TNAD_H
#define TNAD_H

#include <stdint.h>

#if defined(__GNUC__)
#define TNAD_API __attribute__((visibility("default")))
#else
#define TNAD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnad_ctx tnad_ctx;
typedef struct tnad_tape tnad_tape;

enum {
  TNAD_OK = 0,
  TNAD_ERR_ARG = 1,        /* bad argument / dimension mismatch (DimensionMismatch in Julia) */
  TNAD_ERR_CUDA = 2,       /* CUDA runtime error */
  TNAD_ERR_NOMEM = 3,      /* device out of memory */
  TNAD_ERR_NOCONV = 4,     /* Jacobi SVD did not converge */
  TNAD_ERR_NCCL = 5,
  TNAD_ERR_INTERNAL = 6
};

enum { TNAD_POINTER_HOST = 0, TNAD_POINTER_DEVICE = 1 };
enum { TNAD_ENV_RAW = 0, TNAD_ENV_GIVEN = 1 };

/* ---- lifecycle ------------------------------------------------------------------------- */
TNAD_API int tnad_version(void);
TNAD_API int tnad_create(int device, tnad_ctx** out);
TNAD_API int tnad_destroy(tnad_ctx* ctx);
TNAD_API const char* tnad_last_error(tnad_ctx* ctx);       /* ctx may be NULL: message of a failed tnad_create */
TNAD_API int tnad_set_pointer_mode(tnad_ctx* ctx, int mode);
/* A/B switches.  Every TNAD_* environment variable is read once, at tnad_create; afterwards this call
 * changes an entry of the context's table (value == NULL removes it).  Nothing on the hot path reads the environment. */
TNAD_API int tnad_set_option(tnad_ctx* ctx, const char* name, const char* value);
TNAD_API int tnad_synchronize(tnad_ctx* ctx);
/* counters: kernels launched by this library on ctx since creation / last reset */
TNAD_API int64_t tnad_launch_count(tnad_ctx* ctx);
TNAD_API int tnad_reset_launch_count(tnad_ctx* ctx);
/* device memory helpers for callers that keep inputs resident in HBM (bench `value`) */
TNAD_API int tnad_dev_alloc(tnad_ctx* ctx, int64_t ndoubles, double** dptr);
TNAD_API int tnad_dev_free(tnad_ctx* ctx, double* dptr);
TNAD_API int tnad_dev_upload(tnad_ctx* ctx, double* dptr, const double* host, int64_t ndoubles);
TNAD_API int tnad_dev_download(tnad_ctx* ctx, double* host, const double* dptr, int64_t ndoubles);

/* ---- L1 building blocks (step-level entry points used by the parity tests) --------------- */

/* OMEinsum pairwise `ein"spec"(A, B)` (every contraction call site of trg.jl:25, ctmrg.jl:130-140,
 * variationalipeps.jl:52-54 is run as a chain of these).  spec like "iba,ad->ibd"; every label
 * is one ASCII letter.  C = alpha * contraction + beta * C.  dims are the Julia sizes. */
TNAD_API int tnad_contract(tnad_ctx* ctx, const char* spec,
                  const double* A, const int64_t* dimsA, int rankA,
                  const double* B, const int64_t* dimsB, int rankB,
                  double alpha, double beta, double* C);
/* Host-only: the GEMM plan the library would run for `spec` (no GPU needed).  Fills
 * plan[0..63] (layout documented in csrc/contract.cu) so tests can check the stride algebra. */
TNAD_API int tnad_contract_plan(const char* spec, const int64_t* dimsA, int rankA,
                       const int64_t* dimsB, int rankB, int64_t* plan);

/* Gauge of every decomposition returned by this library (the "sign-fix"): LAPACK leaves the column signs of U to
 * chance; here column j of U is oriented so that its largest-magnitude entry (first one on ties)
 * is positive and column j of V follows, fused into the kernel that sorts / truncates the vectors.  lnZ, energies,
 * magnetisation and all gradients do not depend on it; corner / edge become comparable entry by entry. */

/* LinearAlgebra.svd(A) as used at trg.jl:36 and ctmrg.jl:136: thin SVD, A (m x n) = U diag(S) V^T,
 * k = min(m,n), U m x k, S k (descending), V n x k.  One-sided block Jacobi on the device. */
TNAD_API int tnad_svd(tnad_ctx* ctx, const double* A, int m, int n, double* U, double* S, double* V,
             int* sweeps_out /* may be NULL */);

/* svd of a symmetric matrix as it occurs at ctmrg.jl:135-136 (cpmat + cpmat'): A = Q L Q' returned as U = Q,
 * S = |L| (descending), V = Q sign(L).  Solver: the direct one (two-stage tridiagonalisation, divide and conquer,
 * back-transformation; default from n >= 48) or two-sided block Jacobi (below that; option TNAD_SYMEIG = 1 | 2 forces
 * one).  A is symmetrised as (A + A')/2. */
TNAD_API int tnad_svd_sym(tnad_ctx* ctx, const double* A, int n, double* U, double* S, double* V, int* sweeps_out);

/* trg_svd(t, dmax, tol)  (trg.jl:33-44).  t is (d1,d2,d3,d4); u gets (d1,d2,k), v gets (k,d3,d4);
 * the buffers must hold dmax columns/rows; *k_out = kept rank by the reference's rule
 * k = min(searchsortedfirst(s, tol, rev=true), dmax, length(s)) applied to the computed spectrum.
 * Noise floor: singular values <= 16 eps sqrt(max(m,n)) |t|_F are numerically null; LAPACK returns rounding noise
 * there, this library returns exact zeros for them (their factor columns are zero and carry no cotangent).  The kept
 * rank still follows the rule literally, so shapes equal the reference's for tol >= that floor. */
TNAD_API int tnad_trg_svd(tnad_ctx* ctx, const double* t, int d1, int d2, int d3, int d4, int dmax, double tol,
                 double* u, double* v, int* k_out);

/* svd_back(U,S,V,dU,dS,dV; eta)  (trg.jl:72-105), real case.  U m x k, S k, V n x k; any of the
 * cotangents may be NULL (`nothing`); dA is m x n. */
TNAD_API int tnad_svd_back(tnad_ctx* ctx, int m, int n, int k, const double* U, const double* S, const double* V,
                  const double* dU, const double* dS, const double* dV, double eta, double* dA);

/* ---- TRG (trg.jl:13-30 and its Zygote pullback) ------------------------------------------ */
/* a is (d0,d1,d2,d3) with d0==d2, d1==d3 (the reference passes 2x2x2x2). tape may be NULL. */
TNAD_API int tnad_trg_forward(tnad_ctx* ctx, const double* a, int d0, int d1, int chi, int niter, double tol,
                     double* lnZ, tnad_tape** tape);
/* da (same shape as a) = dlnZ * d lnZ / d a */
TNAD_API int tnad_trg_backward(tnad_ctx* ctx, tnad_tape* tape, double dlnZ, double* da);
TNAD_API int tnad_tape_free(tnad_tape* tape);

/* ---- CTMRG (ctmrg.jl:66-153, fixedpoint.jl:11-41) ---------------------------------------- */
/* _initializect_square(bulk, Val(:raw), chi)  (ctmrg.jl:74-86) */
TNAD_API int tnad_ctmrg_init_raw(tnad_ctx* ctx, const double* bulk, int D, int chi, double* corner, double* edge);
/* _initializect_square(bulk, Val(:random), chi)  (ctmrg.jl:66-72) generated on the device: randn from a counter-based generator
 * (reproducible from `seed`; the reference uses Julia's global RNG, whose stream cannot be matched), corner += corner',
 * edge += permutedims(edge, (3,2,1)) */
TNAD_API int tnad_ctmrg_init_random(tnad_ctx* ctx, int D, int chi, unsigned long long seed, double* corner, double* edge);
/* one ctmrgstep (ctmrg.jl:126-153): bulk D^4, corner chi^2, edge chi*D*chi; vals has chi*D entries */
TNAD_API int tnad_ctmrgstep(tnad_ctx* ctx, const double* bulk, int D, int chi,
                   const double* corner_in, const double* edge_in,
                   double* corner_out, double* edge_out, double* vals);
/* Pullback of ONE ctmrgstep (what Zygote derives from ctmrg.jl:126-153 with the rules of autodiff.jl and trg.jl:55-105):
 * given the cotangents of the step's outputs returns those of bulk (D^4, required), corner_in and edge_in (optional).
 * The forward step is recomputed internally. */
TNAD_API int tnad_ctmrgstep_backward(tnad_ctx* ctx, const double* bulk, int D, int chi,
                            const double* corner_in, const double* edge_in,
                            const double* dcorner_out, const double* dedge_out,
                            double* dbulk, double* dcorner_in, double* dedge_in);
/* ctmrg(rt; tol, maxit) with the reference stop rule (counter from -1). corner/edge are in/out.
 * steps_done, vals (chi*D) and tape may be NULL. */
TNAD_API int tnad_ctmrg(tnad_ctx* ctx, const double* bulk, int D, int chi, double* corner, double* edge,
               double tol, int maxit, int* steps_done, double* vals, tnad_tape** tape);
/* Unrolled reverse sweep through every executed step.  dcorner/dedge: cotangents of the returned
 * environment.  dbulk (D^4) is required; dcorner0/dedge0 (cotangents of the initial environment)
 * may be NULL (the reference marks the initialisation @nograd, autodiff.jl:5). */
TNAD_API int tnad_ctmrg_backward(tnad_ctx* ctx, tnad_tape* tape, const double* dcorner, const double* dedge,
                        double* dbulk, double* dcorner0, double* dedge0);

/* ---- iPEPS energy (variationalipeps.jl:28-56, ipeps.jl:32-39) ----------------------------- */
/* expectationvalue(h, ap, rt): h (s,s,s,s), ap (D,D,D,D,s,s) */
TNAD_API int tnad_expectationvalue(tnad_ctx* ctx, const double* h, const double* ap, int D, int s,
                          const double* corner, const double* edge, int chi, double* e);
/* pullback of expectationvalue: cotangents of ap (D,D,D,D,s,s), corner and edge for the output cotangent ybar; any
 * output may be NULL */
TNAD_API int tnad_expectationvalue_backward(tnad_ctx* ctx, const double* h, const double* ap, int D, int s,
                                   const double* corner, const double* edge, int chi, double ybar,
                                   double* dap, double* dcorner, double* dedge);
/* energy(h, ipeps; chi, tol, maxit) and, when gradA != NULL, its gradient w.r.t. ipeps.bulk
 * exactly as Zygote + src/autodiff.jl compute it.  A is (d,d,d,d,s). steps_done may be NULL. */
TNAD_API int tnad_energy(tnad_ctx* ctx, const double* h, const double* A, int d, int s, int chi,
                double tol, int maxit, double* e, double* gradA, int* steps_done);
/* The same energy with the IMPLICIT (fixed-point) gradient, opt-in: CTMRG runs without a tape until the reference's stop
 * rule fires, one more step is recorded at the final environment, and the gradient is the Neumann series
 * sum_k J_a' (J_x')^k xbar of that step's pullback, truncated when the cotangent has decayed by bwd_tol (or after bwd_maxit
 * terms; *bwd_iters returns the number applied).  Agrees with tnad_energy's gradient -- i.e. with what Zygote computes
 * for the reference -- in the limit of a converged environment; memory is one step record instead of one per step.
 * Not what the reference computes for short fixed-maxit runs (it treats the initial environment as a constant). */
TNAD_API int tnad_energy_fixedpoint(tnad_ctx* ctx, const double* h, const double* A, int d, int s, int chi,
                           double tol, int maxit, double bwd_tol, int bwd_maxit,
                           double* e, double* gradA, int* steps_done, int* bwd_iters);
/* magnetisation read-out (exampletensors.jl:63-68): |<env,m>/<env,a>| */
TNAD_API int tnad_magnetisation_readout(tnad_ctx* ctx, const double* a, const double* m, int D,
                               const double* corner, const double* edge, int chi, double* mag);

/* pullback of the read-out (test/ctmrg.jl:44-46 differentiates magnetisation with Zygote): cotangents of a, m,
 * corner, edge for the output cotangent ybar; any output may be NULL.  Chained with tnad_ctmrg_backward (whose
 * dbulk adds to da) this gives d magnetisation / d beta. */
TNAD_API int tnad_magnetisation_backward(tnad_ctx* ctx, const double* a, const double* m, int D,
                                const double* corner, const double* edge, int chi, double ybar,
                                double* da, double* dm, double* dcorner, double* dedge);

/* ---- multi-GPU: independent instances (beta sweeps, parameter scans; no communication) ---------------------- */
/* ninst TRG instances (tensors: ninst arrays (d0,d1,d0,d1) back to back, HOST memory), instance i on device
 * devices[i mod ngpu] (devices == NULL: 0..ngpu-1), one context and one host thread per device.  lnZ[ninst];
 * grads (ninst arrays like the tensors: d lnZ / d a) may be NULL.  errbuf receives the first failure message. */
TNAD_API int tnad_trg_sweep(const double* tensors, int ninst, int d0, int d1, int chi, int niter, double tol,
                   int ngpu, const int* devices, double* lnZ, double* grads, char* errbuf, int errlen);

/* ---- pieces used by the chi-sharded multi-GPU step (tensornetworkad.jl_b200/sharded.py) -------- */
/* svd(A + A') for a square A (ctmrg.jl:133-135: `cpmat += adjoint(cpmat); svd(cpmat)`), same solver as tnad_svd_sym */
TNAD_API int tnad_svd_symmetrized(tnad_ctx* ctx, const double* A, int n, double* U, double* S, double* V, int* sweeps_out);
/* The symmetric eigensolver behind tnad_svd_sym / tnad_svd_symmetrized in three phases, so that several GPUs can
 * share its back-transformation (the chi-sharded step: every rank reduces the replicated matrix, back-transforms its
 * own block of columns, the blocks are all-gathered, every rank finishes):
 *   reduce         A (+ A' if add_transpose) -> tridiagonal (one- or two-stage) and its eigen-decomposition;
 *                  *N_out = padded order N >= n of the eigenvector matrix
 *   backtransform  Zcols (N x ncols, leading dimension N) = columns col0 .. col0+ncols-1 of Q Z
 *   finish         Zfull (N x N, all columns back-transformed) -> U, S, V as tnad_svd_sym returns them */
typedef struct tnad_eig tnad_eig;
TNAD_API int tnad_symeig_reduce(tnad_ctx* ctx, const double* A, int n, int add_transpose, tnad_eig** out, int* N_out);
TNAD_API int tnad_symeig_backtransform(tnad_ctx* ctx, tnad_eig* h, int col0, int ncols, double* Zcols);
TNAD_API int tnad_symeig_finish(tnad_ctx* ctx, tnad_eig* h, const double* Zfull, double* U, double* S, double* V);
TNAD_API int tnad_symeig_free(tnad_eig* h);
/* stages of the direct symmetric eigensolver (tridiag.cu / stedc.cu), exported for the parity tests:
   A = Q T Q' with T = tridiag(d, e) (Householder, LAPACK dsytrd semantics; Q explicit n x n) ... */
TNAD_API int tnad_sytrd(tnad_ctx* ctx, const double* A, int n, double* d, double* e, double* Q);
/* the same factorisation through the two-stage route (band.cu): dense -> band (half bandwidth 32, panel QR inside a
   thread-block cluster + DMMA trailing updates) -> tridiagonal (systolic bulge chasing); Q = Q1 Q2 explicit.
   band (33 x n, lower band storage of the intermediate band matrix) may be NULL.  n >= 3. */
TNAD_API int tnad_sytrd2(tnad_ctx* ctx, const double* A, int n, double* d, double* e, double* Q, double* band);
/* ... and all eigenpairs of tridiag(d, e), ascending (divide and conquer, LAPACK dstedc semantics) */
TNAD_API int tnad_stedc(tnad_ctx* ctx, const double* d, const double* e, int n, double* lam, double* Z);
/* permutedims(in, perm) (0-based perm, Julia semantics: out dim i = in dim perm[i]) */
TNAD_API int tnad_permute(tnad_ctx* ctx, const double* in, const int64_t* dims, int rank, const int* perm, double* out);
/* tail of ctmrgstep (ctmrg.jl:145-150): corner += corner', edge += permutedims(edge,(3,2,1)), each / its norm */
TNAD_API int tnad_ctmrg_finish(tnad_ctx* ctx, const double* c1, const double* e1, int D, int chi,
                               double* corner_out, double* edge_out);

/* ---- chi-sharded ctmrgstep over several GPUs (one process per GPU; north_star: "at chi >= 256 the enlarged-corner GEMMs
 * are sharded along the chi index with an NCCL all-gather over NVLink before the SVD"; replaces the OMEinsum / LAPACK calls of
 * ctmrg.jl:126-142 on every rank).  NCCL is loaded at run time: nccl_path = path of libnccl.so.2, or NULL to use the copy the
 * host process has already mapped.  Rank 0 obtains a 128-byte unique id, the application distributes it (MPI,
 * torch.distributed, a file), every rank calls tnad_comm_init on its own context; tnad_comm_init(ctx, .., 0, 1) without a
 * library makes the sharded entry point run on one GPU.  tnad_ctmrgstep_sharded: every rank passes the same bulk / corner /
 * edge (pointer mode applies; chi divisible by the number of ranks) and receives the full corner_out / edge_out; vals (chi*D,
 * host) and ms3 (host, optional: device ms of contractions, all-gathers, eigen-decomposition) as in tnad_ctmrgstep.
 * Opt-in (tnad_set_option(ctx, "TNAD_SHARDED_LOOP", "1") after tnad_comm_init): the fixed-point loops of this context
 * (tnad_ctmrg, tnad_energy, tnad_energy_fixedpoint) run their steps through the sharded step as well and record them in full, so
 * tapes and gradients work unchanged: the forward pass is shared between the ranks, the reverse sweep is replicated.  Those calls
 * are then COLLECTIVE: every rank of the communicator must make them. */
TNAD_API int tnad_nccl_unique_id(const char* nccl_path, unsigned char* id128);
TNAD_API int tnad_comm_init(tnad_ctx* ctx, const char* nccl_path, const unsigned char* id128, int rank, int world);
TNAD_API int tnad_comm_destroy(tnad_ctx* ctx);
/* ctmrg + fixedpoint + StopFunction (ctmrg.jl:110-117, fixedpoint.jl:11-41) over the communicator: corner / edge in-out as in
 * tnad_ctmrg (no tape: forward only) */
TNAD_API int tnad_ctmrg_sharded(tnad_ctx* ctx, const double* bulk, int D, int chi, double* corner, double* edge, double tol, int maxit,
                                int* steps_done, double* vals);
TNAD_API int tnad_ctmrgstep_sharded(tnad_ctx* ctx, const double* bulk, int D, const double* corner, const double* edge, int chi,
                                    double* corner_out, double* edge_out, double* vals, double* ms3);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------- */
/* pinned host memory for end-to-end runs */
TNAD_API int tnad_host_alloc(tnad_ctx* ctx, int64_t ndoubles, double** hptr);
TNAD_API int tnad_host_free(tnad_ctx* ctx, double* hptr);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel of ctx is launched on) */
TNAD_API int tnad_timer_start(tnad_ctx* ctx);
TNAD_API int tnad_timer_stop(tnad_ctx* ctx, double* ms);
/* per-kernel-family device time: enable, run, then read (16 slots).  families: 0 jacobi_gram, 1 one-stage
 * tridiagonalisation panel / Jacobi pivot kernels, 2 jacobi_update, 3 gemm_tma / gemm_dmma (every einsum contraction and the GEMMs
 * of the back-transformation and of the divide and conquer), 4 other, 5/6 work counters of the Jacobi kernels,
 * 7 k_chase (band -> tridiagonal), 8 k_q2_stage (back-transformation with Q2), 9 k_panel_gram / k_panel_qr (panel QR of the
 * band reduction), 10 k_symm_y (+ k_reduce_g), 11 k_rank64_update, 12 divide-and-conquer kernels other than its GEMMs.
 * ms[i] = summed duration, count[i] = launches.  Slots 13 / 14 describe family 3: ms[13] = sum of 2 M N K batch over its
 * launches, ms[14] = the part of it on the TMA kernel, count[13] / count[14] = products on the TMA kernel / on the cp.async
 * kernel. */
TNAD_API int tnad_set_kernel_timing(tnad_ctx* ctx, int enable);
TNAD_API int tnad_kernel_timing(tnad_ctx* ctx, double* ms /* [16] */, int64_t* count /* [16] */);
/* FP64 tensor-core (DMMA m8n8k4) issue-rate microbenchmark: register-resident operands, no memory
 * traffic.  The measured TFLOP/s is the denominator of the `tensor` roofline (MEASURED_PEAKS.json
 * has no FP64 figure). */
TNAD_API int tnad_dmma_peak(tnad_ctx* ctx, double* tflops);
/* achieved TFLOP/s of the contraction GEMM on a plain m x k x n product (device buffers, timed with events) */
TNAD_API int tnad_gemm_bench(tnad_ctx* ctx, int m, int n, int k, int reps, double* tflops);

/* ---- timing breakdown of the last energy / ctmrg / trg call (milliseconds, CUDA events) ---- */
/* keys: 0 total, 1 svd, 2 contractions (forward), 3 backward total, 4 svd_back. */
TNAD_API int tnad_last_timing(tnad_ctx* ctx, double* ms /* [8] */);

#ifdef __cplusplus
}
#endif
#endif /* TNAD_H */
