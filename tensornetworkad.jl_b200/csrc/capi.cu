// extern "C" surface of libtnad_b200.so (see include/tnad.h).  No exception crosses this boundary.
#include "drivers.h"
#include "eigdc.h"
#include <algorithm>
#include <mutex>
#include <thread>

#include <unistd.h>
extern char** environ;

using namespace tnad;

static std::string g_create_error;

#define TNAD_API_BEGIN(ctx)                       \
  if (!(ctx)) return TNAD_ERR_ARG;                \
  try {                                           \
    TNAD_CUDA(cudaSetDevice((ctx)->device));      \
    tnad::ApiBracket _bracket(ctx);

#define TNAD_API_END(ctx)                                           \
    return TNAD_OK;                                                 \
  } catch (const tnad::Error& e) {                                  \
    (ctx)->err = e.msg;                                             \
    cudaGetLastError();                                             \
    return e.code;                                                  \
  } catch (const std::bad_alloc&) {                                 \
    (ctx)->err = "host out of memory";                              \
    return TNAD_ERR_NOMEM;                                          \
  } catch (const std::exception& e) {                               \
    (ctx)->err = std::string("internal error: ") + e.what();        \
    return TNAD_ERR_INTERNAL;                                       \
  } catch (...) {                                                   \
    (ctx)->err = "unknown internal error";                          \
    return TNAD_ERR_INTERNAL;                                       \
  }

extern "C" {

int tnad_version(void) { return 100; }

int tnad_create(int device, tnad_ctx** out) {
  if (!out) return TNAD_ERR_ARG;
  *out = nullptr;
  tnad_ctx* c = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(TNAD_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                              "); libtnad_b200 has no CPU fallback");
    TNAD_REQUIRE(device >= 0 && device < ndev, "tnad_create: device index out of range");
    cudaDeviceProp prop;
    TNAD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      fail(TNAD_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                              std::to_string(prop.minor) + "; libtnad_b200 is built for sm_100a only");
    TNAD_CUDA(cudaSetDevice(device));
    c = new tnad_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    TNAD_CUDA(cudaDeviceGetAttribute(&c->coop_launch, cudaDevAttrCooperativeLaunch, device));
    for (char** e = environ; e && *e; ++e)     // A/B switches: read the environment once, here
      if (strncmp(*e, "TNAD_", 5) == 0)
        if (const char* eq = strchr(*e, '=')) c->opts[std::string(*e, eq - *e)] = std::string(eq + 1);
    c->host_prof = opt_i(c, "TNAD_HOST_PROF", 0) != 0;
    // the main stream carries the latency-critical pivot kernels of the eigensolver: give it the highest
    // priority so its CTAs are scheduled ahead of the bulk update kernels running on streams 2 and 3
    int prio_lo = 0, prio_hi = 0;
    TNAD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TNAD_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    TNAD_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, opt_i(c, "TNAD_SIDE_SAME_PRIO", 0) ? prio_hi : prio_lo));
    TNAD_CUDA(cudaEventCreateWithFlags(&c->ev_eig, cudaEventDisableTiming));
    TNAD_CUDA(cudaEventCreateWithFlags(&c->ev_rest, cudaEventDisableTiming));
    TNAD_CUDA(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_lo));
    TNAD_CUDA(cudaEventCreateWithFlags(&c->ev_v, cudaEventDisableTiming));
    TNAD_CUDA(cudaMalloc((void**)&c->scal, SCAL_SLOTS * sizeof(double)));
    TNAD_CUDA(cudaMalloc((void**)&c->partial, PARTIAL_SLOTS * sizeof(double)));
    TNAD_CUDA(cudaMallocHost((void**)&c->hpin, HPIN_SLOTS * sizeof(double)));
    TNAD_CUDA(cudaMemset(c->scal, 0, SCAL_SLOTS * sizeof(double)));
    // keep freed blocks in the stream-ordered pool: the CTMRG loop re-allocates the same sizes every step
    cudaMemPool_t pool;
    TNAD_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    TNAD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // Work enqueued on the side stream (explicit-Q mode of the eigensolver) allocates from its OWN pool: blocks that
    // migrate between two streams of one pool made the allocator insert cross-stream waits or grow the pool in the
    // middle of a call -- rare stalls of 0.1 - 2 s (measured: 1 call in 5 at d = 4, chi = 128, none with this pool)
    {
      cudaMemPoolProps pp;
      memset(&pp, 0, sizeof(pp));
      pp.allocType = cudaMemAllocationTypePinned;
      pp.handleTypes = cudaMemHandleTypeNone;
      pp.location.type = cudaMemLocationTypeDevice;
      pp.location.id = device;
      TNAD_CUDA(cudaMemPoolCreate(&c->side_pool, &pp));
      TNAD_CUDA(cudaMemPoolSetAttribute(c->side_pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    *out = c;
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    g_create_error = e.msg;
    if (c) delete c;
    return e.code;
  } catch (...) {
    g_create_error = "unknown error in tnad_create";
    if (c) delete c;
    return TNAD_ERR_INTERNAL;
  }
}

int tnad_destroy(tnad_ctx* c) {
  if (!c) return TNAD_OK;
  if (c->live_tapes > 0) {   // a tape owns device buffers that are released on this context's stream
    c->err = "tnad_destroy: " + std::to_string(c->live_tapes) + " tape(s) of this context are still alive; call tnad_tape_free first";
    return TNAD_ERR_ARG;
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->host_prof) {
    static const char* nm[8] = {"contract_plan", "cudaMallocAsync", "cudaFreeAsync", "tensor-map encode (2 per product)", "TMA GEMM launch", "", "", ""};
    for (int i = 0; i < 5; ++i)
      fprintf(stderr, "[tnad host] %-34s %9lld calls  %8.2f ms total  %6.2f us each\n", nm[i], c->hp_n[i], c->hp_ns[i] * 1e-6,
              c->hp_n[i] ? c->hp_ns[i] * 1e-3 / c->hp_n[i] : 0.0);
  }
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (c->stream3) cudaStreamSynchronize(c->stream3);
  for (auto& s : c->spans) {
    cudaEventDestroy(s.second.first);
    cudaEventDestroy(s.second.second);
  }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  for (auto& k : c->kspans) {
    cudaEventDestroy(k.a);
    cudaEventDestroy(k.b);
  }
  if (c->tstart) cudaEventDestroy(c->tstart);
  if (c->tstop) cudaEventDestroy(c->tstop);
  tnad_comm_destroy(c);
  if (c->side_pool) cudaMemPoolDestroy(c->side_pool);
  cudaFree(c->scal);
  for (int* p : c->gemm_cnt)
    if (p) cudaFree(p);
  cudaFree(c->partial);
  cudaFreeHost(c->hpin);
  tnad::symeig_cache_free(c);
  if (c->ev_eig) cudaEventDestroy(c->ev_eig);
  if (c->ev_rest) cudaEventDestroy(c->ev_rest);
  if (c->ev_v) cudaEventDestroy(c->ev_v);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  cudaStreamDestroy(c->stream);
  delete c;
  return TNAD_OK;
}

int tnad_set_option(tnad_ctx* c, const char* name, const char* value) {
  if (!c || !name) return TNAD_ERR_ARG;
  if (value) c->opts[name] = value;
  else c->opts.erase(name);
  return TNAD_OK;
}

const char* tnad_last_error(tnad_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int tnad_set_pointer_mode(tnad_ctx* c, int mode) {
  if (!c || (mode != TNAD_POINTER_HOST && mode != TNAD_POINTER_DEVICE)) return TNAD_ERR_ARG;
  c->pointer_mode = mode;
  return TNAD_OK;
}

int tnad_synchronize(tnad_ctx* c) {
  TNAD_API_BEGIN(c)
  sync(c);
  TNAD_API_END(c)
}

int64_t tnad_launch_count(tnad_ctx* c) { return c ? c->launches : -1; }
int tnad_reset_launch_count(tnad_ctx* c) {
  if (!c) return TNAD_ERR_ARG;
  c->launches = 0;
  return TNAD_OK;
}

int tnad_dev_alloc(tnad_ctx* c, int64_t n, double** dptr) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(dptr && n >= 0, "tnad_dev_alloc: bad arguments");
  TNAD_CUDA(cudaMalloc((void**)dptr, (size_t)(n ? n + 2 : 2) * sizeof(double)));
  TNAD_API_END(c)
}
int tnad_dev_free(tnad_ctx* c, double* dptr) {
  TNAD_API_BEGIN(c)
  sync(c);
  TNAD_CUDA(cudaFree(dptr));
  TNAD_API_END(c)
}
int tnad_dev_upload(tnad_ctx* c, double* dptr, const double* host, int64_t n) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(dptr && host && n >= 0, "tnad_dev_upload: bad arguments");
  h2d(c, dptr, host, (size_t)n);
  sync(c);
  TNAD_API_END(c)
}
int tnad_dev_download(tnad_ctx* c, double* host, const double* dptr, int64_t n) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(dptr && host && n >= 0, "tnad_dev_download: bad arguments");
  TNAD_CUDA(cudaMemcpyAsync(host, dptr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  TNAD_API_END(c)
}

int tnad_last_timing(tnad_ctx* c, double* ms) {
  if (!c || !ms) return TNAD_ERR_ARG;
  for (int i = 0; i < 8; ++i) ms[i] = c->timing[i];
  return TNAD_OK;
}

// ---- building blocks --------------------------------------------------------------------------------
int tnad_contract(tnad_ctx* c, const char* spec, const double* A, const int64_t* dimsA, int rankA, const double* B,
                  const int64_t* dimsB, int rankB, double alpha, double beta, double* C) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(spec && dimsA && dimsB && rankA >= 0 && rankA <= MAXR && rankB >= 0 && rankB <= MAXR,
               "tnad_contract: bad arguments");
  std::vector<int64_t> da(dimsA, dimsA + rankA), db(dimsB, dimsB + rankB);
  Tens tA = t_in(c, A, da), tB = t_in(c, B, db);
  std::string s(spec);
  size_t arrow = s.find("->"), comma = s.find(',');
  TNAD_REQUIRE(arrow != std::string::npos && comma != std::string::npos, "tnad_contract: bad spec");
  std::string la, lb, lc;
  for (char ch : s.substr(0, comma)) if (ch != ' ') la += ch;
  for (char ch : s.substr(comma + 1, arrow - comma - 1)) if (ch != ' ') lb += ch;
  for (char ch : s.substr(arrow + 2)) if (ch != ' ') lc += ch;
  std::vector<int64_t> dc;
  for (char ch : lc) {
    size_t ia = la.find(ch), ib = lb.find(ch);
    TNAD_REQUIRE(ia != std::string::npos || ib != std::string::npos, "tnad_contract: output label not in inputs");
    dc.push_back(ia != std::string::npos ? da[ia] : db[ib]);
  }
  Tens tC;
  if (beta != 0.0) tC = t_in(c, C, dc);
  else tC = (c->pointer_mode == TNAD_POINTER_DEVICE) ? t_wrap(C, dc) : t_alloc_v(c, dc);
  const int reps = beta == 0.0 ? std::max(1, opt_i(c, "TNAD_CONTRACT_REPS", 1)) : 1;   // measurement aid: the same product enqueued reps times
  for (int r = 0; r < reps; ++r) contract(c, spec, tA, tB, tC, alpha, beta);
  if (c->pointer_mode == TNAD_POINTER_DEVICE) sync(c);
  else t_out(c, tC, C);
  TNAD_API_END(c)
}

int tnad_svd(tnad_ctx* c, const double* A, int m, int n, double* U, double* S, double* V, int* sweeps_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(m >= 1 && n >= 1, "tnad_svd: empty matrix");
  Tens tA = t_in(c, A, {m, n});
  SvdResult r = svd_jacobi(c, tA, false);
  t_out(c, r.U, U);
  t_out(c, r.S, S);
  t_out(c, r.V, V);
  if (sweeps_out) *sweeps_out = r.sweeps;
  TNAD_API_END(c)
}

int tnad_svd_sym(tnad_ctx* c, const double* A, int n, double* U, double* S, double* V, int* sweeps_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(n >= 1, "tnad_svd_sym: empty matrix");
  Tens tA = t_in(c, A, {n, n});
  SvdResult r = svd_symmetric_auto(c, tA, false);
  t_out(c, r.U, U);
  t_out(c, r.S, S);
  t_out(c, r.V, V);
  if (sweeps_out) *sweeps_out = r.sweeps;
  TNAD_API_END(c)
}

int tnad_svd_symmetrized(tnad_ctx* c, const double* A, int n, double* U, double* S, double* V, int* sweeps_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(n >= 1, "tnad_svd_symmetrized: empty matrix");
  Tens tA = t_in(c, A, {n, n});
  SvdResult r = svd_symmetric_auto(c, tA, true);
  t_out(c, r.U, U);
  t_out(c, r.S, S);
  t_out(c, r.V, V);
  if (sweeps_out) *sweeps_out = r.sweeps;
  TNAD_API_END(c)
}

// ---- the symmetric eigensolver in three phases (the back-transformation is independent per column: GPUs share it) ----
struct tnad_eig {
  tnad_ctx* ctx = nullptr;
  tnad::EigFactor f;
};

int tnad_symeig_reduce(tnad_ctx* c, const double* A, int n, int add_transpose, tnad_eig** out, int* N_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(A && n >= 1 && out && N_out, "tnad_symeig_reduce: bad arguments");
  TNAD_REQUIRE(c->coop_launch, "tnad_symeig_reduce: needs cooperative kernel launches");
  Tens tA = t_in(c, A, {n, n});
  Tens Aw;
  load_symmetric(c, tA, add_transpose != 0, Aw);
  tnad_eig* h = new tnad_eig();
  h->ctx = c;
  try {
    symeig_reduce(c, Aw, n, h->f);
  } catch (...) {
    delete h;
    throw;
  }
  sync(c);
  *out = h;
  *N_out = (int)h->f.N;
  TNAD_API_END(c)
}

int tnad_symeig_backtransform(tnad_ctx* c, tnad_eig* h, int col0, int ncols, double* Zcols) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(h && h->ctx == c && Zcols && col0 >= 0 && ncols >= 0 && col0 + ncols <= h->f.N, "tnad_symeig_backtransform: bad arguments");
  const int64_t N = h->f.N;
  if (c->pointer_mode == TNAD_POINTER_DEVICE) {
    TNAD_CUDA(cudaMemcpyAsync(Zcols, h->f.Z.p + (int64_t)col0 * N, (size_t)N * ncols * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    symeig_backtransform(c, h->f, Zcols, N, ncols);
    sync(c);
  } else {
    Tens tmp = t_alloc(c, {N, (int64_t)ncols});
    TNAD_CUDA(cudaMemcpyAsync(tmp.p, h->f.Z.p + (int64_t)col0 * N, (size_t)N * ncols * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    symeig_backtransform(c, h->f, tmp.p, N, ncols);
    t_out(c, tmp, Zcols);
  }
  TNAD_API_END(c)
}

int tnad_symeig_finish(tnad_ctx* c, tnad_eig* h, const double* Zfull, double* U, double* S, double* V) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(h && h->ctx == c && Zfull && U && S && V, "tnad_symeig_finish: bad arguments");
  Tens tZ = t_in(c, Zfull, {h->f.N, h->f.N});
  SvdResult r = symeig_finish(c, h->f, tZ.p);
  t_out(c, r.U, U);
  t_out(c, r.S, S);
  t_out(c, r.V, V);
  TNAD_API_END(c)
}

int tnad_symeig_free(tnad_eig* h) {
  if (!h) return TNAD_OK;
  if (h->ctx) {
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
  }
  delete h;
  return TNAD_OK;
}

int tnad_sytrd(tnad_ctx* c, const double* A, int n, double* d, double* e, double* Q) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(n >= 1 && d && e && Q, "tnad_sytrd: bad arguments");
  Tens tA = t_in(c, A, {n, n});
  Tens Aw = t_clone(c, tA);
  Tens Vh = t_alloc(c, {n, sytrd_vcols(n)}, true), tau = t_alloc(c, {sytrd_vcols(n)}, true), dd = t_alloc(c, {n}), ee = t_alloc(c, {n}, true);
  sytrd(c, Aw.p, n, n, Vh.p, n, tau.p, dd.p, ee.p);
  Tens Qm = t_alloc(c, {n, n}, true);
  set_identity(c, Qm.p, n, n);
  apply_q(c, Vh.p, n, tau.p, n, Qm.p, n, n);
  t_out(c, dd, d);
  if (n > 1) {
    Tens ev = t_wrap(ee.p, {n - 1});
    t_out(c, ev, e);
  }
  t_out(c, Qm, Q);
  TNAD_API_END(c)
}

int tnad_sytrd2(tnad_ctx* c, const double* A, int n, double* d, double* e, double* Q, double* band) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(n >= 3 && d && e && Q, "tnad_sytrd2: bad arguments (n >= 3)");
  Tens tA = t_in(c, A, {n, n});
  Tens Aw = t_clone(c, tA);
  const int64_t vc = sytrd_vcols(n), NP = chase_positions(n), ldv2 = 32 * NP;
  Tens Yst = t_alloc(c, {n, vc}, true), tau1 = t_alloc(c, {vc}, true), dd = t_alloc(c, {n}), ee = t_alloc(c, {n}, true);
  sy2sb(c, Aw.p, n, n, Yst.p, n, tau1.p);
  Tens AB = t_alloc(c, {33, n});
  extract_band(c, Aw.p, n, n, AB.p, 33);
  Tens V2 = t_alloc(c, {ldv2, (int64_t)n - 2}, true), tau2 = t_alloc(c, {NP, (int64_t)n - 2}, true);
  sb2st(c, AB.p, 33, n, dd.p, ee.p, V2.p, ldv2, tau2.p);
  Tens Qm = t_alloc(c, {n, n}, true);
  set_identity(c, Qm.p, n, n);
  apply_q2(c, V2.p, ldv2, tau2.p, n, Qm.p, n, n);
  apply_q(c, Yst.p, n, tau1.p, n, Qm.p, n, n, 32);
  t_out(c, dd, d);
  Tens ev = t_wrap(ee.p, {n - 1});
  t_out(c, ev, e);
  t_out(c, Qm, Q);
  if (band) t_out(c, AB, band);
  TNAD_API_END(c)
}

int tnad_stedc(tnad_ctx* c, const double* d, const double* e, int n, double* lam, double* Z) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(n >= 1 && d && lam && Z && (n == 1 || e), "tnad_stedc: bad arguments");
  Tens td = t_in(c, d, {n});
  Tens te = t_alloc(c, {n}, true);
  if (n > 1) {
    Tens tmp = t_in(c, e, {n - 1});
    TNAD_CUDA(cudaMemcpyAsync(te.p, tmp.p, (size_t)(n - 1) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  Tens l, Zp;
  int64_t N = 0;
  stedc(c, td.p, te.p, n, l, Zp, N);
  std::vector<double> lh((size_t)N);
  d2h(c, lh.data(), l.p, (size_t)N);
  std::vector<int> idx((size_t)N);
  for (int64_t i = 0; i < N; ++i) idx[(size_t)i] = (int)i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lh[x] < lh[y]; });
  Tens lo = t_alloc(c, {n}), Zo = t_alloc(c, {n, n});
  std::vector<double> ls((size_t)n);
  for (int j = 0; j < n; ++j) {
    ls[(size_t)j] = lh[idx[(size_t)j]];
    TNAD_CUDA(cudaMemcpyAsync(Zo.p + (int64_t)j * n, Zp.p + (int64_t)idx[(size_t)j] * N, (size_t)n * sizeof(double),
                              cudaMemcpyDeviceToDevice, c->stream));
  }
  h2d(c, lo.p, ls.data(), (size_t)n);
  t_out(c, lo, lam);
  t_out(c, Zo, Z);
  TNAD_API_END(c)
}

int tnad_trg_svd(tnad_ctx* c, const double* t, int d1, int d2, int d3, int d4, int dmax, double tol, double* u,
                 double* v, int* k_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(d1 >= 1 && d2 >= 1 && d3 >= 1 && d4 >= 1 && dmax >= 1 && k_out, "tnad_trg_svd: bad arguments");
  Tens tt = t_in(c, t, {d1, d2, d3, d4});
  TrgSplit sp = trg_split(c, tt, dmax, tol);
  *k_out = (int)sp.k;
  t_out(c, sp.us, u);
  Tens vt = t_clone(c, t_perm(sp.vs, {2, 0, 1}));   // (k, d3, d4) as the reference returns it
  t_out(c, vt, v);
  TNAD_API_END(c)
}

int tnad_svd_back(tnad_ctx* c, int m, int n, int k, const double* U, const double* S, const double* V,
                  const double* dU, const double* dS, const double* dV, double eta, double* dA) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(m >= 1 && n >= 1 && k == std::min(m, n), "tnad_svd_back: k must equal min(m, n)");
  TNAD_REQUIRE(dU || dS || dV, "tnad_svd_back: all cotangents are nothing");
  Tens tU = t_in(c, U, {m, k}), tS = t_in(c, S, {k}), tV = t_in(c, V, {n, k});
  Tens tdU, tdS, tdV;
  if (dU) tdU = t_in(c, dU, {m, k});
  if (dS) tdS = t_in(c, dS, {k});
  if (dV) tdV = t_in(c, dV, {n, k});
  Tens r = svd_back_dev(c, tU, tS, tV, dU ? &tdU : nullptr, dS ? &tdS : nullptr, dV ? &tdV : nullptr, k, eta);
  t_out(c, r, dA);
  TNAD_API_END(c)
}

// ---- TRG ----------------------------------------------------------------------------------------------
int tnad_trg_forward(tnad_ctx* c, const double* a, int d0, int d1, int chi, int niter, double tol, double* lnZ,
                     tnad_tape** tape) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(lnZ && d0 >= 1 && d1 >= 1, "tnad_trg_forward: bad arguments");
  timing_begin(c);
  Tens ta = t_in(c, a, {d0, d1, d0, d1});
  tnad_tape* tp = nullptr;
  if (tape) {
    tp = new tnad_tape();
    tp->ctx = c;
    tp->kind = 1;
    c->live_tapes++;
  }
  try {
    Span sp(c, 0);
    *lnZ = trg_forward(c, ta, chi, niter, tol, tp ? &tp->trg : nullptr);
  } catch (...) {
    if (tp) c->live_tapes--;
    delete tp;
    throw;
  }
  timing_end(c);
  if (tape) *tape = tp;
  TNAD_API_END(c)
}

int tnad_trg_backward(tnad_ctx* c, tnad_tape* tape, double dlnZ, double* da) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(tape && tape->kind == 1 && tape->ctx == c && da, "tnad_trg_backward: not a TRG tape of this context");
  timing_begin(c);
  Tens g;
  {
    Span sp(c, 0);
    g = trg_backward(c, tape->trg, dlnZ);
  }
  t_out(c, g, da);
  timing_end(c);
  TNAD_API_END(c)
}

int tnad_tape_free(tnad_tape* tape) {
  if (!tape) return TNAD_OK;
  tnad_ctx* c = tape->ctx;
  if (c) cudaSetDevice(c->device);
  delete tape;
  if (c) c->live_tapes--;
  return TNAD_OK;
}

// ---- CTMRG --------------------------------------------------------------------------------------------
int tnad_ctmrg_init_raw(tnad_ctx* c, const double* bulk, int D, int chi, double* corner, double* edge) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1, "tnad_ctmrg_init_raw: bad arguments");
  Tens b = t_in(c, bulk, {D, D, D, D});
  Tens co = t_alloc(c, {chi, chi}), ed = t_alloc(c, {chi, D, chi});
  init_raw(c, b, co, ed);
  t_out(c, co, corner);
  t_out(c, ed, edge);
  TNAD_API_END(c)
}

int tnad_ctmrg_init_random(tnad_ctx* c, int D, int chi, unsigned long long seed, double* corner, double* edge) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1 && corner && edge, "tnad_ctmrg_init_random: bad arguments");
  Tens co, ed;
  init_random(c, D, chi, seed, co, ed);
  t_out(c, co, corner);
  t_out(c, ed, edge);
  TNAD_API_END(c)
}

int tnad_ctmrgstep(tnad_ctx* c, const double* bulk, int D, int chi, const double* corner_in, const double* edge_in,
                   double* corner_out, double* edge_out, double* vals) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1, "tnad_ctmrgstep: bad arguments");
  Tens b = t_in(c, bulk, {D, D, D, D}), co = t_in(c, corner_in, {chi, chi}), ed = t_in(c, edge_in, {chi, D, chi});
  Tens cn, en;
  std::vector<double> v;
  ctmrg_step(c, b, co, ed, cn, en, v, nullptr);
  t_out(c, cn, corner_out);
  t_out(c, en, edge_out);
  if (vals) memcpy(vals, v.data(), v.size() * sizeof(double));
  TNAD_API_END(c)
}

int tnad_ctmrgstep_backward(tnad_ctx* c, const double* bulk, int D, int chi, const double* corner_in,
                            const double* edge_in, const double* dcorner_out, const double* dedge_out, double* dbulk,
                            double* dcorner_in, double* dedge_in) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1 && dcorner_out && dedge_out && dbulk, "tnad_ctmrgstep_backward: bad arguments");
  Tens b = t_in(c, bulk, {D, D, D, D}), co = t_in(c, corner_in, {chi, chi}), ed = t_in(c, edge_in, {chi, D, chi});
  Tens cb3 = t_in(c, dcorner_out, {chi, chi}), eb3 = t_in(c, dedge_out, {chi, D, chi});
  Tens cn, en, cb, eb;
  std::vector<double> v;
  CtmrgStepRec rec;
  ctmrg_step(c, b, co, ed, cn, en, v, &rec);   // forward with a record (the pullback of one step is self-contained)
  Tens bb = t_alloc(c, {D, D, D, D}, true);
  ctmrg_step_backward(c, b, rec, cb3, eb3, bb, cb, eb, 1e-40);
  t_out(c, bb, dbulk);
  if (dcorner_in) t_out(c, cb, dcorner_in);
  if (dedge_in) t_out(c, eb, dedge_in);
  TNAD_API_END(c)
}

int tnad_ctmrg(tnad_ctx* c, const double* bulk, int D, int chi, double* corner, double* edge, double tol, int maxit,
               int* steps_done, double* vals, tnad_tape** tape) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1 && maxit >= 0, "tnad_ctmrg: bad arguments");
  timing_begin(c);
  Tens b = t_in(c, bulk, {D, D, D, D});
  // the environment is in/out: work on private copies so a device-mode caller's buffers are only written at the end
  Tens co = t_clone(c, t_in(c, corner, {chi, chi})), ed = t_clone(c, t_in(c, edge, {chi, D, chi}));
  if (c->pointer_mode == TNAD_POINTER_DEVICE) b = t_clone(c, b);
  tnad_tape* tp = nullptr;
  if (tape) {
    tp = new tnad_tape();
    tp->ctx = c;
    tp->kind = 2;
    c->live_tapes++;
  }
  std::vector<double> v;
  int ns = 0;
  try {
    Span sp(c, 0);
    ns = ctmrg_loop(c, b, co, ed, tol, maxit, v, tp ? &tp->ctmrg : nullptr);
  } catch (...) {
    if (tp) c->live_tapes--;
    delete tp;
    throw;
  }
  t_out(c, co, corner);
  t_out(c, ed, edge);
  if (steps_done) *steps_done = ns;
  if (vals) memcpy(vals, v.data(), v.size() * sizeof(double));
  timing_end(c);
  if (tape) *tape = tp;
  TNAD_API_END(c)
}

int tnad_ctmrg_backward(tnad_ctx* c, tnad_tape* tape, const double* dcorner, const double* dedge, double* dbulk,
                        double* dcorner0, double* dedge0) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(tape && tape->kind == 2 && tape->ctx == c && dbulk, "tnad_ctmrg_backward: not a CTMRG tape of this context");
  timing_begin(c);
  const int64_t D = tape->ctmrg.D, chi = tape->ctmrg.chi;
  Tens cb = t_in(c, dcorner, {chi, chi}), eb = t_in(c, dedge, {chi, D, chi});
  Tens bb, cb0, eb0;
  {
    Span sp(c, 0);
    Span sp3(c, 3);
    ctmrg_backward(c, tape->ctmrg, cb, eb, bb, cb0, eb0, 1e-40);
  }
  t_out(c, bb, dbulk);
  if (dcorner0) t_out(c, cb0, dcorner0);
  if (dedge0) t_out(c, eb0, dedge0);
  timing_end(c);
  TNAD_API_END(c)
}

int tnad_permute(tnad_ctx* c, const double* in, const int64_t* dims, int rank, const int* perm, double* out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(in && dims && perm && out && rank >= 1 && rank <= MAXR, "tnad_permute: bad arguments");
  std::vector<int64_t> din(dims, dims + rank), dout((size_t)rank);
  Tens tin = t_in(c, in, din);
  Tens view = tin;
  for (int i = 0; i < rank; ++i) {
    TNAD_REQUIRE(perm[i] >= 0 && perm[i] < rank, "tnad_permute: bad permutation");
    view.dim[i] = tin.dim[perm[i]];
    view.str[i] = tin.str[perm[i]];
    dout[(size_t)i] = tin.dim[perm[i]];
  }
  Tens tout = (c->pointer_mode == TNAD_POINTER_DEVICE) ? t_wrap(out, dout) : t_alloc_v(c, dout);
  tcopy(c, view, tout, 1.0, 0.0);
  if (c->pointer_mode == TNAD_POINTER_DEVICE) sync(c);
  else t_out(c, tout, out);
  TNAD_API_END(c)
}

int tnad_ctmrg_finish(tnad_ctx* c, const double* c1p, const double* e1p, int D, int chi, double* corner_out,
                      double* edge_out) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1, "tnad_ctmrg_finish: bad arguments");
  Tens c1 = t_in(c, c1p, {chi, chi}), e1 = t_in(c, e1p, {chi, D, chi});
  Tens c2 = t_clone(c, c1), e2 = t_clone(c, e1);
  tcopy(c, t_perm(c1, {1, 0}), c2, 1.0, 1.0);
  tcopy(c, t_perm(e1, {2, 1, 0}), e2, 1.0, 1.0);
  Tens ss = t_alloc(c, {2});
  reduce(c, RED_SUMSQ, c2, nullptr, ss.p);
  reduce(c, RED_SUMSQ, e2, nullptr, ss.p + 1);
  Tens co = t_alloc(c, {chi, chi}), eo = t_alloc(c, {chi, D, chi});
  scale_dev(c, c2, co, ss.p, SC_INVSQRT);
  scale_dev(c, e2, eo, ss.p + 1, SC_INVSQRT);
  t_out(c, co, corner_out);
  t_out(c, eo, edge_out);
  TNAD_API_END(c)
}

// ---- energy -------------------------------------------------------------------------------------------
int tnad_expectationvalue(tnad_ctx* c, const double* h, const double* ap, int D, int s, const double* corner,
                          const double* edge, int chi, double* e) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(e && D >= 1 && s >= 1 && chi >= 1, "tnad_expectationvalue: bad arguments");
  Tens th = t_in(c, h, {s, s, s, s}), tap = t_in(c, ap, {D, D, D, D, s, s});
  Tens co = t_in(c, corner, {chi, chi}), ed = t_in(c, edge, {chi, D, chi});
  *e = expectationvalue(c, th, tap, co, ed, nullptr);
  TNAD_API_END(c)
}

int tnad_expectationvalue_backward(tnad_ctx* c, const double* h, const double* ap, int D, int s, const double* corner,
                                   const double* edge, int chi, double ybar, double* dap, double* dcorner,
                                   double* dedge) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && s >= 1 && chi >= 1, "tnad_expectationvalue_backward: bad arguments");
  Tens th = t_in(c, h, {s, s, s, s}), tap = t_in(c, ap, {D, D, D, D, s, s});
  Tens co = t_in(c, corner, {chi, chi}), ed = t_in(c, edge, {chi, D, chi});
  ExpvalTape et;
  expectationvalue(c, th, tap, co, ed, &et);
  Tens apbar, cbar, ebar;
  expectationvalue_back(c, co, ed, et, ybar, apbar, cbar, ebar);
  if (dap) t_out(c, apbar, dap);
  if (dcorner) t_out(c, cbar, dcorner);
  if (dedge) t_out(c, ebar, dedge);
  TNAD_API_END(c)
}

int tnad_energy(tnad_ctx* c, const double* h, const double* A, int d, int s, int chi, double tol, int maxit,
                double* e, double* gradA, int* steps_done) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(e && d >= 1 && s >= 1, "tnad_energy: bad arguments");
  timing_begin(c);
  Tens th = t_in(c, h, {s, s, s, s}), tA = t_in(c, A, {d, d, d, d, s});
  Tens g;
  {
    Span sp(c, 0);
    *e = energy(c, th, tA, chi, tol, maxit, gradA ? &g : nullptr, steps_done);
  }
  if (gradA) t_out(c, g, gradA);
  timing_end(c);
  TNAD_API_END(c)
}

int tnad_energy_fixedpoint(tnad_ctx* c, const double* h, const double* A, int d, int s, int chi, double tol, int maxit,
                           double bwd_tol, int bwd_maxit, double* e, double* gradA, int* steps_done, int* bwd_iters) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(e && d >= 1 && s >= 1, "tnad_energy_fixedpoint: bad arguments");
  timing_begin(c);
  Tens th = t_in(c, h, {s, s, s, s}), tA = t_in(c, A, {d, d, d, d, s});
  Tens g;
  {
    Span sp(c, 0);
    *e = energy_fixedpoint(c, th, tA, chi, tol, maxit, bwd_tol, bwd_maxit, gradA ? &g : nullptr, steps_done, bwd_iters);
  }
  if (gradA) t_out(c, g, gradA);
  timing_end(c);
  TNAD_API_END(c)
}

int tnad_magnetisation_readout(tnad_ctx* c, const double* a, const double* m, int D, const double* corner,
                               const double* edge, int chi, double* mag) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(mag && D >= 1 && chi >= 1, "tnad_magnetisation_readout: bad arguments");
  Tens ta = t_in(c, a, {D, D, D, D}), tm = t_in(c, m, {D, D, D, D});
  Tens co = t_in(c, corner, {chi, chi}), ed = t_in(c, edge, {chi, D, chi});
  *mag = magnetisation_readout(c, ta, tm, co, ed);
  TNAD_API_END(c)
}

int tnad_magnetisation_backward(tnad_ctx* c, const double* a, const double* m, int D, const double* corner,
                                const double* edge, int chi, double ybar, double* da, double* dm, double* dcorner,
                                double* dedge) {
  TNAD_API_BEGIN(c)
  TNAD_REQUIRE(D >= 1 && chi >= 1, "tnad_magnetisation_backward: bad arguments");
  Tens ta = t_in(c, a, {D, D, D, D}), tm = t_in(c, m, {D, D, D, D});
  Tens co = t_in(c, corner, {chi, chi}), ed = t_in(c, edge, {chi, D, chi});
  MagTape mt;
  magnetisation_readout(c, ta, tm, co, ed, &mt);
  Tens ab, mb, cb, eb;
  magnetisation_readout_back(c, ta, tm, co, ed, mt, ybar, ab, mb, cb, eb);
  if (da) t_out(c, ab, da);
  if (dm) t_out(c, mb, dm);
  if (dcorner) t_out(c, cb, dcorner);
  if (dedge) t_out(c, eb, dedge);
  TNAD_API_END(c)
}

// ---- multi-GPU: independent instances (beta sweeps / parameter scans, no communication) -----------------
int tnad_trg_sweep(const double* tensors, int ninst, int d0, int d1, int chi, int niter, double tol, int ngpu,
                   const int* devices, double* lnZ, double* grads, char* errbuf, int errlen) {
  if (!tensors || !lnZ || ninst < 0 || d0 < 1 || d1 < 1 || ngpu < 1) return TNAD_ERR_ARG;
  const int64_t numel = (int64_t)d0 * d1 * d0 * d1;
  std::vector<int> rc((size_t)ngpu, TNAD_OK);
  std::vector<std::string> msg((size_t)ngpu);
  auto worker = [&](int g) {
    tnad_ctx* ctx = nullptr;
    int r = tnad_create(devices ? devices[g] : g, &ctx);
    if (r != TNAD_OK) {
      rc[(size_t)g] = r;
      msg[(size_t)g] = tnad_last_error(nullptr);
      return;
    }
    for (int i = g; i < ninst && r == TNAD_OK; i += ngpu) {   // instance i -> device i mod ngpu
      tnad_tape* tape = nullptr;
      r = tnad_trg_forward(ctx, tensors + (int64_t)i * numel, d0, d1, chi, niter, tol, lnZ + i, grads ? &tape : nullptr);
      if (r == TNAD_OK && grads) r = tnad_trg_backward(ctx, tape, 1.0, grads + (int64_t)i * numel);
      tnad_tape_free(tape);
    }
    if (r != TNAD_OK) {
      rc[(size_t)g] = r;
      msg[(size_t)g] = tnad_last_error(ctx);
    }
    tnad_destroy(ctx);
  };
  try {
    std::vector<std::thread> th;
    for (int g = 1; g < ngpu; ++g) th.emplace_back(worker, g);
    worker(0);
    for (auto& t : th) t.join();
  } catch (...) {
    return TNAD_ERR_INTERNAL;
  }
  for (int g = 0; g < ngpu; ++g)
    if (rc[(size_t)g] != TNAD_OK) {
      if (errbuf && errlen > 0) snprintf(errbuf, (size_t)errlen, "device slot %d: %s", g, msg[(size_t)g].c_str());
      return rc[(size_t)g];
    }
  return TNAD_OK;
}

}  // extern "C"
