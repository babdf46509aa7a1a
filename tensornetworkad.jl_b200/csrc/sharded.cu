// chi-sharded ctmrgstep inside the library (`tnad_ctmrg_sharded(ctx, ncclComm_t, ...)`).
//
// One process per GPU; every rank holds the full environment (corner chi x chi, edge chi x D x chi: a few MB), computes
// a 1 / world slice of every O(chi^3 D^3) / O(chi^3 D^4) contraction of ctmrg.jl:126-142 on the SLOWEST index of the
// result -- so the gathered buffer is the full tensor without a re-layout -- and the slices are concatenated with
// ncclAllGather on the context's stream.  The eigen-decomposition of the enlarged corner (ctmrg.jl:134-136): the
// latency-bound reduction stages run replicated, every rank back-transforms ITS block of N / world eigenvector columns
// (the 4 n^3 flop) and the blocks are gathered in place.  Same arithmetic as the single-GPU step.
//
// NCCL is not linked: the library is loaded at run time (`tnad_comm_init` takes the path, or finds the copy the host
// process already mapped, e.g. torch's), the communicator is created from a 128-byte unique id that rank 0 obtains with
// `tnad_nccl_unique_id` and the host application distributes (MPI / torch.distributed / a file).
#include "drivers.h"
#include "eigdc.h"
#include <dlfcn.h>
#include <mutex>

namespace tnad {

namespace {

struct NcclId {
  char internal[128];
};
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_get_error_string get_error_string = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
constexpr int NCCL_FLOAT64 = 8;   // ncclDataType_t: ncclFloat64 / ncclDouble

void nccl_load(const char* path) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return;
  void* h = nullptr;
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);       // the copy the process already mapped, or the system one
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) fail(TNAD_ERR_ARG, std::string("tnad_comm: cannot load NCCL (") + (dlerror() ? dlerror() : "no error text") + ")");
  NcclApi a;
  a.lib = h;
  a.get_unique_id = reinterpret_cast<fn_get_unique_id>(dlsym(h, "ncclGetUniqueId"));
  a.comm_init_rank = reinterpret_cast<fn_comm_init_rank>(dlsym(h, "ncclCommInitRank"));
  a.all_gather = reinterpret_cast<fn_all_gather>(dlsym(h, "ncclAllGather"));
  a.comm_destroy = reinterpret_cast<fn_comm_destroy>(dlsym(h, "ncclCommDestroy"));
  a.get_error_string = reinterpret_cast<fn_get_error_string>(dlsym(h, "ncclGetErrorString"));
  if (!a.get_unique_id || !a.comm_init_rank || !a.all_gather || !a.comm_destroy)
    fail(TNAD_ERR_ARG, "tnad_comm: the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclCommDestroy");
  g_nccl = a;
}

void nccl_check(int rc, const char* what) {
  if (rc != 0)
    fail(TNAD_ERR_CUDA, std::string(what) + ": " + (g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "NCCL error") + " (" +
                            std::to_string(rc) + ")");
}

// recv = concatenation over the ranks of `count` doubles; in place when send == recv + rank * count
void all_gather(tnad_ctx* c, const double* send, double* recv, size_t count) {
  if (c->comm_world <= 1) {
    if (send != recv) TNAD_CUDA(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  nccl_check(g_nccl.all_gather(send, recv, count, NCCL_FLOAT64, c->comm, c->stream), "ncclAllGather");
}

}  // namespace

// One ctmrgstep with the contractions and the back-transformation of the eigen-decomposition shared between the ranks of
// the context's communicator.  ms (optional, 3 doubles): device time of contractions (+ finish), all-gathers, decomposition.
void ctmrg_step_sharded(tnad_ctx* c, const Tens& bulk, const Tens& corner, const Tens& edge, Tens& corner_out, Tens& edge_out,
                        std::vector<double>& vals_host, double* ms, CtmrgStepRec* rec) {
  const int world = c->comm_world > 1 ? c->comm_world : 1, rank = c->comm_world > 1 ? c->comm_rank : 0;
  const int64_t D = bulk.dim[0], chi = corner.dim[0], n = chi * D;
  TNAD_REQUIRE(chi % world == 0, "tnad_ctmrgstep_sharded: chi must be divisible by the number of ranks");
  const int64_t w = chi / world, k0 = rank * w;
  cudaStream_t st = c->stream;
  cudaEvent_t ev[8];
  int nev = 0;
  auto mark = [&]() {
    if (!ms) return;
    ev[nev] = get_event(c);
    TNAD_CUDA(cudaEventRecord(ev[nev], st));
    ++nev;
  };
  mark();                                                                       // 0
  // grow: cp[i,j,l,k] = corner[a,d] edge[i,b,a] edge[d,c,l] bulk[j,k,c,b] on this rank's slice of l (ctmrg.jl:130)
  Tens X1 = contract_new(c, "iba,ad->ibd", edge, corner);
  Tens edge_sl = t_slice_last(edge, k0, w);
  Tens X2r = contract_new(c, "ibd,dcl->ibcl", X1, edge_sl);
  Tens cpk = t_alloc(c, {chi, D, D, chi});                                      // [i, j, k, l]: l slowest, so the slices concatenate
  Tens cpk_r = t_slice_last(cpk, k0, w);
  contract(c, "ibcl,jkcb->ijkl", X2r, bulk, cpk_r);
  mark();                                                                       // 1
  all_gather(c, cpk_r.p, cpk.p, (size_t)(chi * D * D * w));
  mark();                                                                       // 2
  Tens X2full;
  if (rec) {   // the reverse sweep wants the full X2[i,b,c,l]: the slices concatenate along l
    X2full = t_alloc(c, {chi, D, D, chi});
    all_gather(c, X2r.p, X2full.p, (size_t)(chi * D * D * w));
  }
  Tens cp = t_alloc(c, {chi, D, chi, D});
  tcopy(c, t_perm(cpk, {0, 1, 3, 2}), cp);
  Tens CP = t_reshape(cp, {n, n});
  mark();                                                                       // 3
  // svd(cp + cp') (ctmrg.jl:134-136): replicated reduction, shared back-transformation
  Tens Aw;
  load_symmetric(c, CP, true, Aw);
  EigFactor f;
  symeig_reduce(c, Aw, n, f, world == 1);
  SvdResult svd;
  if (world == 1) {
    if (f.Qfull.p) {
      TNAD_CUDA(cudaStreamWaitEvent(st, f.q_ready, 0));
      Tens X = t_wrap(f.Z.p, {f.n, f.N});
      X.str[1] = f.N;
      Tens U0 = contract_new(c, "ik,kj->ij", f.Qfull, X);
      svd = symeig_finish(c, f, U0.p, f.n);
      c->event_pool.push_back(f.q_ready);
    } else {
      symeig_backtransform(c, f, f.Z.p, f.N, f.N);
      svd = symeig_finish(c, f, f.Z.p);
    }
  } else if (f.N % world == 0) {
    const int64_t wc = f.N / world;
    double* mine = f.Z.p + (int64_t)rank * wc * f.N;
    symeig_backtransform(c, f, mine, f.N, wc);
    all_gather(c, mine, f.Z.p, (size_t)(wc * f.N));                             // in place
    svd = symeig_finish(c, f, f.Z.p);
  } else {
    symeig_backtransform(c, f, f.Z.p, f.N, f.N);
    svd = symeig_finish(c, f, f.Z.p);
  }
  mark();                                                                       // 4
  // corner = z' cp z and edge = z' (edge * bulk) z on this rank's slice of the kept columns (ctmrg.jl:137-140)
  Tens Zall = t_slice_last(svd.U, 0, chi), Zs = t_slice_last(svd.U, k0, w);
  Tens Wr = contract_new(c, "pq,qj->pj", CP, Zs);
  Tens c1 = t_alloc(c, {chi, chi}), e1 = t_alloc(c, {chi, D, chi});
  Tens c1_r = t_slice_last(c1, k0, w), e1_r = t_slice_last(e1, k0, w);
  contract(c, "pi,pj->ij", Zall, Wr, c1_r);
  Tens zall = t_reshape(Zall, {chi, D, chi}), zs = t_reshape(Zs, {chi, D, w});
  Tens Yr = contract_new(c, "aed,dck->aeck", edge, zs);
  Tens Yb = contract_new(c, "aeck,bjce->abjk", Yr, bulk);
  contract(c, "abi,abjk->ijk", zall, Yb, e1_r);
  mark();                                                                       // 5
  all_gather(c, c1_r.p, c1.p, (size_t)(chi * w));
  all_gather(c, e1_r.p, e1.p, (size_t)(chi * D * w));
  mark();                                                                       // 6
  // symmetrise (ctmrg.jl:145-146) and normalise (ctmrg.jl:149-150)
  Tens c2 = t_clone(c, c1), e2 = t_clone(c, e1);
  tcopy(c, t_perm(c1, {1, 0}), c2, 1.0, 1.0);
  tcopy(c, t_perm(e1, {2, 1, 0}), e2, 1.0, 1.0);
  Tens ss = t_alloc(c, {2});
  reduce(c, RED_SUMSQ, c2, nullptr, ss.p);
  reduce(c, RED_SUMSQ, e2, nullptr, ss.p + 1);
  corner_out = t_alloc(c, {chi, chi});
  edge_out = t_alloc(c, {chi, D, chi});
  scale_dev(c, c2, corner_out, ss.p, SC_INVSQRT);
  scale_dev(c, e2, edge_out, ss.p + 1, SC_INVSQRT);
  mark();                                                                       // 7
  vals_host.resize((size_t)n);
  const double s0 = svd.s_host[0];
  for (int64_t i = 0; i < n; ++i) vals_host[i] = svd.s_host[i] / s0;            // ctmrg.jl:142
  if (rec) {
    // record for ctmrg_step_backward (replicated on every rank): it uses the intermediates of the unsharded association
    // Y1 = z edge, Y2 = Y1 bulk, recomputed here in full (2 of the step's 8 products; the decomposition dominates)
    rec->corner = corner;
    rec->edge = edge;
    rec->X1 = X1;
    rec->X2 = X2full;
    rec->cp = cp;
    rec->svd = svd;
    rec->Y1 = contract_new(c, "abi,aed->ibed", zall, edge);
    rec->Y2 = contract_new(c, "ibed,bjce->ijcd", rec->Y1, bulk);
    rec->c2 = c2;
    rec->e2 = e2;
    rec->ss = ss;
  }
  if (ms) {
    TNAD_CUDA(cudaEventSynchronize(ev[7]));
    float t[7];
    for (int i = 0; i < 7; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
    if (opt_i(c, "TNAD_SHARD_DEBUG", 0))
      fprintf(stderr, "[tnad sharded] rank %d: grow %.2f  gather %.2f  permute %.2f  eig %.2f  project %.2f  gather %.2f  finish %.2f ms\n", rank,
              t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    ms[0] = t[0] + t[2] + t[4] + t[6];
    ms[1] = t[1] + t[5];
    ms[2] = t[3];
    for (int i = 0; i < 8; ++i) c->event_pool.push_back(ev[i]);
  }
}

}  // namespace tnad

using namespace tnad;

extern "C" {

int tnad_nccl_unique_id(const char* nccl_path, unsigned char* id128) {
  try {
    if (!id128) return TNAD_ERR_ARG;
    nccl_load(nccl_path);
    NcclId id;
    memset(&id, 0, sizeof(id));
    nccl_check(g_nccl.get_unique_id(&id), "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    return e.code;
  } catch (...) {
    return TNAD_ERR_INTERNAL;
  }
}

int tnad_comm_init(tnad_ctx* c, const char* nccl_path, const unsigned char* id128, int rank, int world) {
  if (!c) return TNAD_ERR_ARG;
  try {
    TNAD_CUDA(cudaSetDevice(c->device));
    TNAD_REQUIRE(world >= 1 && rank >= 0 && rank < world, "tnad_comm_init: bad rank / world");
    TNAD_REQUIRE(!c->comm, "tnad_comm_init: the context already has a communicator");
    if (world > 1) {
      TNAD_REQUIRE(id128, "tnad_comm_init: null unique id");
      nccl_load(nccl_path);
      NcclId id;
      memcpy(id.internal, id128, 128);
      void* comm = nullptr;
      nccl_check(g_nccl.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank");
      c->comm = comm;
    }
    c->comm_rank = rank;
    c->comm_world = world;
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    c->err = e.msg;
    cudaGetLastError();
    return e.code;
  } catch (...) {
    c->err = "unknown internal error";
    return TNAD_ERR_INTERNAL;
  }
}

int tnad_comm_destroy(tnad_ctx* c) {
  if (!c) return TNAD_ERR_ARG;
  if (c->comm && g_nccl.comm_destroy) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    g_nccl.comm_destroy(c->comm);
  }
  c->comm = nullptr;
  c->comm_rank = 0;
  c->comm_world = 0;
  return TNAD_OK;
}

// ctmrg(a, chi, tol, maxit) of ctmrg.jl:110-117 / fixedpoint.jl:11-41 over the ranks of the communicator: the stop rule
// (counter from -1, NaN never converged) runs on every rank on the same replicated spectrum, so all ranks stop together
int tnad_ctmrg_sharded(tnad_ctx* c, const double* bulk, int D, int chi, double* corner, double* edge, double tol, int maxit,
                       int* steps_done, double* vals) {
  if (!c) return TNAD_ERR_ARG;
  try {
    TNAD_CUDA(cudaSetDevice(c->device));
    tnad::ApiBracket _bracket(c);
    TNAD_REQUIRE(bulk && corner && edge && D >= 1 && chi >= 1 && maxit >= 0, "tnad_ctmrg_sharded: bad arguments");
    TNAD_REQUIRE(c->coop_launch, "tnad_ctmrg_sharded: needs cooperative kernel launches");
    Tens tb = t_in(c, bulk, {D, D, D, D});
    Tens co = t_clone(c, t_in(c, corner, {chi, chi})), ed = t_clone(c, t_in(c, edge, {chi, D, chi}));
    const size_t n = (size_t)chi * D;
    std::vector<double> v(n, INFINITY), old(n, INFINITY);
    long long counter = -1;   // ctmrg.jl:114
    int ns = 0;
    for (;;) {
      counter += 1;           // fixedpoint.jl:32
      if (counter > maxit) break;
      double ss = 0.0;
      bool isnan = false;
      for (size_t i = 0; i < n; ++i) {
        const double d = v[i] - old[i];
        if (d != d) isnan = true;
        ss += d * d;
      }
      if (!isnan && std::sqrt(ss) <= tol) break;
      old = v;
      Tens cn, en;
      ctmrg_step_sharded(c, tb, co, ed, cn, en, v, nullptr, nullptr);
      co = cn;
      ed = en;
      ++ns;
    }
    t_out(c, co, corner);
    t_out(c, ed, edge);
    if (steps_done) *steps_done = ns;
    if (vals) memcpy(vals, v.data(), n * sizeof(double));
    sync(c);
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    c->err = e.msg;
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    c->err = std::string("internal error: ") + e.what();
    return TNAD_ERR_INTERNAL;
  } catch (...) {
    c->err = "unknown internal error";
    return TNAD_ERR_INTERNAL;
  }
}

int tnad_ctmrgstep_sharded(tnad_ctx* c, const double* bulk, int D, const double* corner, const double* edge, int chi,
                           double* corner_out, double* edge_out, double* vals, double* ms3) {
  if (!c) return TNAD_ERR_ARG;
  try {
    TNAD_CUDA(cudaSetDevice(c->device));
    tnad::ApiBracket _bracket(c);
    TNAD_REQUIRE(bulk && corner && edge && corner_out && edge_out && D >= 1 && chi >= 1, "tnad_ctmrgstep_sharded: bad arguments");
    TNAD_REQUIRE(c->coop_launch, "tnad_ctmrgstep_sharded: needs cooperative kernel launches");
    Tens tb = t_in(c, bulk, {D, D, D, D}), tc = t_in(c, corner, {chi, chi}), te = t_in(c, edge, {chi, D, chi});
    Tens co, eo;
    std::vector<double> v;
    ctmrg_step_sharded(c, tb, tc, te, co, eo, v, ms3, nullptr);
    t_out(c, co, corner_out);
    t_out(c, eo, edge_out);
    if (vals) memcpy(vals, v.data(), v.size() * sizeof(double));   // the spectrum is host data in both pointer modes (as in tnad_ctmrgstep)
    sync(c);
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    c->err = e.msg;
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    c->err = std::string("internal error: ") + e.what();
    return TNAD_ERR_INTERNAL;
  } catch (...) {
    c->err = "unknown internal error";
    return TNAD_ERR_INTERNAL;
  }
}

}  // extern "C"
