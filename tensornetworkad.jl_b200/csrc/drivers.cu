// Drivers: TRG (trg.jl:13-44), CTMRG (ctmrg.jl:110-153, fixedpoint.jl), iPEPS energy
// (variationalipeps.jl:28-56) and the reverse sweeps Zygote derives from them with the rules of
// src/autodiff.jl and trg.jl:55-105.  Every contraction is a pairwise einsum executed by the DMMA
// GEMM with permute-on-load (contract.cu / gemm.cu); the sequences below are the optimal-order,
// `tp`-free association documented in DESIGN.md and mirrored one-to-one in oracle/tnad_oracle.py.
#include "drivers.h"
#include "eigdc.h"
#include <cmath>
#include <cstdlib>

namespace tnad {

// =====================================================================================================
// svd_back
// =====================================================================================================
Tens svd_back_dev(tnad_ctx* c, const Tens& U, const Tens& S, const Tens& V, const Tens* dUk, const Tens* dS,
                  const Tens* dVk, int64_t k, double eta, int64_t r_valid, int64_t r_valid_v) {
  Span sp(c, 4);
  const int64_t m = U.dim[0], n = V.dim[0], kk = S.dim[0];
  TNAD_REQUIRE(U.dim[1] == kk && V.dim[1] == kk && k <= kk, "svd_back: shape mismatch");
  // r_valid < kk: the columns r_valid.. of U (exactly-zero singular values) were not computed; their only
  // contribution, sum_i U_i U_i' dU_j / S_j, is applied as the projector (I - U_r U_r') instead.
  const bool proj = r_valid >= 0 && r_valid < kk && dUk != nullptr;
  // the same for V (Jordan-Wielandt route): its null columns are zero, so V V' = V_r V_r' and the block
  // "U Sinv (dV' - dV'V V')" below applies the whole complement of span(V_r) - null triplets included - exactly once
  const bool projv = r_valid_v >= 0 && r_valid_v < kk && dVk != nullptr;
  const int64_t ru = proj ? r_valid : kk;
  Tens Ur = t_slice_last(U, 0, ru);
  Tens G1, G2;
  if (dUk) {
    if (proj) {
      G1 = t_alloc(c, {kk, k}, true);
      if (ru > 0) {
        Tens G1r = t_wrap(G1.p, {ru, k});
        G1r.str[1] = kk;
        contract(c, "mi,mj->ij", Ur, *dUk, G1r, 1.0, 0.0);
      }
    } else {
      G1 = contract_new(c, "mi,mj->ij", U, *dUk);   // (U' dU)[:, :k]
    }
  }
  if (dVk) G2 = contract_new(c, "ni,nj->ij", V, *dVk);
  Tens Rrow = t_alloc(c, {k, kk});
  Tens Rcol = t_alloc(c, {kk, k});
  svdback_panels(c, kk, k, S.p, dUk ? G1.p : nullptr, dVk ? G2.p : nullptr, dS ? dS->p : nullptr, eta, Rrow.p,
                 Rcol.p);
  Tens Uk = t_slice_last(U, 0, k), Vk = t_slice_last(V, 0, k);
  // dA = U[:, :k] (Rrow V') + (U Rcol) V[:, :k]'
  Tens T1 = contract_new(c, "ij,nj->in", Rrow, V);
  Tens dA = contract_new(c, "mi,in->mn", Uk, T1);
  Tens T2;
  if (proj) {
    Tens Rc = t_wrap(Rcol.p, {ru, k});
    Rc.str[1] = kk;
    T2 = t_clone(c, *dUk);                                    // (dU - U_r U_r'dU) Sinv  +  U_r Rcol_r
    if (ru > 0) {
      Tens G1r = t_wrap(G1.p, {ru, k});
      G1r.str[1] = kk;
      contract(c, "mi,ij->mj", Ur, G1r, T2, -1.0, 1.0);
    }
    colscale_sinv(c, T2.p, m, m, k, S.p, eta);
    if (ru > 0) contract(c, "mi,ij->mj", Ur, Rc, T2, 1.0, 1.0);
  } else {
    T2 = contract_new(c, "mi,ij->mj", U, Rcol);
  }
  contract(c, "mj,nj->mn", T2, Vk, dA, 1.0, 1.0);
  if (dUk && m != kk && !proj) {   // (dU - U U'dU) Sinv V'   (trg.jl:97-99); the projector path above already covers it
    Tens P = t_clone(c, *dUk);
    contract(c, "mi,ij->mj", U, G1, P, -1.0, 1.0);
    colscale_sinv(c, P.p, m, m, k, S.p, eta);
    contract(c, "mj,nj->mn", P, Vk, dA, 1.0, 1.0);
  }
  if (dVk && (n != kk || projv)) {   // U Sinv (dV' - dV'V V')   (trg.jl:101-103)
    Tens Q = t_clone(c, *dVk);
    contract(c, "ni,ij->nj", V, G2, Q, -1.0, 1.0);
    colscale_sinv(c, Q.p, n, n, k, S.p, eta);
    contract(c, "mj,nj->mn", Uk, Q, dA, 1.0, 1.0);
  }
  return dA;
}

// =====================================================================================================
// TRG
// =====================================================================================================
int64_t trg_rank_rule(const std::vector<double>& s, int64_t dmax, double tol) {
  // min(searchsortedfirst(s, tol, rev=true), dmax, length(s))   (trg.jl:37)
  int64_t idx = (int64_t)s.size() + 1;
  for (size_t i = 0; i < s.size(); ++i)
    if (s[i] <= tol) {
      idx = (int64_t)i + 1;
      break;
    }
  return std::min<int64_t>(std::min<int64_t>(idx, dmax), (int64_t)s.size());
}

TrgSplit trg_split(tnad_ctx* c, const Tens& t4, int64_t dmax, double tol) {
  // trg_svd (trg.jl:33-44) on the matrix [(d1,d2),(d3,d4)] of the rank-4 view t4
  TrgSplit sp;
  {
    Span s(c, 1);
    // from order m + n >= 96 on: Jordan-Wielandt embedding + the direct symmetric eigensolver (TNAD_TRG_SVD=jacobi
    // forces the one-sided block Jacobi path)
    const int64_t mm = t4.dim[0] * t4.dim[1], nn = t4.dim[2] * t4.dim[3];
    const char* ev = opt_s(c, "TNAD_TRG_SVD");
    const int coop = c->coop_launch;
    const bool jac = (ev && ev[0] == 'j') || mm + nn < 96 || !coop;
    sp.svd = jac ? svd_jacobi(c, t4, false, nullptr, /*complete_null=*/false) : svd_general_dc(c, t4);
  }
  const int64_t m = t4.dim[0] * t4.dim[1], n = t4.dim[2] * t4.dim[3];
  sp.k = trg_rank_rule(sp.svd.s_host, dmax, tol);
  // The rank rule keeps one singular value <= tol (trg.jl:37).  For an exactly rank-deficient input that
  // triplet is numerically null: sqrt(s) is not differentiable at 0 and the reference's result does not
  // depend on it (it only stays finite there because LAPACK's noise is non-zero).  Treat such values as
  // exact zeros: the factor columns vanish and so do their cotangents (dS = dsqrt/(2 sqrt(s)) := 0).
  {
    std::vector<double> seff = sp.svd.s_host;
    bool any = false;
    for (auto& x : seff)
      if (x <= sp.svd.null_thr) {
        x = 0.0;
        any = true;
      }
    if (any) {
      Tens S2 = t_alloc(c, {(int64_t)seff.size()});
      h2d(c, S2.p, seff.data(), seff.size());
      sync(c);
      sp.svd.S = S2;
    }
  }
  Tens us = t_alloc(c, {m, sp.k}), vs = t_alloc(c, {n, sp.k});
  colscale_sqrt(c, sp.svd.U.p, m, sp.svd.S.p, us.p, m, m, sp.k);
  colscale_sqrt(c, sp.svd.V.p, n, sp.svd.S.p, vs.p, n, n, sp.k);
  sp.us = t_reshape(us, {t4.dim[0], t4.dim[1], sp.k});
  sp.vs = t_reshape(vs, {t4.dim[2], t4.dim[3], sp.k});
  return sp;
}

double trg_forward(tnad_ctx* c, const Tens& a0, int chi, int niter, double tol, TrgTape* tape) {
  TNAD_REQUIRE(a0.rank == 4 && a0.dim[0] == a0.dim[2] && a0.dim[1] == a0.dim[3],
               "trg: tensor must be (d0,d1,d0,d1)");
  TNAD_REQUIRE(chi >= 1 && niter >= 0, "trg: chi >= 1 and niter >= 0 required");
  Tens a = t_clone(c, a0);
  double lnZ = 0.0;
  double* sc = c->scal;
  if (tape) {
    tape->it.clear();
    tape->niter = niter;
    tape->dims0.assign(a0.dim, a0.dim + 4);
  }
  for (int n = 1; n <= niter; ++n) {
    TrgIter it;
    it.a_in = a;
    reduce(c, RED_ABSMAX, a, nullptr, sc + 0);
    d2h(c, &it.maxval, sc + 0, 1);
    TNAD_REQUIRE(it.maxval > 0.0 && std::isfinite(it.maxval), "trg: tensor vanished or is not finite");
    lnZ += std::ldexp(1.0, 1 - n) * std::log(it.maxval);
    Tens an = t_alloc_v(c, std::vector<int64_t>(a.dim, a.dim + 4));
    tcopy(c, a, an, 1.0 / it.maxval, 0.0);
    it.a = an;
    // a[u,r,d,l]: dr_ul = [(d,r),(u,l)], ld_ru = [(l,d),(r,u)]   (trg.jl:20-23), permutes folded into the SVD load
    it.s1 = trg_split(c, t_perm(an, {2, 1, 0, 3}), chi, tol);
    it.s2 = trg_split(c, t_perm(an, {3, 2, 1, 0}), chi, tol);
    // a'[u,r,d,l] = sum dr[n,p,u] ld[p,o,r] ul[d,o,m] ru[l,m,n]   (trg.jl:25)
    Tens dr = it.s1.us, ld = it.s2.us;
    Tens ul = t_perm(it.s1.vs, {2, 0, 1}), ru = t_perm(it.s2.vs, {2, 0, 1});
    {
      Span s(c, 2);
      Tens X = contract_new(c, "npu,por->nour", dr, ld);
      Tens Y = contract_new(c, "dom,lmn->dlno", ul, ru);
      a = contract_new(c, "nour,dlno->urdl", X, Y);
    }
    if (tape) tape->it.push_back(it);
  }
  trace_ijij(c, a, sc + 1);
  double trace;
  d2h(c, &trace, sc + 1, 1);
  lnZ += std::log(trace) / std::ldexp(1.0, niter);
  if (tape) {
    tape->a_final = a;
    tape->trace = trace;
  }
  return lnZ;
}

static Tens trg_split_back(tnad_ctx* c, const TrgSplit& sp, const Tens& du /*(d1,d2,k)*/,
                           const Tens& dvt /*(k,d3,d4)*/, double eta) {
  const int64_t m = sp.svd.U.dim[0], n = sp.svd.V.dim[0], k = sp.k;
  Tens dUk = t_alloc(c, {m, k}), dVk = t_alloc(c, {n, k}), dS = t_alloc(c, {k});
  trg_factor_back(c, m, n, k, sp.svd.U.p, m, sp.svd.V.p, n, sp.svd.S.p, du.p, dvt.p, dUk.p, dVk.p, dS.p);
  return svd_back_dev(c, sp.svd.U, sp.svd.S, sp.svd.V, &dUk, &dS, &dVk, k, eta, sp.svd.rank_left, sp.svd.rank_right);
}

Tens trg_backward(tnad_ctx* c, TrgTape& tape, double dlnZ) {
  Span sp(c, 3);
  const double eta = 1e-40;
  const int niter = tape.niter;
  Tens abar = t_alloc_v(c, std::vector<int64_t>(tape.a_final.dim, tape.a_final.dim + 4), true);
  add_diag_trace_back(c, abar, 1.0 / (std::ldexp(1.0, niter) * tape.trace));
  for (int n = niter; n >= 1; --n) {
    TrgIter& it = tape.it[n - 1];
    Tens dr = it.s1.us, ld = it.s2.us;
    Tens ul = t_perm(it.s1.vs, {2, 0, 1}), ru = t_perm(it.s2.vs, {2, 0, 1});
    Tens X = contract_new(c, "npu,por->nour", dr, ld);
    Tens Y = contract_new(c, "dom,lmn->dlno", ul, ru);
    Tens Xbar = contract_new(c, "urdl,dlno->nour", abar, Y);
    Tens Ybar = contract_new(c, "nour,urdl->dlno", X, abar);
    Tens ddr = contract_new(c, "nour,por->npu", Xbar, ld);
    Tens dld = contract_new(c, "npu,nour->por", dr, Xbar);
    Tens dul = contract_new(c, "dlno,lmn->dom", Ybar, ru);
    Tens dru = contract_new(c, "dom,dlno->lmn", ul, Ybar);
    Tens dt1 = trg_split_back(c, it.s1, ddr, dul, eta);   // d[(d,r),(u,l)]
    Tens dt2 = trg_split_back(c, it.s2, dld, dru, eta);   // d[(l,d),(r,u)]
    const int64_t du_ = it.a.dim[0], dr_ = it.a.dim[1], dd_ = it.a.dim[2], dl_ = it.a.dim[3];
    Tens da = t_alloc(c, {du_, dr_, dd_, dl_});
    Tens v1 = t_perm(t_reshape(dt1, {dd_, dr_, du_, dl_}), {2, 1, 0, 3});
    Tens v2 = t_perm(t_reshape(dt2, {dl_, dd_, dr_, du_}), {3, 2, 1, 0});
    tcopy(c, v1, da, 1.0, 0.0);
    tcopy(c, v2, da, 1.0, 1.0);
    Tens da_in = t_alloc(c, {du_, dr_, dd_, dl_});
    trg_maxval_back(c, da, it.a_in, it.maxval, std::ldexp(1.0, 1 - n), da_in);
    abar = da_in;
  }
  if (dlnZ != 1.0) {
    Tens out = t_alloc_v(c, std::vector<int64_t>(abar.dim, abar.dim + 4));
    tcopy(c, abar, out, dlnZ, 0.0);
    abar = out;
  }
  return abar;
}

// =====================================================================================================
// CTMRG
// =====================================================================================================
void ctmrg_step(tnad_ctx* c, const Tens& bulk, const Tens& corner, const Tens& edge, Tens& corner_out,
                Tens& edge_out, std::vector<double>& vals_host, CtmrgStepRec* rec, Tens* Vwarm) {
  const int64_t D = bulk.dim[0], chi = corner.dim[0], n = chi * D;
  Tens X1, X2, cp;
  {
    Span s(c, 2);
    // cp[i,j,l,k] = sum corner[a,d] edge[i,b,a] edge[d,c,l] bulk[j,k,c,b]   (ctmrg.jl:130)
    X1 = contract_new(c, "iba,ad->ibd", edge, corner);
    X2 = contract_new(c, "ibd,dcl->ibcl", X1, edge);
    cp = contract_new(c, "ibcl,jkcb->ijlk", X2, bulk);
  }
  Tens CP = t_reshape(cp, {n, n});
  SvdResult svd;
  {
    Span s(c, 1);
    // svd(cpmat + cpmat') (ctmrg.jl:134-136); warm-started from the previous step's right vectors
    const char* se = opt_s(c, "TNAD_SYMEIG");
    if (se && se[0] == '0') {
      svd = svd_jacobi(c, CP, true, (Vwarm && Vwarm->p) ? Vwarm : nullptr);   // one-sided path (A/B switch)
      if (Vwarm) *Vwarm = svd.V;
    } else {
      svd = svd_symmetric_auto(c, CP, true, (Vwarm && Vwarm->p) ? Vwarm : nullptr);
      if (Vwarm) *Vwarm = svd.U;
    }
  }
  Tens Z = t_slice_last(svd.U, 0, chi);          // u[:, 1:chi]
  Tens z = t_reshape(Z, {chi, D, chi});           // (ctmrg.jl:137)
  Tens c1, e1, Y1, Y2;
  {
    Span s(c, 2);
    // corner = z' cp z (ctmrg.jl:139); edge = z' (edge*bulk) z without forming tp (ctmrg.jl:131,140)
    Tens W = contract_new(c, "pq,qj->pj", CP, Z);
    c1 = contract_new(c, "pi,pj->ij", Z, W);
    Y1 = contract_new(c, "abi,aed->ibed", z, edge);
    Y2 = contract_new(c, "ibed,bjce->ijcd", Y1, bulk);
    e1 = contract_new(c, "ijcd,dck->ijk", Y2, z);
  }
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
  vals_host.resize((size_t)n);
  const double s0 = svd.s_host[0];
  for (int64_t i = 0; i < n; ++i) vals_host[i] = svd.s_host[i] / s0;   // ctmrg.jl:142
  if (rec) {
    rec->corner = corner;
    rec->edge = edge;
    rec->X1 = X1;
    rec->X2 = X2;
    rec->cp = cp;
    rec->svd = svd;
    rec->Y1 = Y1;
    rec->Y2 = Y2;
    rec->c2 = c2;
    rec->e2 = e2;
    rec->ss = ss;
  }
}

int ctmrg_loop(tnad_ctx* c, const Tens& bulk, Tens& corner, Tens& edge, double tol, int maxit,
               std::vector<double>& vals, CtmrgTape* tape) {
  const int64_t D = bulk.dim[0], chi = corner.dim[0];
  TNAD_REQUIRE(bulk.rank == 4 && bulk.dim[1] == D && bulk.dim[2] == D && bulk.dim[3] == D, "ctmrg: bulk must be D^4");
  TNAD_REQUIRE(corner.rank == 2 && corner.dim[1] == chi && edge.rank == 3 && edge.dim[0] == chi &&
                   edge.dim[1] == D && edge.dim[2] == chi,
               "ctmrg: corner must be chi x chi and edge chi x D x chi");
  if (tape) {
    tape->bulk = bulk;
    tape->D = D;
    tape->chi = chi;
    tape->steps.clear();
  }
  const size_t n = (size_t)(chi * D);
  std::vector<double> oldvals(n, INFINITY);
  vals.assign(n, INFINITY);
  const char* wenv = opt_s(c, "TNAD_WARMSTART");
  const bool warm = (wenv && wenv[0] == '1');   // measured: no gain on the two-sided path (cluster-limited sweeps)
  Tens Vwarm;
  long long counter = -1;   // ctmrg.jl:114
  int nsteps = 0;
  for (;;) {
    counter += 1;           // fixedpoint.jl:32
    if (counter > maxit) break;
    double ss = 0.0;
    bool isnan = false;
    for (size_t i = 0; i < n; ++i) {
      const double d = vals[i] - oldvals[i];
      if (d != d) isnan = true;
      ss += d * d;
    }
    const double diff = std::sqrt(ss);
    if (!isnan && diff <= tol) break;   // NaN <= tol is false
    oldvals = vals;
    Tens cn, en;
    CtmrgStepRec rec;
    // OPT-IN (option TNAD_SHARDED_LOOP = 1 on a context that joined a communicator): the CTMRG steps run chi-sharded over the
    // ranks and every rank keeps the full record, so tnad_ctmrg_backward / tnad_energy work unchanged (forward shared, reverse
    // sweep replicated).  Off by default: the call becomes COLLECTIVE -- every rank of the communicator must make it, a rank
    // that calls tnad_energy on its own would wait in ncclAllGather for ever.
    if (c->comm_world >= 1 && opt_i(c, "TNAD_SHARDED_LOOP", 0) != 0 && corner.dim[0] % c->comm_world == 0)
      ctmrg_step_sharded(c, bulk, corner, edge, cn, en, vals, nullptr, tape ? &rec : nullptr);
    else
      ctmrg_step(c, bulk, corner, edge, cn, en, vals, tape ? &rec : nullptr, warm ? &Vwarm : nullptr);
    if (tape) tape->steps.push_back(rec);
    corner = cn;
    edge = en;
    ++nsteps;
  }
  return nsteps;
}

void ctmrg_step_backward(tnad_ctx* c, const Tens& bulk, const CtmrgStepRec& rec, const Tens& cbar3,
                         const Tens& ebar3, Tens& bulkbar, Tens& cornerbar, Tens& edgebar, double eta) {
  const int64_t D = bulk.dim[0], chi = rec.corner.dim[0], n = chi * D;
  const Tens& corner = rec.corner;
  const Tens& edge = rec.edge;
  Tens CP = t_reshape(rec.cp, {n, n});
  Tens Z = t_slice_last(rec.svd.U, 0, chi);
  Tens z = t_reshape(Z, {chi, D, chi});
  double* dots = c->scal + 32;
  // normalise + symmetrise backward
  reduce(c, RED_DOT, cbar3, &rec.c2, dots);
  reduce(c, RED_DOT, ebar3, &rec.e2, dots + 1);
  Tens cbar2 = t_alloc(c, {chi, chi}), ebar2 = t_alloc(c, {chi, D, chi});
  norm_back(c, cbar3, rec.c2, rec.ss.p, dots, cbar2);
  norm_back(c, ebar3, rec.e2, rec.ss.p + 1, dots + 1, ebar2);
  Tens cbar1 = t_clone(c, cbar2), ebar1 = t_clone(c, ebar2);
  tcopy(c, t_perm(cbar2, {1, 0}), cbar1, 1.0, 1.0);
  tcopy(c, t_perm(ebar2, {2, 1, 0}), ebar1, 1.0, 1.0);
  // corner projection c1 = Z' CP Z  (cbar1 is exactly symmetric)
  Tens T = contract_new(c, "pi,ij->pj", Z, cbar1);
  Tens CPbar = contract_new(c, "pj,qj->pq", T, Z);
  Tens Zbar = contract_new(c, "pq,qj->pj", CP, T);
  contract(c, "qp,qj->pj", CP, T, Zbar, 1.0, 1.0);
  Tens zbar = t_reshape(Zbar, {chi, D, chi});   // indexed [d,c,k] / [a,b,i]
  // edge projection: Y1 = z*edge, Y2 = Y1*bulk, e1 = Y2*z
  Tens Y2bar = contract_new(c, "ijk,dck->ijcd", ebar1, z);
  contract(c, "ijcd,ijk->dck", rec.Y2, ebar1, zbar, 1.0, 1.0);
  Tens Y1bar = contract_new(c, "ijcd,bjce->ibed", Y2bar, bulk);
  contract(c, "ibed,ijcd->bjce", rec.Y1, Y2bar, bulkbar, 1.0, 1.0);
  contract(c, "ibed,aed->abi", Y1bar, edge, zbar, 1.0, 1.0);
  edgebar = contract_new(c, "abi,ibed->aed", z, Y1bar);
  // svd backward: only dU[:, 1:chi] is non-zero, dS = dV = nothing
  Tens Mbar = svd_back_dev(c, rec.svd.U, rec.svd.S, rec.svd.V, &Zbar, nullptr, nullptr, chi, eta);
  tcopy(c, Mbar, CPbar, 1.0, 1.0);                       // cpmat += cpmat'  backward
  tcopy(c, t_perm(Mbar, {1, 0}), CPbar, 1.0, 1.0);
  // grow backward
  Tens cpbar = t_reshape(CPbar, {chi, D, chi, D});
  Tens X2bar = contract_new(c, "ijlk,jkcb->ibcl", cpbar, bulk);
  contract(c, "ibcl,ijlk->jkcb", rec.X2, cpbar, bulkbar, 1.0, 1.0);
  Tens X1bar = contract_new(c, "ibcl,dcl->ibd", X2bar, edge);
  contract(c, "ibd,ibcl->dcl", rec.X1, X2bar, edgebar, 1.0, 1.0);
  contract(c, "ibd,ad->iba", X1bar, corner, edgebar, 1.0, 1.0);
  cornerbar = contract_new(c, "iba,ibd->ad", edge, X1bar);
}

void ctmrg_backward(tnad_ctx* c, const CtmrgTape& tape, const Tens& cbar_in, const Tens& ebar_in, Tens& bulkbar,
                    Tens& cbar0, Tens& ebar0, double eta) {
  const int64_t D = tape.D;
  bulkbar = t_alloc(c, {D, D, D, D}, true);
  Tens cbar = cbar_in, ebar = ebar_in;
  for (size_t i = tape.steps.size(); i-- > 0;) {
    Tens cb, eb;
    ctmrg_step_backward(c, tape.bulk, tape.steps[i], cbar, ebar, bulkbar, cb, eb, eta);
    cbar = cb;
    ebar = eb;
  }
  cbar0 = cbar;
  ebar0 = ebar;
}

// =====================================================================================================
// energy
// =====================================================================================================
double expectationvalue(tnad_ctx* c, const Tens& h, const Tens& ap, const Tens& corner, const Tens& edge,
                        ExpvalTape* tape) {
  Span sp(c, 2);
  const int64_t s = h.dim[0];
  Tens ss_ap = t_alloc(c, {1});
  reduce(c, RED_SUMSQ, ap, nullptr, ss_ap.p);
  Tens apn = t_alloc_v(c, std::vector<int64_t>(ap.dim, ap.dim + ap.rank));
  scale_dev(c, ap, apn, ss_ap.p, SC_INVSQRT);                       // ap /= norm(ap)  (variationalipeps.jl:51)
  // l = ein"ab,ica,bde,cjfdlm,eg,gfk -> ijklm"   (variationalipeps.jl:52), pairwise
  Tens CT1 = contract_new(c, "ica,ab->icb", edge, corner);
  Tens CTr = contract_new(c, "eg,gfk->efk", corner, edge);
  Tens X = contract_new(c, "icb,bde->icde", CT1, edge);
  Tens Y = contract_new(c, "icde,cjfdlm->iejflm", X, apn);
  Tens l = contract_new(c, "iejflm,efk->ijklm", Y, CTr);
  // e = <l, l, h>, n = <tr l, tr l>   (variationalipeps.jl:53-54)
  Tens lh = contract_new(c, "abckl,ijkl->abcij", l, h);
  Tens eye = t_alloc(c, {s, s}, true);
  set_identity(c, eye.p, s, s);
  Tens tl = contract_new(c, "abcij,ij->abc", l, eye);
  double* sc = c->scal + 40;
  reduce(c, RED_DOT, l, &lh, sc);
  reduce(c, RED_SUMSQ, tl, nullptr, sc + 1);
  double en[2];
  d2h(c, en, sc, 2);
  if (tape) {
    tape->h = h;
    tape->ap = ap;
    tape->apn = apn;
    tape->CT1 = CT1;
    tape->CTr = CTr;
    tape->X = X;
    tape->Y = Y;
    tape->l = l;
    tape->tl = tl;
    tape->ss_ap = ss_ap;
    tape->e = en[0];
    tape->nn = en[1];
  }
  return en[0] / en[1];
}

void expectationvalue_back(tnad_ctx* c, const Tens& corner, const Tens& edge, const ExpvalTape& t, double ybar,
                           Tens& apbar, Tens& cornerbar, Tens& edgebar) {
  const int64_t s = t.h.dim[0];
  const double ebar = ybar / t.nn, nbar = -ybar * t.e / (t.nn * t.nn);
  // lbar = ebar * l.(h + h^T(kl<->ij)) + 2 nbar tl (x) delta      
  Tens hs = t_clone(c, t.h);
  tcopy(c, t_perm(t.h, {2, 3, 0, 1}), hs, 1.0, 1.0);
  Tens lbar = contract_new(c, "abckl,ijkl->abcij", t.l, hs, ebar);
  Tens eye = t_alloc(c, {s, s}, true);
  set_identity(c, eye.p, s, s);
  contract(c, "abc,ij->abcij", t.tl, eye, lbar, 2.0 * nbar, 1.0);
  Tens Ybar = contract_new(c, "ijklm,efk->iejflm", lbar, t.CTr);
  Tens CTrbar = contract_new(c, "iejflm,ijklm->efk", t.Y, lbar);
  Tens Xbar = contract_new(c, "iejflm,cjfdlm->icde", Ybar, t.apn);
  Tens apnbar = contract_new(c, "icde,iejflm->cjfdlm", t.X, Ybar);
  Tens CT1bar = contract_new(c, "icde,bde->icb", Xbar, edge);
  edgebar = contract_new(c, "icb,icde->bde", t.CT1, Xbar);
  cornerbar = contract_new(c, "efk,gfk->eg", CTrbar, edge);
  contract(c, "eg,efk->gfk", corner, CTrbar, edgebar, 1.0, 1.0);
  contract(c, "icb,ab->ica", CT1bar, corner, edgebar, 1.0, 1.0);
  contract(c, "ica,icb->ab", edge, CT1bar, cornerbar, 1.0, 1.0);
  double* dot = c->scal + 44;
  reduce(c, RED_DOT, apnbar, &t.ap, dot);
  apbar = t_alloc_v(c, std::vector<int64_t>(t.ap.dim, t.ap.dim + t.ap.rank));
  norm_back(c, apnbar, t.ap, t.ss_ap.p, dot, apbar);
}

double energy(tnad_ctx* c, const Tens& h, const Tens& A, int chi, double tol, int maxit, Tens* gradA, int* steps) {
  TNAD_REQUIRE(A.rank == 5 && A.dim[0] == A.dim[1] && A.dim[1] == A.dim[2] && A.dim[2] == A.dim[3],
               "size of tensor error, should be (d, d, d, d, s)");   // ipeps.jl:19-20
  const int64_t s = A.dim[4];
  TNAD_REQUIRE(h.rank == 4 && h.dim[0] == s && h.dim[1] == s && h.dim[2] == s && h.dim[3] == s,
               "energy: h must be (s,s,s,s)");
  TNAD_REQUIRE(chi >= 1 && maxit >= 0, "energy: chi >= 1, maxit >= 0 required");
  const double eta = 1e-40;
  // ipeps = indexperm_symmetrize(ipeps)   (variationalipeps.jl:29)
  Tens xsum, As;
  Tens ss_sym = t_alloc(c, {1});
  ipeps_symmetrize(c, A, xsum, As, ss_sym.p);
  Tens ap, a;
  double_layer(c, As, ap, a);                              // variationalipeps.jl:30-34
  const int64_t D = a.dim[0];
  Tens corner = t_alloc(c, {(int64_t)chi, (int64_t)chi}), edge = t_alloc(c, {(int64_t)chi, D, (int64_t)chi});
  init_raw(c, a, corner, edge);                            // variationalipeps.jl:36 (constant under AD)
  CtmrgTape tape;
  std::vector<double> vals;
  int nsteps = ctmrg_loop(c, a, corner, edge, tol, maxit, vals, gradA ? &tape : nullptr);
  if (steps) *steps = nsteps;
  ExpvalTape et;
  double y = expectationvalue(c, h, ap, corner, edge, gradA ? &et : nullptr);
  if (gradA) {
    Span sp(c, 3);
    Tens apbar, cbar, ebar, abar, cb0, eb0, Asbar;
    expectationvalue_back(c, corner, edge, et, 1.0, apbar, cbar, ebar);
    ctmrg_backward(c, tape, cbar, ebar, abar, cb0, eb0, eta);
    double_layer_back(c, As, apbar, abar, Asbar);
    ipeps_symmetrize_back(c, Asbar, xsum, ss_sym.p, *gradA);
  }
  return y;
}

// Implicit (fixed-point) gradient of the energy -- opt-in alternative to the unrolled reverse sweep (north_star item 3,
// SURVEY 8f.2).  At a converged environment x* = f(x*, a) of the gauge-fixed step f (canonical column signs make the fixed
// point elementwise), the cotangent of a is  J_a' (1 - J_x')^{-1} xbar  =  sum_k J_a' (J_x')^k xbar: the pullback of ONE
// recorded step is applied repeatedly to the shrinking cotangent (Neumann series) until it has decayed by bwd_tol.  No
// tape over the forward iterations: one step record instead of one per step.  Equal to the reference's gradient (which
// back-propagates through every executed step with a constant initialisation, autodiff.jl:5) in the limit of a converged
// CTMRG; it is NOT what the reference computes for short fixed-maxit runs.
double energy_fixedpoint(tnad_ctx* c, const Tens& h, const Tens& A, int chi, double tol, int maxit, double bwd_tol,
                         int bwd_maxit, Tens* gradA, int* steps, int* bwd_iters) {
  TNAD_REQUIRE(A.rank == 5 && A.dim[0] == A.dim[1] && A.dim[1] == A.dim[2] && A.dim[2] == A.dim[3],
               "size of tensor error, should be (d, d, d, d, s)");   // ipeps.jl:19-20
  const int64_t s = A.dim[4];
  TNAD_REQUIRE(h.rank == 4 && h.dim[0] == s && h.dim[1] == s && h.dim[2] == s && h.dim[3] == s, "energy: h must be (s,s,s,s)");
  TNAD_REQUIRE(chi >= 1 && maxit >= 0 && bwd_maxit >= 1 && bwd_tol > 0.0, "energy_fixedpoint: bad arguments");
  const double eta = 1e-40;
  Tens xsum, As;
  Tens ss_sym = t_alloc(c, {1});
  ipeps_symmetrize(c, A, xsum, As, ss_sym.p);
  Tens ap, a;
  double_layer(c, As, ap, a);
  const int64_t D = a.dim[0];
  Tens corner = t_alloc(c, {(int64_t)chi, (int64_t)chi}), edge = t_alloc(c, {(int64_t)chi, D, (int64_t)chi});
  init_raw(c, a, corner, edge);
  std::vector<double> vals;
  int nsteps = ctmrg_loop(c, a, corner, edge, tol, maxit, vals, nullptr);     // no tape
  // one recorded step at the (converged) environment: its pullback is the linear map of the Neumann series
  CtmrgStepRec rec;
  Tens cn, en;
  ctmrg_step(c, a, corner, edge, cn, en, vals, gradA ? &rec : nullptr);
  ++nsteps;
  if (steps) *steps = nsteps;
  ExpvalTape et;
  const double y = expectationvalue(c, h, ap, cn, en, gradA ? &et : nullptr);
  if (bwd_iters) *bwd_iters = 0;
  if (gradA) {
    Span sp(c, 3);
    // the implicit formula needs an ELEMENTWISE fixed point of the gauge-fixed step: f(x*) = x*.  With degenerate singular
    // values at (or inside) the kept spectrum the eigenvectors of a multiplet rotate from step to step and only gauge
    // invariants converge: refuse instead of returning a wrong gradient (the unrolled sweep handles that case).
    {
      Tens dc = t_clone(c, cn), de = t_clone(c, en);
      tcopy(c, corner, dc, -1.0, 1.0);
      tcopy(c, edge, de, -1.0, 1.0);
      double* sq = c->scal + 56;
      reduce(c, RED_SUMSQ, dc, nullptr, sq);
      reduce(c, RED_SUMSQ, de, nullptr, sq + 1);
      double v[2];
      d2h(c, v, sq, 2);
      const double res = std::sqrt(v[0] + v[1]);       // corner and edge have unit norm
      if (!(res <= opt_d(c, "TNAD_FIXEDPOINT_RES", 1e-7)))
        fail(TNAD_ERR_NOCONV, "energy_fixedpoint: the environment is not an elementwise fixed point of the step (|f(x) - x| = " +
                                  std::to_string(res) + "): not converged, or degenerate singular values inside the kept spectrum; "
                                  "use the unrolled gradient (tnad_energy)");
    }
    Tens apbar, cbar, ebar, Asbar;
    expectationvalue_back(c, cn, en, et, 1.0, apbar, cbar, ebar);
    Tens abar = t_alloc(c, {D, D, D, D}, true);
    double* sq = c->scal + 52;
    double nrm0 = 0.0;
    int k = 0;
    for (; k < bwd_maxit; ++k) {
      reduce(c, RED_SUMSQ, cbar, nullptr, sq);
      reduce(c, RED_SUMSQ, ebar, nullptr, sq + 1);
      double v[2];
      d2h(c, v, sq, 2);
      const double nrm = std::sqrt(v[0] + v[1]);
      TNAD_REQUIRE(std::isfinite(nrm), "energy_fixedpoint: the cotangent series diverged (is the environment converged?)");
      if (k == 0) nrm0 = nrm;
      if (nrm <= bwd_tol * nrm0) break;
      Tens cb, eb;
      ctmrg_step_backward(c, a, rec, cbar, ebar, abar, cb, eb, eta);   // abar += J_a' lambda_k; (cb, eb) = J_x' lambda_k
      cbar = cb;
      ebar = eb;
    }
    if (bwd_iters) *bwd_iters = k;
    double_layer_back(c, As, apbar, abar, Asbar);
    ipeps_symmetrize_back(c, Asbar, xsum, ss_sym.p, *gradA);
  }
  return y;
}

double magnetisation_readout(tnad_ctx* c, const Tens& a, const Tens& m, const Tens& corner, const Tens& edge,
                             MagTape* tape) {
  // exampletensors.jl:63-68
  Tens ct = contract_new(c, "ia,ajb->ijb", corner, edge);
  Tens ctc = contract_new(c, "ijb,bk->ijk", ct, corner);
  Tens e1 = contract_new(c, "alc,ckd->alkd", ctc, edge);
  Tens e2 = contract_new(c, "bjd,bia->jdia", ctc, edge);
  Tens env = contract_new(c, "alkd,jdia->ijkl", e1, e2);
  double* sc = c->scal + 48;
  reduce(c, RED_DOT, env, &m, sc);
  reduce(c, RED_DOT, env, &a, sc + 1);
  double v[2];
  d2h(c, v, sc, 2);
  if (tape) {
    tape->ct = ct; tape->ctc = ctc; tape->e1 = e1; tape->e2 = e2; tape->env = env;
    tape->mag = v[0]; tape->nrm = v[1];
  }
  return std::fabs(v[0] / v[1]);
}

// Reverse of the read-out as Zygote derives it (test/ctmrg.jl:44-46 differentiates magnetisation): y = |mag/norm|.
void magnetisation_readout_back(tnad_ctx* c, const Tens& a, const Tens& m, const Tens& corner, const Tens& edge,
                                const MagTape& t, double ybar, Tens& abar, Tens& mbar, Tens& cornerbar, Tens& edgebar) {
  const double sg = (t.mag / t.nrm) >= 0.0 ? 1.0 : -1.0;
  const double magbar = ybar * sg / t.nrm, nrmbar = -ybar * sg * t.mag / (t.nrm * t.nrm);
  const std::vector<int64_t> d4(a.dim, a.dim + 4);
  Tens envbar = t_alloc_v(c, d4);
  tcopy(c, m, envbar, magbar, 0.0);
  tcopy(c, a, envbar, nrmbar, 1.0);
  abar = t_alloc_v(c, d4);
  mbar = t_alloc_v(c, d4);
  tcopy(c, t.env, abar, nrmbar, 0.0);
  tcopy(c, t.env, mbar, magbar, 0.0);
  Tens e1bar = contract_new(c, "ijkl,jdia->alkd", envbar, t.e2);
  Tens e2bar = contract_new(c, "alkd,ijkl->jdia", t.e1, envbar);
  Tens ctcbar = contract_new(c, "alkd,ckd->alc", e1bar, edge);
  contract(c, "jdia,bia->bjd", e2bar, edge, ctcbar, 1.0, 1.0);
  edgebar = contract_new(c, "alc,alkd->ckd", t.ctc, e1bar);
  contract(c, "bjd,jdia->bia", t.ctc, e2bar, edgebar, 1.0, 1.0);
  Tens ctbar = contract_new(c, "ijk,bk->ijb", ctcbar, corner);
  cornerbar = contract_new(c, "ijb,ijk->bk", t.ct, ctcbar);
  contract(c, "ijb,ajb->ia", ctbar, edge, cornerbar, 1.0, 1.0);
  contract(c, "ia,ijb->ajb", corner, ctbar, edgebar, 1.0, 1.0);
}

}  // namespace tnad
