// Whole-algorithm drivers (device-resident state; host code only sequences kernels and applies the
// reference's scalar control rules: the TRG rank rule trg.jl:37 and the CTMRG stop rule fixedpoint.jl:31-41).
#pragma once
#include "common.h"

namespace tnad {

// svd_back (trg.jl:72-105) for cotangents that live in the first k columns of U / V.
// U: m x kk, S: kk, V: n x kk (kk = min(m,n)); dUk: m x k, dS: k, dVk: n x k (any may be null).
Tens svd_back_dev(tnad_ctx* c, const Tens& U, const Tens& S, const Tens& V, const Tens* dUk, const Tens* dS,
                  const Tens* dVk, int64_t k, double eta, int64_t r_valid = -1, int64_t r_valid_v = -1);

// ---- TRG ------------------------------------------------------------------------------------------
struct TrgSplit {
  SvdResult svd;
  int64_t k = 0;
  Tens us;   // U[:, :k] * sqrt(s)  as (d1, d2, k)
  Tens vs;   // V[:, :k] * sqrt(s)  as (d3, d4, k)
};
struct TrgIter {
  Tens a_in, a;
  double maxval = 1.0;
  TrgSplit s1, s2;
};
struct TrgTape {
  std::vector<TrgIter> it;
  Tens a_final;
  double trace = 0.0;
  int niter = 0;
  std::vector<int64_t> dims0;
};
double trg_forward(tnad_ctx* c, const Tens& a0, int chi, int niter, double tol, TrgTape* tape);
Tens trg_backward(tnad_ctx* c, TrgTape& tape, double dlnZ);
int64_t trg_rank_rule(const std::vector<double>& s, int64_t dmax, double tol);
TrgSplit trg_split(tnad_ctx* c, const Tens& t4, int64_t dmax, double tol);

// ---- CTMRG ----------------------------------------------------------------------------------------
struct CtmrgStepRec {
  Tens corner, edge;         // inputs of the step
  Tens X1, X2, cp;           // grow intermediates; cp is (chi, D, chi, D)
  SvdResult svd;
  Tens Y1, Y2;
  Tens c2, e2;               // symmetrised, un-normalised outputs
  Tens ss;                   // device scalars: ss[0] = |c2|^2, ss[1] = |e2|^2
};
struct CtmrgTape {
  Tens bulk;
  int64_t D = 0, chi = 0;
  std::vector<CtmrgStepRec> steps;
};
// one step (ctmrg.jl:126-153); vals_host gets s ./ s[1]
void ctmrg_step(tnad_ctx* c, const Tens& bulk, const Tens& corner, const Tens& edge, Tens& corner_out,
                Tens& edge_out, std::vector<double>& vals_host, CtmrgStepRec* rec, Tens* Vwarm = nullptr);
// the fixed-point loop (ctmrg.jl:110-117 + fixedpoint.jl); returns the number of steps executed
int ctmrg_loop(tnad_ctx* c, const Tens& bulk, Tens& corner, Tens& edge, double tol, int maxit,
               std::vector<double>& vals_host, CtmrgTape* tape);
void ctmrg_step_backward(tnad_ctx* c, const Tens& bulk, const CtmrgStepRec& rec, const Tens& cbar3,
                         const Tens& ebar3, Tens& bulkbar /*accumulated*/, Tens& cornerbar, Tens& edgebar, double eta);
void ctmrg_backward(tnad_ctx* c, const CtmrgTape& tape, const Tens& cbar, const Tens& ebar, Tens& bulkbar,
                    Tens& cbar0, Tens& ebar0, double eta);

// ---- energy ---------------------------------------------------------------------------------------
struct ExpvalTape {
  Tens h, ap, apn, CT1, CTr, X, Y, l, tl, ss_ap;
  double e = 0.0, nn = 0.0;
};
double expectationvalue(tnad_ctx* c, const Tens& h, const Tens& ap, const Tens& corner, const Tens& edge,
                        ExpvalTape* tape);
void expectationvalue_back(tnad_ctx* c, const Tens& corner, const Tens& edge, const ExpvalTape& t, double ybar,
                           Tens& apbar, Tens& cornerbar, Tens& edgebar);
double energy(tnad_ctx* c, const Tens& h, const Tens& A, int chi, double tol, int maxit, Tens* gradA, int* steps);
double energy_fixedpoint(tnad_ctx* c, const Tens& h, const Tens& A, int chi, double tol, int maxit, double bwd_tol,
                         int bwd_maxit, Tens* gradA, int* steps, int* bwd_iters);
struct MagTape {
  Tens ct, ctc, e1, e2, env;
  double mag = 0.0, nrm = 1.0;
};
double magnetisation_readout(tnad_ctx* c, const Tens& a, const Tens& m, const Tens& corner, const Tens& edge,
                             MagTape* tape = nullptr);
void magnetisation_readout_back(tnad_ctx* c, const Tens& a, const Tens& m, const Tens& corner, const Tens& edge,
                                const MagTape& t, double ybar, Tens& abar, Tens& mbar, Tens& cornerbar, Tens& edgebar);

// chi-sharded ctmrgstep over the ranks of the context's communicator (sharded.cu)
// rec (optional): the record ctmrg_step_backward needs, so that the unrolled reverse sweep (replicated on every rank) can follow
// a forward pass whose steps were sharded
void ctmrg_step_sharded(tnad_ctx* c, const Tens& bulk, const Tens& corner, const Tens& edge, Tens& corner_out, Tens& edge_out,
                        std::vector<double>& vals_host, double* ms, CtmrgStepRec* rec = nullptr);

}  // namespace tnad

struct tnad_tape {
  tnad_ctx* ctx = nullptr;
  int kind = 0;   // 1 = TRG, 2 = CTMRG
  tnad::TrgTape trg;
  tnad::CtmrgTape ctmrg;
};
