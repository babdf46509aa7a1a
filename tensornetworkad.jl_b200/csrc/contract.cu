// Pairwise einsum -> multi-level-stride GEMM descriptor (host logic only; the kernel is gemm.cu).
//
// Replaces OMEinsum's pairwise contraction (permutedims + reshape + BLAS gemm) behind every
// `ein"..."` call of the reference.  Labels are classified as batch (in A, B and C), M (A and C),
// N (B and C) or K (A and B); each group becomes up to MAXL (extent, stride) levels per operand.
// Adjacent labels are merged when they are jointly contiguous in both operands that carry them.
//
// tnad_contract_plan() exports the descriptor as int64[128] so that the CPU test-suite can check
// the stride algebra against numpy.einsum without a GPU:
//   [0..7]   M, N, K, batch, a_kfast, b_kfast, a_vec, b_vec
//   [8+9*i]  level set i = {nl, n0..n3, s0..s3}, i in order am, ak, bk, bn, cm, cn, ab, bb, cb
#include "common.h"
#include <algorithm>

namespace tnad {

namespace {

struct Lab {
  char ch;
  int64_t n;
  int64_t sa, sb, sc;
  bool inA, inB, inC;
};

struct Level {
  int64_t n, s1, s2, s3;
};

void build_levels(std::vector<Lab*>& labs, int which1, int which2, int which3, LvlSet* o1, LvlSet* o2,
                  LvlSet* o3, const char* what) {
  auto get = [](const Lab* l, int w) -> int64_t { return w == 0 ? l->sa : (w == 1 ? l->sb : l->sc); };
  std::vector<Level> lv;
  for (Lab* l : labs) {
    Level x{l->n, get(l, which1), get(l, which2), which3 >= 0 ? get(l, which3) : 0};
    if (!lv.empty()) {
      Level& p = lv.back();
      bool ok = x.s1 == p.s1 * p.n && x.s2 == p.s2 * p.n && (which3 < 0 || x.s3 == p.s3 * p.n);
      if (ok) {
        p.n *= x.n;
        continue;
      }
    }
    lv.push_back(x);
  }
  if (lv.empty()) lv.push_back(Level{1, 0, 0, 0});
  if ((int)lv.size() > MAXL)
    fail(TNAD_ERR_INTERNAL, std::string("contract: too many stride levels in group ") + what);
  LvlSet* outs[3] = {o1, o2, o3};
  for (int w = 0; w < 3; ++w) {
    if (!outs[w]) continue;
    LvlSet& L = *outs[w];
    L.nl = (int)lv.size();
    for (int i = 0; i < MAXL; ++i) {
      L.n[i] = 1;
      L.s[i] = 0;
    }
    for (size_t i = 0; i < lv.size(); ++i) {
      if (lv[i].n > 0x7fffffffLL) fail(TNAD_ERR_ARG, "contract: extent too large");
      L.n[i] = (int)lv[i].n;
      L.s[i] = w == 0 ? lv[i].s1 : (w == 1 ? lv[i].s2 : lv[i].s3);
    }
  }
}

bool all_even_except_unit(const LvlSet& fast, const LvlSet& o1, const LvlSet& o2) {
  // fast.s[0] is the unit stride; everything else must be even so 16-byte chunks stay aligned
  for (int l = 1; l < fast.nl; ++l)
    if (fast.s[l] & 1) return false;
  for (int l = 0; l < o1.nl; ++l)
    if (o1.n[l] > 1 && (o1.s[l] & 1)) return false;
  for (int l = 0; l < o2.nl; ++l)
    if (o2.n[l] > 1 && (o2.s[l] & 1)) return false;
  return true;
}

}  // namespace

GemmDesc contract_plan(const char* spec, const Tens& A, const Tens& B, const Tens& C) {
  std::string s(spec);
  s.erase(std::remove(s.begin(), s.end(), ' '), s.end());
  size_t comma = s.find(','), arrow = s.find("->");
  TNAD_REQUIRE(comma != std::string::npos && arrow != std::string::npos && comma < arrow,
               "contract: spec must look like 'ab,bc->ac'");
  std::string la = s.substr(0, comma), lb = s.substr(comma + 1, arrow - comma - 1), lc = s.substr(arrow + 2);
  TNAD_REQUIRE((int)la.size() == A.rank && (int)lb.size() == B.rank && (int)lc.size() == C.rank,
               std::string("contract: rank mismatch for spec ") + spec);

  std::vector<Lab> labs;
  auto find = [&](char ch) -> Lab* {
    for (auto& l : labs)
      if (l.ch == ch) return &l;
    return nullptr;
  };
  auto add = [&](const std::string& ls, const Tens& T, int which) {
    for (size_t i = 0; i < ls.size(); ++i) {
      char ch = ls[i];
      Lab* l = find(ch);
      if (!l) {
        labs.push_back(Lab{ch, T.dim[i], 0, 0, 0, false, false, false});
        l = &labs.back();
      }
      TNAD_REQUIRE(l->n == T.dim[i], std::string("contract: extent mismatch for label '") + ch + "' in " + spec);
      bool& in = which == 0 ? l->inA : (which == 1 ? l->inB : l->inC);
      TNAD_REQUIRE(!in, std::string("contract: repeated label within one operand in ") + spec);
      in = true;
      (which == 0 ? l->sa : (which == 1 ? l->sb : l->sc)) = T.str[i];
    }
  };
  add(la, A, 0);
  add(lb, B, 1);
  add(lc, C, 2);

  std::vector<Lab*> gb, gm, gn, gk;
  for (auto& l : labs) {
    if (l.n == 1) continue;
    if (l.inA && l.inB && l.inC) gb.push_back(&l);
    else if (l.inA && l.inC && !l.inB) gm.push_back(&l);
    else if (l.inB && l.inC && !l.inA) gn.push_back(&l);
    else if (l.inA && l.inB && !l.inC) gk.push_back(&l);
    else fail(TNAD_ERR_ARG, std::string("contract: label '") + l.ch + "' appears in only one tensor in " + spec);
  }
  auto has_unit = [](const std::vector<Lab*>& g, int w) {
    for (auto* l : g)
      if ((w == 0 ? l->sa : (w == 1 ? l->sb : l->sc)) == 1) return true;
    return false;
  };
  auto sort_by = [](std::vector<Lab*>& g, int w) {
    std::stable_sort(g.begin(), g.end(), [w](const Lab* x, const Lab* y) {
      int64_t a = w == 0 ? x->sa : (w == 1 ? x->sb : x->sc);
      int64_t b = w == 0 ? y->sa : (w == 1 ? y->sb : y->sc);
      return a < b;
    });
  };
  const bool a_mfast = has_unit(gm, 0);
  const bool b_nfast = has_unit(gn, 1);
  sort_by(gm, a_mfast ? 0 : 2);
  sort_by(gn, b_nfast ? 1 : 2);
  if (!a_mfast && has_unit(gk, 0)) sort_by(gk, 0);
  else if (!b_nfast && has_unit(gk, 1)) sort_by(gk, 1);
  else sort_by(gk, 0);
  sort_by(gb, 2);

  GemmDesc d;
  memset(&d, 0, sizeof(d));
  build_levels(gm, 0, 2, -1, &d.am, &d.cm, nullptr, "M");
  build_levels(gn, 1, 2, -1, &d.bn, &d.cn, nullptr, "N");
  build_levels(gk, 0, 1, -1, &d.ak, &d.bk, nullptr, "K");
  build_levels(gb, 0, 1, 2, &d.ab, &d.bb, &d.cb, "batch");
  auto total = [](const LvlSet& L) {
    int64_t t = 1;
    for (int i = 0; i < L.nl; ++i) t *= L.n[i];
    return t;
  };
  int64_t M = total(d.am), N = total(d.bn), K = total(d.ak), Bt = total(d.ab);
  TNAD_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31) && Bt < 65536, "contract: problem too large");
  d.M = (int)M;
  d.N = (int)N;
  d.K = (int)K;
  d.batch = (int)Bt;
  d.A = A.p;
  d.B = B.p;
  d.C = C.p;
  d.alpha = 1.0;
  d.beta = 0.0;
  d.a_kfast = a_mfast ? 0 : 1;
  d.b_kfast = b_nfast ? 0 : 1;
  const bool a_al = (reinterpret_cast<uintptr_t>(A.p) & 15) == 0;
  const bool b_al = (reinterpret_cast<uintptr_t>(B.p) & 15) == 0;
  if (a_mfast)
    d.a_vec = a_al && d.am.s[0] == 1 && (d.am.n[0] % 2 == 0) && all_even_except_unit(d.am, d.ak, d.ab);
  else
    d.a_vec = a_al && d.ak.s[0] == 1 && (d.ak.n[0] % 2 == 0) && all_even_except_unit(d.ak, d.am, d.ab);
  if (b_nfast)
    d.b_vec = b_al && d.bn.s[0] == 1 && (d.bn.n[0] % 2 == 0) && all_even_except_unit(d.bn, d.bk, d.bb);
  else
    d.b_vec = b_al && d.bk.s[0] == 1 && (d.bk.n[0] % 2 == 0) && all_even_except_unit(d.bk, d.bn, d.bb);
  return d;
}

void contract(tnad_ctx* c, const char* spec, const Tens& A, const Tens& B, Tens& C, double alpha, double beta) {
  GemmDesc d;
  {
    HostTimer ht(c, 0);
    d = contract_plan(spec, A, B, C);
  }
  d.alpha = alpha;
  d.beta = beta;
  gemm_run(c, d);
}

Tens contract_new(tnad_ctx* c, const char* spec, const Tens& A, const Tens& B, double alpha) {
  std::string s(spec);
  s.erase(std::remove(s.begin(), s.end(), ' '), s.end());
  size_t comma = s.find(','), arrow = s.find("->");
  TNAD_REQUIRE(comma != std::string::npos && arrow != std::string::npos, "contract: bad spec");
  std::string la = s.substr(0, comma), lb = s.substr(comma + 1, arrow - comma - 1), lc = s.substr(arrow + 2);
  TNAD_REQUIRE((int)la.size() == A.rank && (int)lb.size() == B.rank,
               std::string("contract: rank mismatch for spec ") + spec);
  std::vector<int64_t> dims;
  for (char ch : lc) {
    int64_t n = -1;
    size_t ia = la.find(ch), ib = lb.find(ch);
    if (ia != std::string::npos) n = A.dim[ia];
    else if (ib != std::string::npos) n = B.dim[ib];
    TNAD_REQUIRE(n >= 0, std::string("contract: output label not found in inputs: ") + spec);
    dims.push_back(n);
  }
  Tens C = t_alloc_v(c, dims);
  contract(c, spec, A, B, C, alpha, 0.0);
  return C;
}

}  // namespace tnad

extern "C" int tnad_contract_plan(const char* spec, const int64_t* dimsA, int rankA, const int64_t* dimsB,
                                  int rankB, int64_t* plan) {
  using namespace tnad;
  try {
    if (!spec || !dimsA || !dimsB || !plan || rankA > MAXR || rankB > MAXR) return TNAD_ERR_ARG;
    std::vector<int64_t> da(dimsA, dimsA + rankA), db(dimsB, dimsB + rankB);
    Tens A = t_wrap(nullptr, da), B = t_wrap(nullptr, db);
    std::string s(spec);
    s.erase(std::remove(s.begin(), s.end(), ' '), s.end());
    size_t comma = s.find(','), arrow = s.find("->");
    if (comma == std::string::npos || arrow == std::string::npos) return TNAD_ERR_ARG;
    std::string la = s.substr(0, comma), lb = s.substr(comma + 1, arrow - comma - 1), lc = s.substr(arrow + 2);
    if ((int)la.size() != rankA || (int)lb.size() != rankB) return TNAD_ERR_ARG;
    std::vector<int64_t> dc;
    for (char ch : lc) {
      size_t ia = la.find(ch), ib = lb.find(ch);
      if (ia != std::string::npos) dc.push_back(da[ia]);
      else if (ib != std::string::npos) dc.push_back(db[ib]);
      else return TNAD_ERR_ARG;
    }
    Tens C = t_wrap(nullptr, dc);
    GemmDesc d = contract_plan(spec, A, B, C);
    for (int i = 0; i < 128; ++i) plan[i] = 0;
    plan[0] = d.M; plan[1] = d.N; plan[2] = d.K; plan[3] = d.batch;
    plan[4] = d.a_kfast; plan[5] = d.b_kfast; plan[6] = d.a_vec; plan[7] = d.b_vec;
    const LvlSet* sets[9] = {&d.am, &d.ak, &d.bk, &d.bn, &d.cm, &d.cn, &d.ab, &d.bb, &d.cb};
    for (int i = 0; i < 9; ++i) {
      int64_t* q = plan + 8 + 9 * i;
      q[0] = sets[i]->nl;
      for (int l = 0; l < MAXL; ++l) {
        q[1 + l] = sets[i]->n[l];
        q[5 + l] = sets[i]->s[l];
      }
    }
    return TNAD_OK;
  } catch (const tnad::Error& e) {
    return e.code;
  } catch (...) {
    return TNAD_ERR_INTERNAL;
  }
}
