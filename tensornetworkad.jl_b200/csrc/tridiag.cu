// Householder tridiagonalisation of a dense symmetric matrix and the matching back-transformation: the O(n^3)
// direct route to the eigen-decomposition ctmrgstep needs (ctmrg.jl:133-136), next to the block-Jacobi solver of
// symeig.cu (which spends ~20 sweeps x 8 n^3 flops; this path needs ~(4/3 + 2 + 2) n^3).
//
//   A = Q T Q',  Q = H_0 H_1 ... H_{n-3},  H_j = I - tau_j v_j v_j'
//
// Three kernels share the work (sytrd() below picks per column range):
//   k_sytrd_panel1        default for the columns whose trailing matrix does not fit a cluster: ONE persistent cooperative
//                         kernel per panel of 32 columns, ONE grid-wide exchange per column (index ownership, matvec by
//                         column ownership on a vector that does not need the previous column's pending scalar, cancellation
//                         guard, sector-packed exchange buffers, TMA-staged shared-memory cache of the CTA's columns).
//   k_sytrd_tail_cluster  the last <= 416 columns (and whole problems up to that order) on one thread-block cluster with the
//                         trailing matrix resident in distributed shared memory (unblocked dsytd2 update, DSMEM exchange,
//                         two hardware cluster barriers per column).
//   k_sytrd_panel         the first version (A/B switch TNAD_SYTRD_1B=0): dlatrd-style, two grid barriers per column;
//                         phase B forms p_c = A22[:, c] . v as independent coalesced column dots (A22 is stored in full),
//                         phase C forms w and brings the next column of A up to date.
// The rank-64 trailing update A22 -= V W' + W V' between panels is one DMMA GEMM (contract) with K = 64.
// Algorithmic traffic: the trailing matrix once per column, 8 n^3 / 3 bytes per decomposition.
// Also here: the compact-WY back-transformation (apply_q), the symmetric driver (svd_symmetric_dc) and the general
// driver through the Jordan-Wielandt embedding (svd_general_dc).
#include "drivers.h"
#include "eigdc.h"
#include <algorithm>
#include <numeric>
#include <cooperative_groups.h>

namespace tnad {

namespace {

constexpr int ST_NT = 512;          // threads per CTA of the panel kernel (one CTA per SM)
constexpr int ST_NW = ST_NT / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All CTAs of a cooperative launch meet here: CTA b publishes `epoch` in its own flag line (release) and every CTA
// polls all flags (one acquire load per thread) - no contended atomic, one L2 round trip after the last arrival.
// Counter variant (A/B switch TNAD_SYTRD_BAR=1): one release-add per CTA on a single counter, thread 0 polls it.
__device__ __forceinline__ void grid_barrier_ctr(unsigned int* ctr, unsigned int target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    const long long t0 = clock64();
    unsigned int v;
    for (;;) {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 6000000000LL) {
        *err = 1;
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void grid_barrier(unsigned int* flags, unsigned int epoch, int G, int* err, int* s_to) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(flags + 32 * blockIdx.x), "r"(epoch) : "memory");
  const long long t0 = clock64();
  for (;;) {
    int bad = 0;
    for (int q = threadIdx.x; q < G; q += ST_NT) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * q) : "memory");
      bad |= ((int)(v - epoch) < 0) ? 1 : 0;
    }
    if (threadIdx.x == 0 && clock64() - t0 > 6000000000LL) {   // ~3 s: never hang the device on a lost CTA
      *s_to = 1;
      *err = 1;
    }
    if (__syncthreads_count(bad) == 0 || *s_to) break;
  }
}

// block-wide sum, identical result in every thread; red: >= ST_NW doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < ST_NW; ++w) s += red[w];
  return s;
}

// P1 = [V W], P2 = [W V]  (n x 2nb each, leading dimension ldp): A22 -= P1 P2' after the panel.
// Work item of phase B = (column, chunk of ST_CH rows): 16 independent loads per lane in flight; chunk partials
// land in ppart[chunk][column] / spart[chunk][slot] and are summed (fixed order) by their consumers in phase C.
// Loads that do not depend on the other CTAs (trailing-matrix chunks, panel rows) are issued BEFORE the barrier
// they would otherwise wait behind, so the critical path per column is two barriers and two L2 round trips.
constexpr int ST_CH = 512;
constexpr int ST_PJ = 64;   // max chunks (n <= ST_CH * ST_PJ)
__global__ void __launch_bounds__(ST_NT, 1) k_sytrd_panel(double* A, long long lda, int n, int j0, int nbc, int nb, double* P1,
                                                          double* P2, long long ldp, double* Vh, long long ldv, double* tau,
                                                          double* dd, double* ee, double* ppart, double* spart,
                                                          double* pvpart, unsigned int* flags, unsigned int epoch0, int* err, int dbg) {
  extern __shared__ double sm[];
  double* vs = sm;                 // v, indexed by absolute row
  double* sv = sm + n;             // V'v   (nb)
  double* sw = sv + nb;            // W'v   (nb)
  double* vj = sw + nb;            // V[jn, :]
  double* wj = vj + nb;            // W[jn, :]
  double* pjs = wj + nb;           // chunk partials of p[jn]  (ST_PJ)
  double* red = pjs + ST_PJ;       // ST_NW + 8
  __shared__ int s_to;
  const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int gw = b * ST_NW + warp, GW = G * ST_NW;
  unsigned int epoch = epoch0;
  if (t == 0) s_to = 0;

  for (int i = 0; i < nbc; ++i) {
    const int j = j0 + i, jn = j + 1;
    const int m = n - jn, ncol = m + 2 * i;
    const int nch = (m + ST_CH - 1) / ST_CH;
    const long long nitem = (long long)ncol * nch;
    // ---- prefetch the first phase-B item if it is a chunk of the trailing matrix (independent of the barrier) -----
    double x[16];
    const bool pre = gw < nitem && (gw / nch) < m;
    if (pre) {
      const int cc = gw / nch, q = gw - cc * nch;
      const int rb = jn + q * ST_CH + lane, rend = min(n, jn + (q + 1) * ST_CH);
      const double* col = A + (long long)(jn + cc) * lda;
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? col[rb + 32 * u] : 0.0;
    }
    if (i > 0 && !(dbg & 2)) grid_barrier(flags, ++epoch, G, err, &s_to);   // column j of A is up to date
    // ---- reflector (every CTA, identical arithmetic, norm from the shared-memory copy) ---------------------------
    double ss = 0.0;
    for (int r = jn + t; r < n; r += ST_NT) {
      const double xr = __ldcg(A + r + (long long)j * lda);
      vs[r] = xr;
      if (r > jn) ss += xr * xr;
    }
    const double xn2 = block_sum(ss, red);
    const double alpha = vs[jn];
    double beta, tauj, scale;
    if (xn2 == 0.0) {
      beta = alpha;
      tauj = 0.0;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      tauj = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (int r = jn + t; r < n; r += ST_NT) vs[r] = (r == jn) ? 1.0 : vs[r] * scale;
    __syncthreads();
    for (int r = jn + b * ST_NT + t; r < n; r += G * ST_NT) {
      const double v = vs[r];
      P1[r + (long long)i * ldp] = v;
      P2[r + (long long)(nb + i) * ldp] = v;
      Vh[r + (long long)j * ldv] = v;
    }
    if (b == 0 && t == 0) {
      ee[j] = beta;
      tau[j] = tauj;
      dd[j] = __ldcg(A + j + (long long)j * lda);
    }
    // ---- phase B: chunk partials of p = A22 v and of the panel dots V'v, W'v --------------------------------------
    double pvacc = 0.0;
    for (long long it = gw; it < ((dbg & 1) ? 0 : nitem); it += GW) {
      const int cc = (int)(it / nch), q = (int)(it - (long long)cc * nch);
      const int rb = jn + q * ST_CH + lane, rend = min(n, jn + (q + 1) * ST_CH);
      if (!(pre && it == gw)) {
        if (cc < m) {
          const double* col = A + (long long)(jn + cc) * lda;   // never written inside this kernel
#pragma unroll
          for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? col[rb + 32 * u] : 0.0;
        } else {
          const int kk = cc - m;
          const double* col = P1 + (long long)(kk < i ? kk : nb + (kk - i)) * ldp;   // written by other CTAs: bypass L1
#pragma unroll
          for (int u = 0; u < 16; ++u) x[u] = (rb + 32 * u < rend) ? __ldcg(col + rb + 32 * u) : 0.0;
        }
      }
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int u = 0; u < 16; u += 2) {
        s0 += x[u] * ((rb + 32 * u < rend) ? vs[rb + 32 * u] : 0.0);
        s1 += x[u + 1] * ((rb + 32 * (u + 1) < rend) ? vs[rb + 32 * (u + 1)] : 0.0);
      }
      const double s = warp_sum(s0 + s1);
      if (cc < m) {
        if (lane == 0) ppart[(long long)q * n + jn + cc] = s;
        pvacc += s * vs[jn + cc];
      } else if (lane == 0) {
        const int kk = cc - m;
        spart[q * 2 * nb + (kk < i ? kk : nb + (kk - i))] = s;
      }
    }
    __syncthreads();
    if (lane == 0) red[warp] = pvacc;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int w = 0; w < ST_NW; ++w) s += red[w];
      pvpart[b] = s;
    }
    // ---- prefetch for phase C: this warp's first row of the panels and of column jn (written in earlier columns) --
    const bool next = (i + 1 < nbc);
    const int rfirst = jn + gw;
    double pv_r = 0.0, pw_r = 0.0, pa_r = 0.0;
    if (rfirst < n) {
      if (lane < i) {
        pv_r = __ldcg(P1 + rfirst + (long long)lane * ldp);
        pw_r = __ldcg(P1 + rfirst + (long long)(nb + lane) * ldp);
      }
      if (next) pa_r = __ldcg(A + rfirst + (long long)jn * lda);
    }
    if (!(dbg & 2)) grid_barrier(flags, ++epoch, G, err, &s_to);
    // ---- phase C: w, then column jn of A brought up to date -----------------------------------------------------
    if (t < 2 * i) {
      const int slot = t < i ? t : nb + (t - i);
      double s = 0.0;
      for (int q = 0; q < nch; ++q) s += __ldcg(spart + q * 2 * nb + slot);
      if (t < i) {
        sv[t] = s;
        vj[t] = __ldcg(P1 + jn + (long long)t * ldp);
      } else {
        sw[t - i] = s;
        wj[t - i] = __ldcg(P1 + jn + (long long)(nb + t - i) * ldp);
      }
    } else if (t >= 64 && t < 64 + nch) {
      pjs[t - 64] = __ldcg(ppart + (long long)(t - 64) * n + jn);
    }
    double pv = 0.0;
    for (int q = t; q < G; q += ST_NT) pv += __ldcg(pvpart + q);
    pv = block_sum(pv, red);   // (also orders the sv/sw/vj/wj/pjs stores)
    double cross = 0.0, pcj = 0.0, p_jn = 0.0;
    if (lane < i) {
      cross = sv[lane] * sw[lane];
      pcj = vj[lane] * sw[lane] + wj[lane] * sv[lane];
    }
    cross = warp_sum(cross);
    pcj = warp_sum(pcj);
    for (int q = 0; q < nch; ++q) p_jn += pjs[q];
    const double pcv = pv - 2.0 * cross;
    const double c1 = tauj, c2 = 0.5 * tauj * tauj * pcv;
    const double w_jn = c1 * (p_jn - pcj) - c2;
    for (int r = rfirst; r < ((dbg & 4) ? 0 : n); r += GW) {
      double vr = 0.0, wr = 0.0, ar = 0.0, pr = 0.0;
      for (int q = lane; q < nch; q += 32) pr += __ldcg(ppart + (long long)q * n + r);
      if (r == rfirst) {
        vr = pv_r;
        wr = pw_r;
        ar = pa_r;
      } else {
        if (lane < i) {
          vr = __ldcg(P1 + r + (long long)lane * ldp);
          wr = __ldcg(P1 + r + (long long)(nb + lane) * ldp);
        }
        if (next) ar = __ldcg(A + r + (long long)jn * lda);
      }
      double acc = 0.0, acc2 = 0.0;
      if (lane < i) {
        acc = vr * sw[lane] + wr * sv[lane];
        acc2 = vr * wj[lane] + wr * vj[lane];
      }
      acc = warp_sum(acc);
      acc2 = warp_sum(acc2);
      pr = warp_sum(pr);
      const double w = c1 * (pr - acc) - c2 * vs[r];
      if (lane == 0) {
        P1[r + (long long)(nb + i) * ldp] = w;
        P2[r + (long long)i * ldp] = w;
        if (next) A[r + (long long)jn * lda] = ar - acc2 - (vs[r] * w_jn + w);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// One grid barrier per column (tools/sytrd1b_proto.py is the NumPy statement of the same data flow).
// Index r (row and column) is owned by CTA r mod G; the owner keeps its rows of the panels V, W and of the running
// vectors in shared memory.  Per column ONE exchange carries the matvec partials of A g (column ownership: CTA b
// multiplies its own columns by its own entries of g), the panel dots with g and four scalars; the scalar of the
// previous column that is still unknown when the matvec starts (c_p = tau_p^2 (y_p'v_p)/2) enters afterwards through
//     x = g + 2 c_p v_p,   v = s (x - beta e),   A v = s (A g + 2 c_p (A v_p - A[:, j]) - beta A[:, jn]).
// When the expansion of |x|^2 cancels (sig2 < theta * pieces) the column is redone with c_p folded in (one extra
// barrier) so the backward error stays at the eps level on rank-deficient / graded matrices.
constexpr int S1_NB = 32;   // panel width (lanes <-> panel columns)
constexpr int S1_SS = 72;   // exchanged scalars per CTA: [0,32) V'g, [32,64) W'g, 64 gg, 65 gv, 66 vv, 67 y_p'v_p
// Ownership in groups of 4 consecutive indices (one 32-byte sector of the exchange buffers): group q belongs to CTA
// q mod G; local index l <-> row 4 (b + (l/4) G) + l%4.
__device__ __forceinline__ int s1_row(int b, int G, int l) { return 4 * (b + (l >> 2) * G) + (l & 3); }
__device__ __forceinline__ int s1_first(int b, int G, int x) {   // first local index whose row is >= x
  if (x <= 0) return 0;
  const int gx = x >> 2;
  const int gl = gx > b ? (gx - b + G - 1) / G : 0;
  return (b + gl * G == gx) ? 4 * gl + (x & 3) : 4 * gl;
}
__global__ void __launch_bounds__(ST_NT, 1) k_sytrd_panel1(const double* __restrict__ A, long long lda, int n, int j0, int nbc,
                                                           double* P1, double* P2, long long ldp, double* Vh, long long ldv,
                                                           double* tau, double* dd, double* ee, double* upart, double* spart,
                                                           double* pub, unsigned int* flags, unsigned int epoch0, int* err,
                                                           double theta, int* redo_count, long long* prof, int bar_mode,
                                                           unsigned int ctr_base, int cache_cap) {
  extern __shared__ __align__(16) double sm[];
  const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  unsigned int ctr_target = ctr_base;
  const int NQ = (n + 3) >> 2;                 // row quads
  const int NL = 4 * ((NQ + G - 1) / G);
  double* gs = sm;               // running column g (owned rows)
  double* vps = gs + NL;         // previous reflector v_p
  double* w0ps = vps + NL;       // tau_p y_p
  double* yps = w0ps + NL;       // y_p
  double* avps = yps + NL;       // A v_p
  double* us = avps + NL;        // (A g) summed over CTAs
  double* ajs = us + NL;         // A[r, j]
  double* ajns = ajs + NL;       // A[r, jn]
  double* Vs = ajns + NL;        // NL x 32
  double* Ws = Vs + NL * S1_NB;  // NL x 32
  double* sums = Ws + NL * S1_NB;   // 80
  double* svp = sums + 80;       // V_k'v_p of the previous column, [2][32] (column parity)
  double* swp = svp + 2 * S1_NB; // W_k'v_p, [2][32]
  double* vjn = swp + 2 * S1_NB; // V[jn, :]
  double* wjn = vjn + S1_NB;     // W[jn, :] (final values)
  double* vjr = wjn + S1_NB;     // V[j, :]
  double* wjr = vjr + S1_NB;     // W[j, :]
  double* misc = wjr + S1_NB;    // 24
  double* cache = misc + 24;     // panel-resident copy of this CTA's highest columns of the trailing matrix (TMA)
  __shared__ int s_to;
  __shared__ __align__(8) unsigned long long mbar;
  double* gpub = pub;                   // [2][n]
  double* apub = pub + 2 * (long long)n;
  double* wpub = pub + 4 * (long long)n;
  unsigned int epoch = epoch0, xc = 0;
  const int NLb = s1_first(b, G, n);           // owned rows with r < n
  const bool vec_ok = ((lda & 1) == 0) && ((reinterpret_cast<unsigned long long>(A) & 15ull) == 0);
  // ---- panel-resident column cache: the trailing matrix is constant during the panel, so the columns this CTA
  // multiplies in every matvec are staged ONCE per launch by bulk TMA copies (one elected thread, mbarrier with the
  // expected byte count) and read from shared memory by all 32 columns of the panel.  What does not fit streams
  // from L2 as before.
  const int r_lo = 4 * ((j0 + 1) >> 2);
  const int Lr = n - r_lo;
  const int lc0 = s1_first(b, G, j0 + 1);
  int ncache = (vec_ok && Lr > 0 && (Lr & 1) == 0) ? max(0, min(NLb - lc0, cache_cap / Lr)) : 0;
  if (2 * ncache < NLb - lc0) ncache = 0;      // not worth it unless at least half of the owned columns fit
  const int lcs = NLb - ncache;                 // local columns [lcs, NLb) are cached
  const unsigned int mbar_s = (unsigned int)__cvta_generic_to_shared(&mbar);
  if (t == 0) {
    s_to = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (t == 0 && ncache > 0) {
    const unsigned int bytes = (unsigned int)Lr * 8u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes * (unsigned int)ncache) : "memory");
    for (int q = 0; q < ncache; ++q) {
      const double* src = A + (long long)s1_row(b, G, lcs + q) * lda + r_lo;
      const unsigned int dst = (unsigned int)__cvta_generic_to_shared(cache + (long long)q * Lr);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                   "r"(bytes), "r"(mbar_s)
                   : "memory");
    }
  }
  for (int l = t; l < NL; l += ST_NT) {
    const int r = s1_row(b, G, l);
    gs[l] = (r > j0 && r < n) ? A[r + (long long)j0 * lda] : 0.0;
    vps[l] = w0ps[l] = yps[l] = avps[l] = us[l] = 0.0;
    ajs[l] = (r < n) ? A[r + (long long)j0 * lda] : 0.0;
    ajns[l] = (r < n && j0 + 1 < n) ? A[r + (long long)(j0 + 1) * lda] : 0.0;
  }
  for (int q = t; q < NL * S1_NB; q += ST_NT) Vs[q] = Ws[q] = 0.0;
  if (t < 2 * S1_NB) svp[t] = swp[t] = 0.0;
  if (ncache > 0) {   // all threads: wait for the staged columns (phase 0 of the barrier)
    unsigned int ok = 0;
    while (!ok)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok)
                   : "r"(mbar_s), "r"(0u)
                   : "memory");
  }
  __syncthreads();

  int i = 0;
  bool pending = false;
  double tau_p = 0.0;
  long long tprev = clock64();
#define S1_PROF(slot)                                            \
  if (prof && b == 0 && t == 0) {                                \
    const long long tn = clock64();                              \
    prof[slot] += tn - tprev;                                    \
    tprev = tn;                                                  \
  }
  for (;;) {
    const bool fin = (i == nbc);
    const int j = j0 + i, jn = j + 1;
    const int par = (int)(xc & 1u);
    ++xc;
    const int lmin = s1_first(b, G, jn);        // first owned local row with r >= jn
    const int lminj = s1_first(b, G, j);        // ... with r >= j
    // ---------------- pre-barrier ----------------
    if (!fin) {
      // matvec partial: u_b[r] = sum over owned columns c >= jn of A[r, c] g[c].  A warp takes 128 consecutive rows,
      // lane L the row pairs (2L, 2L+1) and (64+2L, 64+2L+1): every load is a contiguous 512 bytes per warp
      // (coalesced from L2, conflict-free from the shared-memory cache) and the two halves of every 32-byte exchange
      // sector are written by neighbouring lanes.
      double* up = upart + (long long)par * NQ * G * 4;
      if (vec_ok && ncache > 0) {
        const int rbase = 4 * (jn >> 2);
        for (int blk = warp; rbase + blk * 128 < n; blk += ST_NW) {
          const int ra = rbase + blk * 128 + 2 * lane, rb = ra + 64;
          const bool oka = ra < n, okb = rb < n;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          const int lend = min(NLb, lcs);
#pragma unroll 8
          for (int l = lmin; l < lend; ++l) {       // streamed columns
            const double* col = A + (long long)s1_row(b, G, l) * lda;
            const double gl = gs[l];
            if (oka) {
              const double2 x = *reinterpret_cast<const double2*>(col + ra);
              a0 += x.x * gl;
              a1 += x.y * gl;
            }
            if (okb) {
              const double2 x = *reinterpret_cast<const double2*>(col + rb);
              a2 += x.x * gl;
              a3 += x.y * gl;
            }
          }
#pragma unroll 8
          for (int l = max(lmin, lcs); l < NLb; ++l) {   // cached columns
            const double* col = cache + (long long)(l - lcs) * Lr - r_lo;
            const double gl = gs[l];
            if (oka) {
              const double2 x = *reinterpret_cast<const double2*>(col + ra);
              a0 += x.x * gl;
              a1 += x.y * gl;
            }
            if (okb) {
              const double2 x = *reinterpret_cast<const double2*>(col + rb);
              a2 += x.x * gl;
              a3 += x.y * gl;
            }
          }
          if (oka) *reinterpret_cast<double2*>(up + ((long long)(ra >> 2) * G + b) * 4 + (ra & 3)) = make_double2(a0, a1);
          if (okb) *reinterpret_cast<double2*>(up + ((long long)(rb >> 2) * G + b) * 4 + (rb & 3)) = make_double2(a2, a3);
        }
      } else {
        for (int q = (jn >> 2) + t; q < NQ; q += ST_NT) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          const int r0 = 4 * q;
          if (vec_ok && r0 + 3 < n) {      // streamed, one 32-byte row quad per thread
#pragma unroll 8
            for (int l = lmin; l < NLb; ++l) {
              const double2* col = reinterpret_cast<const double2*>(A + (long long)s1_row(b, G, l) * lda + r0);
              const double gl = gs[l];
              const double2 x0 = col[0], x1 = col[1];
              a0 += x0.x * gl;
              a1 += x0.y * gl;
              a2 += x1.x * gl;
              a3 += x1.y * gl;
            }
            double2* dst = reinterpret_cast<double2*>(up + ((long long)q * G + b) * 4);
            dst[0] = make_double2(a0, a1);
            dst[1] = make_double2(a2, a3);
            continue;
          }
          for (int l = lmin; l < NLb; ++l) {
            const double* col = A + (long long)s1_row(b, G, l) * lda + r0;
            const double gl = gs[l];
            a0 += col[0] * gl;
            if (r0 + 1 < n) a1 += col[1] * gl;
            if (r0 + 2 < n) a2 += col[2] * gl;
            if (r0 + 3 < n) a3 += col[3] * gl;
          }
          double2* dst = reinterpret_cast<double2*>(up + ((long long)q * G + b) * 4);
          dst[0] = make_double2(a0, a1);
          dst[1] = make_double2(a2, a3);
        }
      }
      S1_PROF(0)
      for (int l = lminj + t; l < NLb; l += ST_NT) {
        const int r = s1_row(b, G, l);
        gpub[(long long)par * n + r] = gs[l];
        apub[(long long)par * n + r] = avps[l];
        wpub[(long long)par * n + r] = w0ps[l];
      }
    }
    {
      // partial sums over owned rows: 4 lanes per slot (272 threads), two shuffle steps
      const int l1 = s1_first(b, G, jn + 1);      // rows > jn
      double* sp = spart + (long long)par * S1_SS * G;
      const int slot = t >> 2, part = t & 3;
      if (t < 288) {   // warps 0..8 (slots 68..71 are idle lanes that only take part in the shuffles)
        double s = 0.0;
        bool need = slot < 68;
        if (slot >= 68) {
        } else if (slot < 64) {
          const int k = slot & 31;
          need = !fin && k < i;
          if (need) {
            const double* X = (slot < 32) ? Vs : Ws;
            for (int l = lmin + part; l < NLb; l += 4) s += X[l * S1_NB + k] * gs[l];
          }
        } else if (slot == 64) {
          for (int l = l1 + part; l < NLb; l += 4) s += gs[l] * gs[l];
        } else if (slot == 65) {
          for (int l = l1 + part; l < NLb; l += 4) s += gs[l] * vps[l];
        } else if (slot == 66) {
          for (int l = l1 + part; l < NLb; l += 4) s += vps[l] * vps[l];
        } else {
          for (int l = lminj + part; l < NLb; l += 4) s += yps[l] * vps[l];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (need && part == 0) sp[(long long)slot * G + b] = s;
      }
    }
    S1_PROF(1)
    const long long tw0 = (prof && t == 0) ? clock64() : 0;
    if (bar_mode) {
      ctr_target += (unsigned int)G;
      grid_barrier_ctr(flags + 32 * G, ctr_target, err);
    } else {
      grid_barrier(flags, ++epoch, G, err, &s_to);
    }
    if (prof && t == 0) prof[8 + b] += clock64() - tw0;   // per-CTA time inside the barrier (straggler analysis)
    S1_PROF(2)
    // ---------------- post-barrier: gather.  All loads are issued from fully unrolled register arrays before the
    // first reduction, so the phase costs one L2 round trip (G <= 160). ----------------
    {
      const double* sp = spart + (long long)par * S1_SS * G;
      const double* up = upart + (long long)par * NQ * G * 4;
      // round A (all threads): slots 0..63 = V'g, W'g, 8 lanes per slot
      const int slotA = t >> 3, partA = t & 7;
      const bool needA = !fin && (slotA & 31) < i;
      double xa[20];
#pragma unroll
      for (int u = 0; u < 20; ++u) {
        const int q = partA + 8 * u;
        xa[u] = (needA && q < G) ? __ldcg(sp + (long long)slotA * G + q) : 0.0;
      }
      // round B: threads 0..31 the four scalars (8 lanes each), warp 1 u[jn], warps 2.. the owned quads (one warp
      // per quad), the last threads the rows jn / j of the panels and the published scalars
      double xb[20];
      const int glq = (lmin >> 2) + (warp - 2);
      const bool quadw = !fin && warp >= 2 && 4 * glq < NLb;
#pragma unroll
      for (int u = 0; u < 20; ++u) xb[u] = 0.0;
      if (t < 32) {
        const int slotB = 64 + (t >> 3);
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int q = (t & 7) + 8 * u;
          if (q < G) xb[u] = __ldcg(sp + (long long)slotB * G + q);
        }
      } else if (t < 64) {
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int q = lane + 32 * u;
          if (!fin && q < G) xb[u] = __ldcg(up + ((long long)(jn >> 2) * G + q) * 4 + (jn & 3));
        }
      } else if (quadw) {
        const int qd = b + glq * G;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int q = lane + 32 * u;
          if (q < G) {
            const double2* src = reinterpret_cast<const double2*>(up + ((long long)qd * G + q) * 4);
            const double2 y0 = __ldcg(src), y1 = __ldcg(src + 1);
            xb[4 * u] = y0.x;
            xb[4 * u + 1] = y0.y;
            xb[4 * u + 2] = y1.x;
            xb[4 * u + 3] = y1.y;
          }
        }
      }
      if (!fin) {
        const int tt = ST_NT - 1 - t;
        if (tt < i) {
          vjn[tt] = __ldcg(P1 + jn + (long long)tt * ldp);
          vjr[tt] = __ldcg(P1 + j + (long long)tt * ldp);
          const bool stale = pending && tt == i - 1;    // W column i-1 is not final yet: use the published w0_p
          wjn[tt] = stale ? 0.0 : __ldcg(P1 + jn + (long long)(S1_NB + tt) * ldp);
          wjr[tt] = stale ? 0.0 : __ldcg(P1 + j + (long long)(S1_NB + tt) * ldp);
        }
        if (tt == 32) misc[0] = __ldcg(gpub + (long long)par * n + jn);
        if (tt == 33) misc[1] = __ldcg(wpub + (long long)par * n + jn);
        if (tt == 34) misc[2] = __ldcg(wpub + (long long)par * n + j);
        if (tt == 35) misc[3] = __ldcg(apub + (long long)par * n + jn);
        if (tt == 36) misc[4] = A[jn + (long long)j * lda];
        if (tt == 37) misc[5] = A[jn + (long long)jn * lda];
        if (tt == 38) misc[6] = A[j + (long long)j * lda];
        if (tt == 39) misc[7] = (pending && i > 0) ? __ldcg(P1 + jn + (long long)(i - 1) * ldp) : 0.0;   // v_p[jn]
      }
      // reductions
      {
        double sA = 0.0;
#pragma unroll
        for (int u = 0; u < 20; ++u) sA += xa[u];
        sA += __shfl_xor_sync(0xffffffffu, sA, 1);
        sA += __shfl_xor_sync(0xffffffffu, sA, 2);
        sA += __shfl_xor_sync(0xffffffffu, sA, 4);
        if (needA && partA == 0) sums[slotA] = sA;
      }
      if (t < 32) {
        double sB = 0.0;
#pragma unroll
        for (int u = 0; u < 20; ++u) sB += xb[u];
        sB += __shfl_xor_sync(0xffffffffu, sB, 1);
        sB += __shfl_xor_sync(0xffffffffu, sB, 2);
        sB += __shfl_xor_sync(0xffffffffu, sB, 4);
        if ((t & 7) == 0) sums[64 + (t >> 3)] = sB;
      } else if (t < 64) {
        double sB = (xb[0] + xb[1]) + (xb[2] + xb[3]) + xb[4];
        sB = warp_sum(sB);
        if (lane == 0) misc[8] = sB;   // u[jn]
      } else if (quadw) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          s0 += xb[4 * u];
          s1 += xb[4 * u + 1];
          s2 += xb[4 * u + 2];
          s3 += xb[4 * u + 3];
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        s3 = warp_sum(s3);
        if (lane == 0) {
          us[4 * glq] = s0;
          us[4 * glq + 1] = s1;
          us[4 * glq + 2] = s2;
          us[4 * glq + 3] = s3;
        }
      }
      // more owned quads than warps (n > ~8000): the rest, one warp per quad
      if (!fin) {
        for (int gl = (lmin >> 2) + (ST_NW - 2) + warp; 4 * gl < NLb; gl += ST_NW) {
          const int qd = b + gl * G;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          for (int q = lane; q < G; q += 32) {
            const double2* src = reinterpret_cast<const double2*>(up + ((long long)qd * G + q) * 4);
            const double2 y0 = __ldcg(src), y1 = __ldcg(src + 1);
            s0 += y0.x;
            s1 += y0.y;
            s2 += y1.x;
            s3 += y1.y;
          }
          s0 = warp_sum(s0);
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          s3 = warp_sum(s3);
          if (lane == 0) {
            us[4 * gl] = s0;
            us[4 * gl + 1] = s1;
            us[4 * gl + 2] = s2;
            us[4 * gl + 3] = s3;
          }
        }
      }
    }
    if (fin) {
      __syncthreads();
      const double c_l = 0.5 * tau_p * tau_p * sums[67];
      for (int l = lminj + t; l < NLb; l += ST_NT) {
        const int r = s1_row(b, G, l);
        const double w = w0ps[l] - c_l * vps[l];
        P1[r + (long long)(S1_NB + nbc - 1) * ldp] = w;
        P2[r + (long long)(nbc - 1) * ldp] = w;
      }
      break;
    }
    __syncthreads();
    S1_PROF(3)
    // ---------------- global scalars and panel dots, redundantly in every warp (lane k <-> panel column k): no
    // block-wide synchronisation between the gather and the row updates ----------------
    const double gam_p = sums[67], s_gv = sums[65], s_vv = sums[66];
    const double g_jn = misc[0], w0p_jn = misc[1], w0p_j = misc[2], avp_jn = misc[3], A_jn_j = misc[4], A_jn_jn = misc[5],
                 A_j_j = misc[6], vp_jn = misc[7], u_jn = misc[8];
    // one thread does the square root and the division (FP64 issue is per warp instruction: doing this in all
    // 16 warps costs more than the extra __syncthreads)
    if (t == 0) {
      const double gg = sums[64];
      const double cp = pending ? 0.5 * tau_p * tau_p * gam_p : 0.0;
      const double al = g_jn + 2.0 * cp * vp_jn;
      const double sg2 = gg + 4.0 * cp * s_gv + 4.0 * cp * cp * s_vv;
      double be = 0.0, tq = 0.0, sq = 0.0, rd = 0.0;
      if (pending && sg2 < theta * (gg + 4.0 * cp * cp * s_vv)) {
        rd = 1.0;
      } else if (!(sg2 > 0.0)) {
        be = al;
      } else {
        be = -copysign(sqrt(al * al + sg2), al);
        const double amb = al - be;
        const double rr = 1.0 / (be * amb);     // one division: 1/(alpha-beta) = rr*beta, 1/beta = rr*(alpha-beta)
        sq = rr * be;
        tq = -amb * (rr * amb);
      }
      misc[10] = cp;
      misc[11] = be;
      misc[12] = tq;
      misc[13] = sq;
      misc[14] = rd;
    }
    __syncthreads();
    const double c_p = misc[10], beta = misc[11], tj = misc[12], s = misc[13], rdo = misc[14];
    if (rdo != 0.0) {
      // cancellation: fold c_p in, finish W column i-1, redo this column's exchange with nothing pending
      __syncthreads();
      for (int l = lminj + t; l < NLb; l += ST_NT) {
        const int r = s1_row(b, G, l);
        const double w = w0ps[l] - c_p * vps[l];
        Ws[l * S1_NB + (i - 1)] = w;
        P1[r + (long long)(S1_NB + i - 1) * ldp] = w;
        P2[r + (long long)(i - 1) * ldp] = w;
        gs[l] = (r >= jn) ? gs[l] + 2.0 * c_p * vps[l] : 0.0;
      }
      if (b == 0 && t == 0) atomicAdd(redo_count, 1);
      pending = false;
      __syncthreads();
      continue;
    }
    const double* svq = svp + (i & 1) * S1_NB;
    const double* swq = swp + (i & 1) * S1_NB;
    double sv_k = 0.0, sw_k = 0.0, vjn_k = 0.0, wjn_k = 0.0, vjr_k = 0.0, wjr_k = 0.0;
    if (lane < i) {
      const int k = lane;
      vjn_k = vjn[k];
      vjr_k = vjr[k];
      double Vk_vp, Wk_vp, Wk_g;
      if (k < i - 1 || !pending) {
        Vk_vp = svq[k] - vjr_k;
        Wk_vp = swq[k] - wjr[k];
        Wk_g = sums[32 + k];
        wjn_k = wjn[k];
        wjr_k = wjr[k];
      } else {
        const double vv = s_vv + vp_jn * vp_jn, gv = s_gv + g_jn * vp_jn;
        Vk_vp = vv;
        Wk_vp = (tau_p * gam_p - w0p_j) - c_p * vv;
        Wk_g = sums[32 + k] - c_p * gv;
        wjn_k = w0p_jn - c_p * vp_jn;
        wjr_k = w0p_j - c_p;
      }
      const double tv = pending ? 2.0 * c_p * Vk_vp : 0.0, tw = pending ? 2.0 * c_p * Wk_vp : 0.0;
      if (s != 0.0) {
        sv_k = s * (sums[k] + tv - beta * vjn_k);
        sw_k = s * (Wk_g + tw - beta * wjn_k);
      } else {
        sv_k = vjn_k;
        sw_k = wjn_k;
      }
    }
    const double a_d = warp_sum(vjr_k * wjr_k);
    const double q_y = warp_sum(vjn_k * sw_k + wjn_k * sv_k);
    const double Av_jn = (s != 0.0) ? s * (u_jn + 2.0 * c_p * (avp_jn - A_jn_j) - beta * A_jn_jn) : A_jn_jn;
    const double w0_jn = tj * (Av_jn - q_y);
    if (b == 0 && t == 0) {
      dd[j] = A_j_j - 2.0 * a_d;
      ee[j] = beta;
      tau[j] = tj;
    }
    S1_PROF(4)
    // ---------------- owned rows: finish W column i-1, new v / y / w0, next g ----------------
    for (int l = lmin + warp; l < NLb; l += ST_NW) {
      const int r = s1_row(b, G, l);
      double vk = 0.0, wk = 0.0;
      if (lane < i) {
        vk = Vs[l * S1_NB + lane];
        wk = Ws[l * S1_NB + lane];
        if (pending && lane == i - 1) wk = w0ps[l] - c_p * vps[l];
      }
      const double x = gs[l] + 2.0 * c_p * vps[l];
      const double v = (r == jn) ? 1.0 : s * x;
      const double Av = (s != 0.0) ? s * (us[l] + 2.0 * c_p * (avps[l] - ajs[l]) - beta * ajns[l]) : ajns[l];
      double acc = 0.0, acc2 = 0.0;
      double a_next = 0.0;
      if (lane == 1 && jn + 1 < n) a_next = A[r + (long long)(jn + 1) * lda];   // A[r, jn+1] for the next column
      if (lane < i) {
        acc = vk * sw_k + wk * sv_k;
        acc2 = vk * wjn_k + wk * vjn_k;
      }
      acc = warp_sum(acc);
      acc2 = warp_sum(acc2);
      const double y = Av - acc, w0 = tj * y;
      const double gnew = (r > jn) ? (ajns[l] - acc2) - v * w0_jn - w0 : 0.0;
      __syncwarp();
      if (pending && lane == i - 1) {
        Ws[l * S1_NB + lane] = wk;
        P1[r + (long long)(S1_NB + lane) * ldp] = wk;
        P2[r + (long long)lane * ldp] = wk;
      }
      if (lane == 0) {
        Vs[l * S1_NB + i] = v;
        Ws[l * S1_NB + i] = w0;
        vps[l] = v;
        w0ps[l] = w0;
        yps[l] = y;
        avps[l] = Av;
        gs[l] = gnew;
        P1[r + (long long)i * ldp] = v;
        P2[r + (long long)(S1_NB + i) * ldp] = v;
        Vh[r + (long long)j * ldv] = v;
      }
      const double ajn_old = ajns[l];
      __syncwarp();
      if (lane == 1) {
        ajs[l] = ajn_old;
        ajns[l] = a_next;
      }
    }
    if (warp == 0) {
      svp[((i + 1) & 1) * S1_NB + lane] = (lane < i) ? sv_k : 0.0;
      swp[((i + 1) & 1) * S1_NB + lane] = (lane < i) ? sw_k : 0.0;
    }
    tau_p = tj;
    pending = true;
    ++i;
    __syncthreads();
    S1_PROF(6)
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tail of the tridiagonalisation on ONE thread-block cluster.  Once the trailing matrix fits into the shared memory of
// a cluster (t = n - j_start <= 416 with 8 CTAs), a grid-wide exchange per column (~3 L2 round trips, 9 us) is the wrong
// tool: the cluster keeps the whole trailing matrix in distributed shared memory (CTA k owns the columns c = k mod 8),
// exchanges v and the matvec partials through DSMEM and synchronises with two hardware cluster barriers per column.  Unblocked
// dsytd2-style update A -= v w' + w v' directly on the resident matrix (no panels, no trailing GEMM).
constexpr int TC_CS = 8;       // portable cluster size
constexpr int TC_NT = 512;
constexpr int TC_TMAX = 416;   // ceil(t/8) * t doubles of columns + work vectors must fit into 227 KB
__global__ void __launch_bounds__(TC_NT, 1) k_sytrd_tail_cluster(const double* __restrict__ A, long long lda, int n, int j_start,
                                                                 double* Vh, long long ldv, double* tau, double* dd,
                                                                 double* ee, long long* prof) {
  namespace cg = cooperative_groups;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define TC_PROF(slot)                          \
  if (prof) {                                  \
    const long long tn = clock64();            \
    pacc[slot] += tn - tprev;                  \
    tprev = tn;                                \
  }
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double sm[];
  const int t = n - j_start;                 // order of the resident trailing matrix
  const int rank = (int)cluster.block_rank(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nslot = (t + TC_CS - 1) / TC_CS;
  double* cols = sm;                          // [nslot][t]: column c = rank + slot * 8 (tail coordinates)
  double* vbuf = cols + (size_t)nslot * t;    // [2][t]  reflector (parity)
  double* ybuf = vbuf + 2 * t;                // [t]     A v, assembled by the slice owners of all CTAs
  double* ubuf = ybuf + t;                    // [t]     this CTA's matvec partial
  double* wbuf = ubuf + t;                    // [t]
  double* scal = wbuf + t;                    // [2][4] tau, beta (parity); [8..16) gamma partials; [16..) reduction scratch
  for (int idx = tid; idx < nslot * t; idx += TC_NT) {
    const int slot = idx / t, r = idx - slot * t, cl = rank + slot * TC_CS;
    cols[idx] = (cl < t) ? A[(long long)(j_start + r) + (long long)(j_start + cl) * lda] : 0.0;
  }
  for (int idx = tid; idx < 2 * t; idx += TC_NT) vbuf[idx] = 0.0;
  __syncthreads();
  cluster.sync();
  for (int jj = 0; jj + 2 < t; ++jj) {
    const int par = jj & 1, owner = jj % TC_CS, j1 = jj + 1;
    double* v = vbuf + par * t;
    // ---- 1. owner: broadcast the RAW column first (the DSMEM stores drain while the norm, the square root and the
    // division are computed), then the three scalars; every CTA scales its own copy after the barrier ----
    if (rank == owner) {
      const double* x = cols + (size_t)(jj / TC_CS) * t;
      double ss = 0.0;
      for (int r = jj + tid; r < t; r += TC_NT) {
        const double xr = (r < j1) ? 0.0 : x[r];
#pragma unroll
        for (int dst = 0; dst < TC_CS; ++dst) cluster.map_shared_rank(v, dst)[r] = xr;
        if (r > j1) ss += xr * xr;
      }
      ss = warp_sum(ss);
      if (lane == 0) scal[16 + warp] = ss;
      __syncthreads();
      if (tid < TC_CS) {
        double xn2 = 0.0;
        for (int w = 0; w < TC_NT / 32; ++w) xn2 += scal[16 + w];
        const double alpha = x[j1];
        double beta, tj, sc;
        if (!(xn2 > 0.0)) {
          beta = alpha;
          tj = 0.0;
          sc = 0.0;
        } else {
          beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
          const double amb = alpha - beta, rr = 1.0 / (beta * amb);
          sc = rr * beta;
          tj = -amb * (rr * amb);
        }
        double* rs = cluster.map_shared_rank(scal, tid);   // thread k serves CTA k (8 threads repeat the scalar math)
        rs[par * 4 + 0] = tj;
        rs[par * 4 + 1] = beta;
        rs[par * 4 + 2] = sc;
        if (tid == 0) {
          dd[j_start + jj] = x[jj];
          ee[j_start + jj] = beta;
          tau[j_start + jj] = tj;
        }
      }
    }
    TC_PROF(0)
    cluster.sync();   // #1: v, tau in every CTA
    TC_PROF(1)
    const double tj = scal[par * 4 + 0];
    {
      const double sc = scal[par * 4 + 2];
      for (int r = j1 + tid; r < t; r += TC_NT) {
        const double val = (r == j1) ? 1.0 : v[r] * sc;
        v[r] = val;
        if (rank == owner) Vh[(long long)(j_start + r) + (long long)(j_start + jj) * ldv] = val;
      }
      __syncthreads();
    }
    // ---- 2. matvec partial over the owned columns, then every CTA reduces its row slice through DSMEM ----
    {
      const int s0 = (j1 > rank) ? (j1 - rank + TC_CS - 1) / TC_CS : 0;     // first owned slot with column >= j1
      const int s1 = (t - rank + TC_CS - 1) / TC_CS;
      for (int r = j1 + tid; r < t; r += TC_NT) {
        double a0 = 0.0, a1 = 0.0;
        const double* cp = cols + (size_t)s0 * t + r;
        int cl = rank + s0 * TC_CS, slot = s0;
        for (; slot + 1 < s1; slot += 2, cp += 2 * t, cl += 2 * TC_CS) {
          a0 += cp[0] * v[cl];
          a1 += cp[t] * v[cl + TC_CS];
        }
        if (slot < s1) a0 += cp[0] * v[cl];
        ubuf[r] = a0 + a1;
      }
    }
    TC_PROF(2)
    cluster.sync();   // #2: all partials written
    TC_PROF(3)
    {
      // every CTA assembles the full y = A v itself: 8 DSMEM loads per row, all in flight (one row per thread), which
      // saves the third cluster barrier a slice-wise reduction + broadcast would need
      double gpart = 0.0;
      for (int r = j1 + tid; r < t; r += TC_NT) {
        double x0[TC_CS];
#pragma unroll
        for (int src = 0; src < TC_CS; ++src) x0[src] = cluster.map_shared_rank(ubuf, src)[r];
        double y = 0.0;
#pragma unroll
        for (int src = 0; src < TC_CS; ++src) y += x0[src];
        ybuf[r] = y;
        gpart += y * v[r];
      }
      gpart = warp_sum(gpart);
      if (lane == 0) scal[16 + warp] = gpart;
      __syncthreads();
      double gamma = 0.0;
      for (int w = 0; w < TC_NT / 32; ++w) gamma += scal[16 + w];
      const double c2 = 0.5 * tj * tj * gamma;
      for (int r = j1 + tid; r < t; r += TC_NT) wbuf[r] = tj * ybuf[r] - c2 * v[r];
      __syncthreads();
      TC_PROF(4)
      // ---- 3. rank-2 update of the owned columns c >= j1, rows >= j1 ----
      const int s0 = (j1 > rank) ? (j1 - rank + TC_CS - 1) / TC_CS : 0;
      const int s1 = (t - rank + TC_CS - 1) / TC_CS;        // owned slots with a column < t
      for (int r = j1 + tid; r < t; r += TC_NT) {            // one row per thread, v[r], w[r] in registers
        const double vr = v[r], wr = wbuf[r];
        double* cp = cols + (size_t)s0 * t + r;
        int cl = rank + s0 * TC_CS;
#pragma unroll 4
        for (int slot = s0; slot < s1; ++slot, cp += t, cl += TC_CS) *cp -= vr * wbuf[cl] + wr * v[cl];
      }
      __syncthreads();
      TC_PROF(5)
    }
  }
  if (prof && rank == 0 && tid == 0)
    for (int q = 0; q < 6; ++q) prof[q] = pacc[q];
  // the last 2 x 2 block
  {
    const int c2i = t - 2, c1i = t - 1;
    if (t >= 2 && rank == c2i % TC_CS && tid == 0) {
      const double* x = cols + (size_t)(c2i / TC_CS) * t;
      dd[j_start + c2i] = x[c2i];
      ee[j_start + c2i] = x[c1i];
    }
    if (rank == c1i % TC_CS && tid == 0) {
      const double* x = cols + (size_t)(c1i / TC_CS) * t;
      dd[j_start + c1i] = x[c1i];
    }
  }
  cluster.sync();   // nobody may exit while its shared memory can still be addressed remotely
}

__global__ void k_sytrd_tail(const double* A, long long lda, int n, double* dd, double* ee) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (n >= 2) {
      dd[n - 2] = A[(n - 2) + (long long)(n - 2) * lda];
      ee[n - 2] = A[(n - 1) + (long long)(n - 2) * lda];
    }
    dd[n - 1] = A[(n - 1) + (long long)(n - 1) * lda];
  }
}

// dst = src (+ src') for an n x n strided source; dst column-major with leading dimension ldd
__global__ void k_load_symm(double* dst, long long ldd, const double* __restrict__ A, long long s0, long long s1, long long n,
                            int add_t) {
  const long long total = n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % n, j = idx / n;
    double v;
    if (add_t) v = A[i * s0 + j * s1] + A[j * s0 + i * s1];
    else v = 0.5 * (A[i * s0 + j * s1] + A[j * s0 + i * s1]);
    dst[i + j * ldd] = v;
  }
}

// Compact WY factor of one panel (dlarft, forward / columnwise): T upper triangular, H_0..H_{k-1} = I - V T V'.
// G = V'V (k x k, ldg).  T[t][c] = -tau_c sum_{t <= q < c} T[t][q] G[q][c] only involves ROW t of T, so the rows are
// independent: one WARP per row (grid: panels x k/16, 16 warps per CTA), lane l keeps T[t][l + 32 j] in registers, the dot
// product of a column is a 4-term partial sum per lane + one shuffle reduction, and the G column of the next step is
// loaded before the reduction of the current one.  (The column-oriented form -- one barrier pair per column, dot products
// read from global memory inside the dependent loop -- took 245 us per launch, 18 % of an energy + gradient call at n = 80.)
__global__ void __launch_bounds__(512) k_larft(const double* __restrict__ G, int ldg, const double* __restrict__ tau, int k, double* T, int ldt) {
  G += (long long)blockIdx.x * ldg * k;   // G, T are (k, k, panel), tau is (k, panel); k <= 128
  T += (long long)blockIdx.x * ldt * k;
  tau += (long long)blockIdx.x * k;
  const int lane = threadIdx.x & 31, t = blockIdx.y * 16 + (threadIdx.x >> 5);
  if (t >= k) return;
  double tr[4] = {0.0, 0.0, 0.0, 0.0};
  {
    const double tt = tau[t];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j == t) tr[j] = tt;
  }
  double gn[4];
  auto load_col = [&](int c, double (&g)[4]) {
    const double* gc = G + (long long)c * ldg;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = lane + 32 * j;
      g[j] = (c < k && q >= t && q < c) ? __ldg(gc + q) : 0.0;
    }
  };
  load_col(t + 1, gn);
  for (int c = t + 1; c < k; ++c) {
    double g[4] = {gn[0], gn[1], gn[2], gn[3]};
    load_col(c + 1, gn);
    double sum = fma(tr[0], g[0], tr[1] * g[1]) + fma(tr[2], g[2], tr[3] * g[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double v = -__ldg(tau + c) * sum;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j == c) tr[j] = v;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    if (c < k) T[t + (long long)c * ldt] = tr[j];
  }
}

__global__ void k_gather_cols(const double* __restrict__ Q, long long ldq, const int* __restrict__ perm,
                              const double* __restrict__ sgn, long long n, double* __restrict__ U, double* __restrict__ V) {
  // fused: sort by |lambda| (perm), V = U sign(lambda), canonical column sign (largest-magnitude entry positive)
  __shared__ double sh[SIGNFIX_SH];
  const long long j = blockIdx.x;
  const double* src = Q + (long long)perm[j] * ldq;
  const double g = block_canonical_sign(src, n, sh);
  const double s = sgn[j] * g;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = src[i];
    U[i + j * n] = v * g;
    V[i + j * n] = v * s;
  }
}

// H[0:m, m:m+n] = A, H[m:, 0:m] = A' for the (d1,d2) x (d3,d4) strided view A (H zero-initialised, ld = ldh)
__global__ void k_load_jw(double* H, long long ldh, const double* __restrict__ A, long long d1, long long d2, long long d3,
                          long long d4, long long s1, long long s2, long long s3, long long s4) {
  const long long m = d1 * d2, n = d3 * d4, total = m * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % m, j = idx / m;
    const double v = A[(i % d1) * s1 + (i / d1) * s2 + (j % d3) * s3 + (j / d3) * s4];
    H[i + (m + j) * ldh] = v;
    H[(m + j) + i * ldh] = v;
  }
}

// U[:, q] = sqrt(2) Z[0:m, pos[q]], V[:, q] = sqrt(2) Z[m:m+n, pos[q]]
__global__ void k_gather_jw(const double* __restrict__ Z, long long ldz, const int* __restrict__ pos, long long m, long long n,
                            double* __restrict__ U, double* __restrict__ V) {
  __shared__ double sh[SIGNFIX_SH];
  const long long q = blockIdx.x;
  const double* src = Z + (long long)pos[q] * ldz;
  const double r2 = 1.4142135623730951 * block_canonical_sign(src, m, sh);   // canonical gauge: largest |U[:, q]| entry positive
  for (long long i = threadIdx.x; i < m; i += blockDim.x) U[i + q * m] = r2 * src[i];
  for (long long i = threadIdx.x; i < n; i += blockDim.x) V[i + q * n] = r2 * src[m + i];
}

Tens view2(double* p, int64_t rows, int64_t cols, int64_t ld) {
  Tens t = t_wrap(p, {rows, cols});
  t.str[0] = 1;
  t.str[1] = ld;
  return t;
}

}  // namespace

// A (n x n, leading dimension lda, full symmetric storage) is overwritten; Vh (n x n, zero-initialised by the
// caller) receives v_j in column j (rows > j, v[j+1] = 1); tau, dd (n), ee (n-1).
void sytrd(tnad_ctx* c, double* A, int64_t lda, int64_t n, double* Vh, int64_t ldv, double* tau, double* dd, double* ee) {
  TNAD_REQUIRE(n >= 1, "sytrd: empty matrix");
  const int nb = 32;
  cudaStream_t st = c->stream;
  // columns [0, j_tail) go through the grid-wide panel kernel, the rest (trailing order t <= 416) through the cluster kernel
  int64_t j_tail = n;
  if (opt_i(c, "TNAD_SYTRD_TAIL", 1) && n >= 3) {
    j_tail = n > TC_TMAX ? ((n - TC_TMAX + nb - 1) / nb) * nb : 0;
    if (n - j_tail < 3) j_tail = n;
  }
  const int64_t nref = j_tail < n ? j_tail : (n >= 3 ? n - 2 : 0);    // reflector columns done by the panel kernel
  if (nref > 0) {
    const size_t smem = (size_t)(n + 4 * nb + ST_PJ + ST_NW + 16) * sizeof(double);
    TNAD_REQUIRE(n <= (int64_t)ST_CH * ST_PJ, "sytrd: n too large");
    TNAD_REQUIRE(smem <= 220 * 1024, "sytrd: matrix too large for the shared-memory reflector (n <= 28000)");
    TNAD_CUDA(cudaFuncSetAttribute(k_sytrd_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TNAD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sytrd_panel, ST_NT, smem));
    TNAD_REQUIRE(per_sm >= 1, "sytrd: panel kernel does not fit on an SM");
    const int G = c->num_sms;
    const int64_t ldp = n;
    const bool one_barrier = opt_i(c, "TNAD_SYTRD_1B", 1) != 0;
    const int64_t NQ = (n + 3) / 4, NL = 4 * ((NQ + G - 1) / G);
    const size_t smem1 = (size_t)(8 * NL + 2 * NL * S1_NB + 80 + 8 * S1_NB + 24) * sizeof(double);
    Tens upart, spart1, pub, redo, prof;
    size_t smem_total = 0;
    int cache_cap_max = 0;
    double theta = 0.1;
    theta = opt_d(c, "TNAD_SYTRD_THETA", theta);
    if (one_barrier) {
      TNAD_REQUIRE(smem1 <= 200 * 1024, "sytrd: matrix too large");
      smem_total = opt_i(c, "TNAD_SYTRD_CACHE", 1) ? (size_t)(232448 - 256) : smem1;   // everything left of the 227 KB goes to the column cache
      cache_cap_max = (int)((smem_total - smem1) / sizeof(double)) & ~1;
      TNAD_CUDA(cudaFuncSetAttribute(k_sytrd_panel1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
      upart = t_alloc(c, {4 * NQ, (int64_t)G, 2});
      spart1 = t_alloc(c, {S1_SS, (int64_t)G, 2}, true);
      pub = t_alloc(c, {n, 6}, true);
      redo = t_alloc(c, {2}, true);
      prof = t_alloc(c, {8 + (int64_t)G}, true);
    }
    Tens P1 = t_alloc(c, {ldp, 2 * nb}, true), P2 = t_alloc(c, {ldp, 2 * nb}, true);
    const int64_t nchmax = (n + ST_CH - 1) / ST_CH;
    Tens ppart = t_alloc(c, {n, nchmax}), spart = t_alloc(c, {2 * nb, nchmax}), pvpart = t_alloc(c, {G});
    Tens ctl = t_alloc(c, {16 * (int64_t)G + 32}, true);   // one 128-byte flag line per CTA, the counter line, the error flag
    unsigned int* bar = reinterpret_cast<unsigned int*>(ctl.p);
    int* err = reinterpret_cast<int*>(ctl.p + 16 * (int64_t)G + 16);
    unsigned int bar_base = 0;
    int dbg = opt_i(c, "TNAD_SYTRD_DBG", 0);
    int bar_mode = opt_i(c, "TNAD_SYTRD_BAR", 1);   // 1: single release-add counter (measured 3.6k vs 4.4k cycles per barrier), 0: per-CTA flags
    unsigned int ctr_base = 0;
    for (int64_t j0 = 0; j0 < nref; j0 += nb) {
      int nbc = (int)std::min<int64_t>(nb, nref - j0);
      int ni = (int)n, j0i = (int)j0, nbi = nb;
      long long lda_ = lda, ldp_ = ldp, ldv_ = ldv;
      double *P1p = P1.p, *P2p = P2.p, *pp = ppart.p, *sp = spart.p, *pvp = pvpart.p;
      void* args[] = {&A, &lda_, &ni, &j0i, &nbc, &nbi, &P1p, &P2p, &ldp_, &Vh, &ldv_, &tau, &dd, &ee, &pp, &sp, &pvp, &bar, &bar_base, &err, &dbg};
      const double* Ac = A;
      // the column cache takes the whole shared memory (and with it most of L1): only worth it when at least half of
      // a CTA's columns fit, otherwise launch with the small footprint and stream (measured at n = 6400: 231 vs 259 ms)
      const int64_t Lr_p = n - 4 * ((j0 + 1) / 4), ncol_p = 4 * (((n - j0 + 3) / 4 + G - 1) / G);
      const bool use_cache = cache_cap_max > 0 && Lr_p > 0 && 2 * (cache_cap_max / Lr_p) >= ncol_p;
      int cache_cap = use_cache ? cache_cap_max : 0;
      const size_t smem_launch = use_cache ? smem_total : smem1;
      double *upp = upart.p, *spp = spart1.p, *pubp = pub.p;
      int* redop = reinterpret_cast<int*>(redo.p);
      long long* profp = opt_i(c, "TNAD_DC_DEBUG", 0) ? reinterpret_cast<long long*>(prof.p) : nullptr;
      void* args1[] = {&Ac, &lda_, &ni, &j0i, &nbc, &P1p, &P2p, &ldp_, &Vh, &ldv_, &tau, &dd, &ee, &upp, &spp, &pubp, &bar, &bar_base, &err, &theta, &redop, &profp, &bar_mode, &ctr_base, &cache_cap};
      {
        KTimer kt(c, KF_EIG);
        if (one_barrier) TNAD_CUDA(cudaLaunchCooperativeKernel((void*)k_sytrd_panel1, dim3(G), dim3(ST_NT), args1, smem_launch, st));
        else TNAD_CUDA(cudaLaunchCooperativeKernel((void*)k_sytrd_panel, dim3(G), dim3(ST_NT), args, smem, st));
      }
      c->launches++;
      bar_base += (unsigned int)(2 * nbc + 2);   // barrier epochs this launch may consume
      if (bar_mode) {   // the counter variant needs the exact number of barriers of the launch: reset the counter instead
        TNAD_CUDA(cudaMemsetAsync(bar + 32 * G, 0, sizeof(unsigned int), st));
      }
      // trailing update A22 -= [V W] [W V]'
      const int64_t jt = j0 + nbc, nt = n - jt;
      if (nt > 0) {
        Tens A22 = view2(A + jt + jt * lda, nt, nt, lda);
        Tens L = view2(P1.p + jt, nt, 2 * nb, ldp), R = view2(P2.p + jt, nt, 2 * nb, ldp);
        if (nbc < nb) {   // unused panel columns of a short last panel must not contribute
          for (int q = nbc; q < nb; ++q) {
            TNAD_CUDA(cudaMemsetAsync(P1.p + (int64_t)q * ldp, 0, (size_t)n * sizeof(double), st));
            TNAD_CUDA(cudaMemsetAsync(P1.p + (int64_t)(nb + q) * ldp, 0, (size_t)n * sizeof(double), st));
          }
        }
        contract(c, "ik,jk->ij", L, R, A22, -1.0, 1.0);
      }
    }
    int herr = 0;
    TNAD_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    sync(c);
    if (herr) fail(TNAD_ERR_INTERNAL, "sytrd: grid barrier timed out");
    if (one_barrier && opt_i(c, "TNAD_DC_DEBUG", 0)) {
      int rc = 0;
      TNAD_CUDA(cudaMemcpy(&rc, redo.p, sizeof(int), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[tnad dc] sytrd: %d of %lld columns redone (cancellation guard, theta %.3g)\n", rc, (long long)nref, theta);
      std::vector<long long> ph((size_t)8 + G);
      TNAD_CUDA(cudaMemcpy(ph.data(), prof.p, ph.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      {
        std::vector<std::pair<long long, int>> w;
        for (int q = 0; q < G; ++q) w.push_back({ph[(size_t)8 + q], q});
        std::sort(w.begin(), w.end());
        fprintf(stderr, "[tnad dc] sytrd barrier wait per CTA (cycles/column): min %.0f (cta %d), %.0f (cta %d), %.0f (cta %d) ... median %.0f ... max %.0f (cta %d)\n",
                (double)w[0].first / nref, w[0].second, (double)w[1].first / nref, w[1].second, (double)w[2].first / nref, w[2].second,
                (double)w[(size_t)G / 2].first / nref, (double)w[(size_t)G - 1].first / nref, w[(size_t)G - 1].second);
      }
      fprintf(stderr, "[tnad dc] sytrd CTA0 cycles/column: matvec %.0f  sums+publish %.0f  barrier %.0f  gather %.0f  scalars %.0f  warp0 %.0f  rows %.0f\n",
              (double)ph[0] / nref, (double)ph[1] / nref, (double)ph[2] / nref, (double)ph[3] / nref, (double)ph[4] / nref,
              (double)ph[5] / nref, (double)ph[6] / nref);
    }
  }
  if (j_tail < n) {
    const int64_t t = n - j_tail, nslot = (t + TC_CS - 1) / TC_CS;
    const size_t smem = (size_t)(nslot * t + 5 * t + 64) * sizeof(double);
    TNAD_REQUIRE(smem <= 232448 - 256, "sytrd: tail does not fit into the cluster");
    TNAD_CUDA(cudaFuncSetAttribute(k_sytrd_tail_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TC_CS);
    cfg.blockDim = dim3(TC_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TC_CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const double* Ac = A;
    long long lda_ = lda, ldv_ = ldv;
    int ni = (int)n, jt = (int)j_tail;
    Tens tp = t_alloc(c, {8}, true);
    long long* tprof = opt_i(c, "TNAD_DC_DEBUG", 0) ? reinterpret_cast<long long*>(tp.p) : nullptr;
    {
      KTimer kt(c, KF_EIG);
      TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_sytrd_tail_cluster, Ac, lda_, ni, jt, Vh, ldv_, tau, dd, ee, tprof));
    }
    c->launches++;
    if (tprof) {
      long long ph[8];
      sync(c);
      TNAD_CUDA(cudaMemcpy(ph, tp.p, sizeof(ph), cudaMemcpyDeviceToHost));
      const double nc = (double)std::max<int64_t>(1, t - 2);
      fprintf(stderr, "[tnad dc] cluster tail t=%lld cycles/column (rank 0): reflector %.0f  sync1 %.0f  matvec %.0f  sync2 %.0f  assemble+w %.0f  update %.0f\n",
              (long long)t, ph[0] / nc, ph[1] / nc, ph[2] / nc, ph[3] / nc, ph[4] / nc, ph[5] / nc);
    }
  } else {
    k_sytrd_tail<<<1, 32, 0, st>>>(A, lda, (int)n, dd, ee);
    c->launches++;
    TNAD_CUDA(cudaGetLastError());
  }
}

// Number of columns the reflector store Vh (and tau) must have: panels of the back-transformation are read as a
// batch of equal width, so the column count is rounded up (extra columns stay zero, tau = 0).
int64_t sytrd_vcols(int64_t n) { return (n + 127) / 128 * 128; }

// Z[0:n, 0:ncols] <- Q Z with Q = H_0 ... H_{n-3} from sytrd (panels applied last to first, compact WY):
// Gram matrices, T factors and V T of ALL panels come from three batched launches; each panel then costs two GEMMs.
void apply_q(tnad_ctx* c, const double* Vh, int64_t ldv, const double* tau, int64_t n, double* Z, int64_t ldz, int64_t ncols,
             int64_t offset) {
  // reflector of column j: unit element at row j + offset, so the last useful column is n - 2 - offset
  const int64_t nref = n >= offset + 2 ? n - 1 - offset : 0;
  if (nref == 0) return;
  const int kb = opt_i(c, "TNAD_APPLYQ_NB", 128) >= 128 ? 128 : 64;   // measured: 128 wins at n = 2048 (2.5 vs 3.4 ms) and 6400 (46 vs 66 ms)
  const int64_t npan = (nref + kb - 1) / kb;
  TNAD_REQUIRE(npan * kb <= sytrd_vcols(n), "apply_q: reflector store too narrow");
  double* V = const_cast<double*>(Vh);
  Tens Vall;   // (row, panel, column in panel)
  Vall.p = V;
  Vall.rank = 3;
  Vall.dim[0] = n; Vall.dim[1] = npan; Vall.dim[2] = kb;
  Vall.str[0] = 1; Vall.str[1] = (int64_t)kb * ldv; Vall.str[2] = ldv;
  Tens Gall = t_alloc(c, {kb, kb, npan}), Tall = t_alloc(c, {kb, kb, npan}), VT = t_alloc(c, {n, kb, npan});
  contract(c, "rpi,rpj->ijp", Vall, Vall, Gall);
  k_larft<<<dim3((unsigned)npan, (unsigned)((kb + 15) / 16)), 512, 0, c->stream>>>(Gall.p, kb, tau, kb, Tall.p, kb);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  contract(c, "rpi,ijp->rjp", Vall, Tall, VT);
  Tens Y = t_alloc(c, {kb, ncols});
  for (int64_t pi = npan - 1; pi >= 0; --pi) {
    const int64_t j0 = pi * kb, jr = j0 + offset, rows = n - jr;
    Tens Vp = view2(V + jr + j0 * ldv, rows, kb, ldv);
    Tens VTp = view2(VT.p + jr + pi * (int64_t)kb * n, rows, kb, n);
    Tens Zr = view2(Z + jr, rows, ncols, ldz);
    contract(c, "ki,kj->ij", Vp, Zr, Y);
    contract(c, "ik,kj->ij", VTp, Y, Zr, -1.0, 1.0);
  }
}

// Eigen-decomposition route in three phases (the split lets several GPUs share the back-transformation, see
// tnad_symeig_* in capi.cu and tensornetworkad.jl_b200/sharded.py):
//   reduce          A -> tridiagonal T (one- or two-stage) and the eigen-decomposition of T by divide and conquer
//   backtransform   columns [col0, col0 + ncols) of Z <- Q Z   (independent per column)
//   finish          sort by |lambda|, V = U sign(lambda), canonical column signs
void symeig_reduce(tnad_ctx* c, Tens& Aw, int64_t n, EigFactor& f, bool want_q) {
  cudaStream_t st = c->stream;
  // scale to max|a_ij| = 1 (the reflector norms are plain sums of squares; LAPACK rescales inside dlarfg instead)
  double amax = 0.0;
  {
    Tens sc = t_alloc(c, {2});
    reduce(c, RED_ABSMAX, Aw, nullptr, sc.p);
    d2h(c, &amax, sc.p, 1);
    if (amax > 0.0 && std::isfinite(amax)) scale_dev(c, Aw, Aw, sc.p, SC_INV);
    else amax = 1.0;
  }
  f.n = n;
  f.Vh = t_alloc(c, {n, sytrd_vcols(n)}, true);
  f.tau = t_alloc(c, {sytrd_vcols(n)}, true);
  Tens dd = t_alloc(c, {n}), ee = t_alloc(c, {n}, true);
  const bool debug = opt_i(c, "TNAD_DC_DEBUG", 0) != 0;
  // two-stage route (band.cu) from n >= TNAD_2STAGE_MIN on; TNAD_EIG_2STAGE = 0 | 1 forces one route
  const int ts_mode = opt_i(c, "TNAD_EIG_2STAGE", -1);
  f.two_stage = n >= 67 && (ts_mode == 1 || (ts_mode < 0 && n >= opt_i(c, "TNAD_2STAGE_MIN", 512)));
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (debug)
    for (auto& e : ev) e = get_event(c);
  if (debug) TNAD_CUDA(cudaEventRecord(ev[0], st));
  Tens lam;
  if (f.two_stage) {
    const int64_t NP = chase_positions(n);
    f.ldv2 = 32 * NP;
    sy2sb(c, Aw.p, n, n, f.Vh.p, n, f.tau.p);                      // A = Q1 B Q1'
    Tens AB = t_alloc(c, {33, n});
    extract_band(c, Aw.p, n, n, AB.p, 33);
    if (debug) TNAD_CUDA(cudaEventRecord(ev[1], st));
    f.V2 = t_alloc(c, {f.ldv2, n - 2}, true);
    f.tau2 = t_alloc(c, {NP, n - 2}, true);
    // Explicit-Q mode: the chase (one cluster, latency bound) and most of the divide and conquer leave the device idle.
    // Q1 = (stage-1 reflectors) I is formed on stream2 while the chase runs, Q2 = (chase reflectors) I and Qfull = Q1 Q2
    // while the divide and conquer runs; the back-transformation is then a single product Qfull Z instead of the
    // reflector passes (n = 2048: 4.8 ms of back-transformation after the divide and conquer become about 2 ms).
    const int xq = opt_i(c, "TNAD_EXPLICIT_Q", 2);
    const bool explicit_q = want_q && xq > 0 && c->stream2 && n <= opt_i(c, "TNAD_EXPLICITQ_MAX", 4096);
    cudaStream_t side = c->stream2;
    cudaEvent_t e_band = nullptr, e_chase = nullptr;
    if (explicit_q) {
      f.Q1x = t_alloc(c, {n, n});
      f.Q2x = t_alloc(c, {n, n});
      f.Qfull = t_alloc(c, {n, n});
      e_band = get_event(c);
      e_chase = get_event(c);
      f.q_ready = get_event(c);
      TNAD_CUDA(cudaEventRecord(e_band, st));      // stage-1 reflectors and the three buffers exist
    }
    auto on_side = [&](auto&& body) {              // enqueue `body` on stream2 (allocations inside follow that stream)
      c->stream = side;
      try {
        body();
      } catch (...) {
        c->stream = st;
        c->gemm_grid_cap = 0;
        throw;
      }
      c->stream = st;
      c->gemm_grid_cap = 0;
    };
    std::function<void(int)> overlap;
    if (explicit_q)
      overlap = [&](int chase_ctas) {
        on_side([&] {
          TNAD_CUDA(cudaStreamWaitEvent(side, e_band, 0));
          c->gemm_grid_cap = std::max(8, c->num_sms - chase_ctas);   // the chase keeps its SMs: persistent grids stay off them
          set_identity(c, f.Q1x.p, n, n);
          apply_q(c, f.Vh.p, n, f.tau.p, n, f.Q1x.p, n, n, 32);
        });
      };
    sb2st(c, AB.p, 33, n, dd.p, ee.p, f.V2.p, f.ldv2, f.tau2.p, overlap);   // B = Q2 T Q2'
    if (explicit_q) {
      TNAD_CUDA(cudaEventRecord(e_chase, st));
      on_side([&] {
        TNAD_CUDA(cudaStreamWaitEvent(side, e_chase, 0));
        set_identity(c, f.Q2x.p, n, n);
        apply_q2(c, f.V2.p, f.ldv2, f.tau2.p, n, f.Q2x.p, n, n, opt_i(c, "TNAD_Q2_IDENT", 1) != 0);
        contract(c, "ik,kj->ij", f.Q1x, f.Q2x, f.Qfull);
        TNAD_CUDA(cudaEventRecord(f.q_ready, side));
      });
      c->event_pool.push_back(e_band);
      c->event_pool.push_back(e_chase);
    }
  } else {
    sytrd(c, Aw.p, n, n, f.Vh.p, n, f.tau.p, dd.p, ee.p);
    if (debug) TNAD_CUDA(cudaEventRecord(ev[1], st));
  }
  if (debug) TNAD_CUDA(cudaEventRecord(ev[2], st));
  stedc(c, dd.p, ee.p, n, lam, f.Z, f.N);   // Z: N x N (ld N), lam: N, pads carry eigenvalues above the spectrum
  if (debug) {
    TNAD_CUDA(cudaEventRecord(ev[3], st));
    TNAD_CUDA(cudaEventSynchronize(ev[3]));
    float t[3];
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
    if (f.two_stage)
      fprintf(stderr, "[tnad dc] n=%lld N=%lld two-stage: sy2sb %.2f ms  chase %.2f ms  stedc %.2f ms", (long long)n, (long long)f.N, t[0], t[1], t[2]);
    else
      fprintf(stderr, "[tnad dc] n=%lld N=%lld sytrd %.2f ms  stedc %.2f ms", (long long)n, (long long)f.N, t[0] + t[1], t[2]);
    for (auto& e : ev) c->event_pool.push_back(e);
  }
  f.lam.resize((size_t)f.N);
  d2h(c, f.lam.data(), lam.p, (size_t)f.N);
  for (auto& v : f.lam) v *= amax;
}

// Zc (N x ncols, leading dimension ldz) holds columns col0.. of the tridiagonal eigenvectors on entry, of Q times them on exit
void symeig_backtransform(tnad_ctx* c, const EigFactor& f, double* Zc, int64_t ldz, int64_t ncols) {
  const bool debug = opt_i(c, "TNAD_DC_DEBUG", 0) != 0;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (debug) {
    for (auto& e : ev) e = get_event(c);
    TNAD_CUDA(cudaEventRecord(ev[0], c->stream));
  }
  if (f.Qfull.p) {
    // explicit-Q mode: one product, through a temporary (the reflector passes work in place)
    TNAD_CUDA(cudaStreamWaitEvent(c->stream, f.q_ready, 0));
    if (debug) TNAD_CUDA(cudaEventRecord(ev[1], c->stream));
    Tens X = view2(Zc, f.n, ncols, ldz);
    Tens Y = contract_new(c, "ik,kj->ij", f.Qfull, X);
    tcopy(c, Y, X);
  } else if (f.two_stage) {
    apply_q2(c, f.V2.p, f.ldv2, f.tau2.p, f.n, Zc, ldz, ncols);
    if (debug) TNAD_CUDA(cudaEventRecord(ev[1], c->stream));
    apply_q(c, f.Vh.p, f.n, f.tau.p, f.n, Zc, ldz, ncols, 32);
  } else {
    if (debug) TNAD_CUDA(cudaEventRecord(ev[1], c->stream));
    apply_q(c, f.Vh.p, f.n, f.tau.p, f.n, Zc, ldz, ncols);
  }
  if (debug) {
    TNAD_CUDA(cudaEventRecord(ev[2], c->stream));
    TNAD_CUDA(cudaEventSynchronize(ev[2]));
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    fprintf(stderr, "  apply_q2 %.2f ms  apply_q1 %.2f ms  (%lld columns)\n", a, b, (long long)ncols);
    for (auto& e : ev) c->event_pool.push_back(e);
  }
}

// Zfull: the back-transformed N x N eigenvector matrix (ld N).  U, S, V as LinearAlgebra.svd returns them for a symmetric matrix.
SvdResult symeig_finish(tnad_ctx* c, const EigFactor& f, const double* Zfull, int64_t ldz) {
  const int64_t n = f.n, N = f.N;
  if (ldz <= 0) ldz = N;
  cudaStream_t st = c->stream;
  const std::vector<double>& lh = f.lam;
  // the N - n pad eigenvalues are the largest ones (stedc puts them above 3 |T|)
  std::vector<int> idx((size_t)N);
  for (int64_t i = 0; i < N; ++i) idx[(size_t)i] = (int)i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lh[x] < lh[y]; });
  std::vector<int> perm(idx.begin(), idx.begin() + n);
  std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return std::fabs(lh[x]) > std::fabs(lh[y]); });
  std::vector<double> sval((size_t)n), sgn((size_t)n);
  double fro2 = 0.0;
  for (int64_t r = 0; r < n; ++r) {
    sval[r] = std::fabs(lh[perm[r]]);
    sgn[r] = lh[perm[r]] < 0.0 ? -1.0 : 1.0;
    fro2 += sval[r] * sval[r];
  }
  Tens meta = t_alloc(c, {3 * n + 4});
  int* dperm = reinterpret_cast<int*>(meta.p);
  double* dsgn = meta.p + (n + 1) / 2 + 1;
  double* dsval = dsgn + n;
  TNAD_CUDA(cudaMemcpyAsync(dperm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  TNAD_CUDA(cudaMemcpyAsync(dsgn, sgn.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
  TNAD_CUDA(cudaMemcpyAsync(dsval, sval.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
  SvdResult res;
  res.U = t_alloc(c, {n, n});
  res.V = t_alloc(c, {n, n});
  res.S = t_alloc(c, {n});
  k_gather_cols<<<(int)n, 128, 0, st>>>(Zfull, ldz, dperm, dsgn, n, res.U.p, res.V.p);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  TNAD_CUDA(cudaMemcpyAsync(res.S.p, dsval, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  sync(c);
  res.s_host = sval;
  res.null_thr = 16.0 * 2.220446049250313e-16 * std::sqrt(fro2);   // same absolute level as the Jacobi solver
  res.sweeps = 0;
  return res;
}

void load_symmetric(tnad_ctx* c, const Tens& A, bool sym_add_transpose, Tens& Aw) {
  TNAD_REQUIRE(A.rank == 2 && A.dim[0] == A.dim[1], "symmetric eigensolver: need a square matrix");
  const int64_t n = A.dim[0];
  Aw = t_alloc(c, {n, n});
  const int nbk = (int)std::max<long long>(1, std::min<long long>((n * n + 1023) / 1024, 148 * 8));
  k_load_symm<<<nbk, 256, 0, c->stream>>>(Aw.p, n, A.p, A.str[0], A.str[1], n, sym_add_transpose ? 1 : 0);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
}

// Eigen-decomposition route: tridiagonalise, divide and conquer, back-transform.
SvdResult svd_symmetric_dc(tnad_ctx* c, const Tens& A, bool sym_add_transpose) {
  Tens Aw;
  load_symmetric(c, A, sym_add_transpose, Aw);
  EigFactor f;
  symeig_reduce(c, Aw, A.dim[0], f, true);
  if (f.Qfull.p) {
    // U0 = Qfull Z[0:n, :] straight into a fresh buffer (leading dimension n); the pad rows of Z are not needed
    const bool debug = opt_i(c, "TNAD_DC_DEBUG", 0) != 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (debug) {
      for (auto& e : ev) e = get_event(c);
      TNAD_CUDA(cudaEventRecord(ev[0], c->stream));
    }
    TNAD_CUDA(cudaStreamWaitEvent(c->stream, f.q_ready, 0));
    Tens X = view2(f.Z.p, f.n, f.N, f.N);
    Tens U0 = contract_new(c, "ik,kj->ij", f.Qfull, X);
    if (debug) {
      TNAD_CUDA(cudaEventRecord(ev[1], c->stream));
      TNAD_CUDA(cudaEventSynchronize(ev[1]));
      float a = 0;
      cudaEventElapsedTime(&a, ev[0], ev[1]);
      fprintf(stderr, "  wait for Qfull + Qfull * Z %.2f ms\n", a);
      for (auto& e : ev) c->event_pool.push_back(e);
    }
    SvdResult r = symeig_finish(c, f, U0.p, f.n);
    c->event_pool.push_back(f.q_ready);
    return r;
  }
  symeig_backtransform(c, f, f.Z.p, f.N, f.N);
  return symeig_finish(c, f, f.Z.p);
}

// General m x n matrix through the Jordan-Wielandt embedding H = [0 A; A' 0] (order m + n): the eigenpairs of H are
// (+-sigma_i, [u_i; +-v_i]/sqrt(2)), so the positive eigenvalues above the noise level give (U_r, S_r, V_r) with the
// absolute accuracy eps |A| of a LAPACK SVD.  The null triplets are NOT formed (the eigenvectors of the 2(n - r)-fold
// zero eigenvalue mix left and right null vectors): rank_left = rank_right = r tells svd_back to apply their
// contribution through the projectors I - U_r U_r', I - V_r V_r' (DESIGN.md 4.3).
SvdResult svd_general_dc(tnad_ctx* c, const Tens& A4) {
  // A4: rank-2 (m x n) or rank-4 [(d1,d2),(d3,d4)] strided view
  TNAD_REQUIRE(A4.rank == 2 || A4.rank == 4, "svd_general_dc: need a rank-2 or rank-4 view");
  Tens v = A4;
  if (v.rank == 2) {
    v.rank = 4;
    v.dim[3] = 1; v.str[3] = 0;
    v.dim[2] = A4.dim[1]; v.str[2] = A4.str[1];
    v.dim[1] = 1; v.str[1] = 0;
  }
  const int64_t m = v.dim[0] * v.dim[1], n = v.dim[2] * v.dim[3], nh = m + n, kk = std::min(m, n);
  cudaStream_t st = c->stream;
  Tens H = t_alloc(c, {nh, nh}, true);
  {
    const long long total = m * n;
    const int nbk = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16));
    k_load_jw<<<nbk, 256, 0, st>>>(H.p, nh, v.p, v.dim[0], v.dim[1], v.dim[2], v.dim[3], v.str[0], v.str[1], v.str[2], v.str[3]);
    c->launches++;
    TNAD_CUDA(cudaGetLastError());
  }
  // reduce H to tridiagonal form and solve the tridiagonal problem; only the eigenvectors of the POSITIVE eigenvalues
  // above the noise level are needed (eigenvalues come in +-sigma pairs), so only those columns are back-transformed:
  // half of the 4 n^3 flop of the back-transformation
  EigFactor f;
  symeig_reduce(c, H, nh, f);
  const std::vector<double>& lh = f.lam;
  const int64_t N = f.N;
  // real eigenvalues = the nh smallest (pads are above the spectrum); positive ones, descending
  std::vector<int> idx((size_t)N);
  for (int64_t i = 0; i < N; ++i) idx[(size_t)i] = (int)i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lh[x] < lh[y]; });
  idx.resize((size_t)nh);
  double fro2 = 0.0;
  for (int i : idx) fro2 += 0.5 * lh[i] * lh[i];
  const double thr = 16.0 * 2.220446049250313e-16 * std::sqrt(fro2);
  std::vector<int> pos;
  for (auto it = idx.rbegin(); it != idx.rend() && (int64_t)pos.size() < kk; ++it)
    if (lh[*it] > thr) pos.push_back(*it);
  const int64_t r = (int64_t)pos.size();
  std::vector<double> sval((size_t)kk, 0.0);
  for (int64_t i = 0; i < r; ++i) sval[(size_t)i] = lh[pos[(size_t)i]];
  SvdResult res;
  res.U = t_alloc(c, {m, kk}, true);
  res.V = t_alloc(c, {n, kk}, true);
  res.S = t_alloc(c, {kk});
  h2d(c, res.S.p, sval.data(), (size_t)kk);
  if (r > 0) {
    // selected columns of the tridiagonal eigenvector matrix, back-transformed in place, then split into U and V
    const int64_t rc = (r + 1) & ~1LL;
    Tens Zsel = t_alloc(c, {N, rc}, true);
    for (int64_t q = 0; q < r; ++q)
      TNAD_CUDA(cudaMemcpyAsync(Zsel.p + q * N, f.Z.p + (int64_t)pos[(size_t)q] * N, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    symeig_backtransform(c, f, Zsel.p, N, rc);
    std::vector<int> ident((size_t)r);
    for (int64_t q = 0; q < r; ++q) ident[(size_t)q] = (int)q;
    Tens meta = t_alloc(c, {r / 2 + 2});
    int* dpos = reinterpret_cast<int*>(meta.p);
    TNAD_CUDA(cudaMemcpyAsync(dpos, ident.data(), (size_t)r * sizeof(int), cudaMemcpyHostToDevice, st));
    k_gather_jw<<<(int)r, 128, 0, st>>>(Zsel.p, N, dpos, m, n, res.U.p, res.V.p);
    c->launches++;
    TNAD_CUDA(cudaGetLastError());
    sync(c);      // ident / Zsel must outlive the kernel
  }
  sync(c);
  res.s_host = sval;
  res.null_thr = thr;
  res.rank_left = r;
  res.rank_right = r;
  res.sweeps = 0;
  return res;
}

SvdResult svd_symmetric_auto(tnad_ctx* c, const Tens& A, bool sym_add_transpose, const Tens* Q0) {
  const int mode = opt_i(c, "TNAD_SYMEIG", -1);
  const int64_t n = A.dim[0];
  // the direct solver needs cooperative (co-resident) launches; a device / partition without them keeps the Jacobi path
  const int coop = c->coop_launch;
  TNAD_REQUIRE(coop || mode != 2, "TNAD_SYMEIG=2 needs cooperative kernel launches, which this device does not support");
  const bool dc = coop && (mode == 2 || (mode < 0 && n >= opt_i(c, "TNAD_DC_MIN", 48)));
  return dc ? svd_symmetric_dc(c, A, sym_add_transpose) : svd_symmetric(c, A, sym_add_transpose, Q0);
}

}  // namespace tnad
