// Two-stage tridiagonalisation of the symmetric eigensolver (CTMRG svd(cpmat + cpmat'), ctmrg.jl:134-136; and through
// the Jordan-Wielandt embedding the TRG splits, trg.jl:35-36):
//
//   stage 1  sy2sb     A = Q1 B Q1'   dense -> band (half bandwidth 32): Householder QR of every 32-column sub-band panel
//                                     inside ONE thread-block cluster (panel resident in distributed shared memory, one
//                                     hardware cluster barrier per column, T factor built on the fly), trailing matrix
//                                     updated by two DMMA GEMMs per panel (Z = A22 Y, A22 -= [Y W][W Y]')
//   stage 2  sb2st     B = Q2 T Q2'   band -> tridiagonal by bulge chasing, organised as a systolic array: position t
//                                     (one warp) owns the 32 x 64 window [E | D] of rows s+1+32t .. s+32+32t, which slides
//                                     down the band by one row per sweep; reflectors travel down the array, the rows that
//                                     enter / leave a window travel up, both through seq-tagged 8-byte mailbox words
//                                     (shared memory inside a CTA, L2 between CTAs) -- no grid barrier, no atomics
//   back     U = Q1 (Q2 E)            Q2: register-resident systolic pass (slot q keeps rows s+1+32q .. of 2 columns per
//                                     thread in registers, window slides up one row per sweep, rows are handed from slot
//                                     to slot through shared memory, from stage to stage through a stream in HBM/L2);
//                                     Q1: compact-WY GEMMs (apply_q of tridiag.cu with reflector offset 32)
//
// tools/twostage_proto.py states the same data flow in NumPy (sb2st_systolic, apply_q2_systolic) and is tested on the CPU.
#include "eigdc.h"
#include <cooperative_groups.h>
#include <algorithm>

namespace cg = cooperative_groups;

namespace tnad {

namespace {

constexpr int CB = 32;     // half bandwidth = reflector length of the chase
constexpr int WLD = 33;    // leading dimension of the window arrays in shared memory

#define LAUNCH_CHECK(c)            \
  do {                             \
    (c)->launches++;               \
    TNAD_CUDA(cudaGetLastError()); \
  } while (0)

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- lower band storage from the reduced dense matrix ---------------------------------------------------------------
// AB[(i - j) + j * ldab] = A[i, j] for 0 <= i - j <= 32
__global__ void k_extract_band(const double* __restrict__ A, long long lda, int n, double* __restrict__ AB, int ldab) {
  const long long total = (long long)n * (CB + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / (CB + 1)), o = (int)(idx - (long long)j * (CB + 1));
    const int i = j + o;
    AB[o + (long long)j * ldab] = i < n ? A[i + (long long)j * lda] : 0.0;
  }
}

// =====================================================================================================================
// stage 2: bulge chasing as a systolic array
// =====================================================================================================================
// Mailbox words: (seq << 32) | 32 payload bits; a double travels as two words.  Every 8-byte store is single-copy atomic,
// so the receiver needs no fence: it polls until both words of its double carry the expected sequence number.
// Single-slot boxes are enough: position t sends v(s+1) only after it consumed row(s) from t+1, which t+1 sent after it
// consumed v(s); the same argument holds for the rows.
constexpr int MB_WORDS = 2 * (CB + 1);     // 33 doubles per message

__device__ __forceinline__ void mb_send(unsigned long long* box, unsigned seq, int idx, double val) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(val);
  volatile unsigned long long* vb = box;
  vb[2 * idx] = ((unsigned long long)seq << 32) | (b & 0xffffffffull);
  vb[2 * idx + 1] = ((unsigned long long)seq << 32) | (b >> 32);
}

// The whole warp receives a 33-double message into dst (shared memory).  false on time-out (sets *err).
__device__ __forceinline__ bool mb_recv(unsigned long long* box, unsigned seq, double* dst, int lane, int* err) {
  volatile unsigned long long* vb = box;
  unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  long long t0 = 0;
  unsigned spins = 0;
  for (;;) {
    w0 = vb[2 * lane];
    w1 = vb[2 * lane + 1];
    bool ok = (unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq;
    if (lane == 0) {
      w2 = vb[2 * CB];
      w3 = vb[2 * CB + 1];
      ok = ok && (unsigned)(w2 >> 32) == seq && (unsigned)(w3 >> 32) == seq;
    }
    if (__all_sync(0xffffffffu, ok)) break;
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      bool to = (now - t0 > 4000000000ll) || (*(volatile int*)err != 0);
      if (__any_sync(0xffffffffu, to)) {
        if (lane == 0) atomicExch(err, 1);
        return false;
      }
    }
  }
  dst[lane] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  if (lane == 0) dst[CB] = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  __syncwarp();
  return true;
}

// Householder reflector of the vector whose element `lane` is x (dlarfg): returns v_lane; tau, beta through references.
__device__ __forceinline__ double warp_house(double x, int lane, double& tau, double& beta) {
  const double xn2 = wsum(lane >= 1 ? x * x : 0.0);
  const double alpha = __shfl_sync(0xffffffffu, x, 0);
  if (xn2 == 0.0) {
    tau = 0.0;
    beta = alpha;
    return lane == 0 ? 1.0 : 0.0;
  }
  // one reciprocal square root and one reciprocal instead of a square root and two divisions (all three are software
  // sequences on the critical path of the pipeline): beta = -sign(alpha) |x|, tau = (beta - alpha) / beta = 1 + |alpha| / |x|,
  // v = x / (alpha - beta) = sign(alpha) x / (|alpha| + |x|)
  const double s2 = fma(alpha, alpha, xn2);
  const double rn = rsqrt(s2);
  const double nrm = s2 * rn;
  const double aa = fabs(alpha);
  beta = alpha >= 0.0 ? -nrm : nrm;
  tau = fma(aa, rn, 1.0);
  const double rc = 1.0 / (aa + nrm);
  const double scale = alpha >= 0.0 ? rc : -rc;
  return lane == 0 ? 1.0 : x * scale;
}

struct ChaseArgs {
  const double* AB;
  int ldab;
  int n, NP, W;                  // W positions (warps) per CTA
  double* d;
  double* e;
  double* V2;                    // (n-2) x ldv, zero-initialised
  long long ldv;
  double* tau2;                  // (n-2) x NP, zero-initialised
  unsigned long long* gbox;      // NP x 2 x MB_WORDS, zero-initialised: [t][0] = reflector box, [t][1] = row box
  int* err;
  long long* prof;               // optional: per position 6 cycle counters (wait row, wait v, E phase, D phase, sends, total)
};

// per-position shared memory (doubles): Ew (1 + 32*33 + 1), Dw (32*33), 6 scratch vectors of 34, two mailboxes
constexpr int POS_DOUBLES = (2 + CB * WLD) + CB * WLD + 6 * 34 + 2 * MB_WORDS;

__global__ void __launch_bounds__(256, 1) k_chase(const ChaseArgs a) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // zero the local mailboxes of every position of this CTA before any neighbour may write into them
  for (int i = threadIdx.x; i < a.W * POS_DOUBLES; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  cluster.sync();
  const int t = blockIdx.x * a.W + wid;
  const int n = a.n;
  if (t < a.NP) {
  double* base = sm + (size_t)wid * POS_DOUBLES;
  double* Ew = base + 1;                       // one double of head room: the shifted store of row 1 touches Ew[-1]
  double* Dw = base + 2 + CB * WLD;
  double* sv = Dw + CB * WLD;                  // own reflector (34)
  double* svp = sv + 34;                       // previous reflector of the sweep + tau (34)
  double* srow = svp + 34;                     // entering row (34)
  double* sw = srow + 34;                      // E-phase coefficients w_k (34)
  double* sd = sw + 34;                        // D-phase vector w (34)
  double* stop = sd + 34;                      // leaving top row (34)
  unsigned long long* lbox = reinterpret_cast<unsigned long long*>(stop + 34);   // [0] reflector box, [1] row box (local)
  // A mailbox lives with its READER: in the reader's shared memory when the writer sits in the same CTA or in the same
  // cluster (then the writer pushes through distributed shared memory), in global memory (L2) between clusters.
  constexpr int BOXOFF = POS_DOUBLES - 2 * MB_WORDS;   // offset of a position's local boxes inside its shared-memory slice
  const bool prev_cta = t > 0 && wid == 0, next_cta = t + 1 < a.NP && wid + 1 == a.W;
  const bool prev_far = prev_cta && crank == 0, next_far = next_cta && crank + 1 == csize;   // neighbour in another cluster
  unsigned long long* my_vbox = prev_far ? a.gbox + ((size_t)t * 2 + 0) * MB_WORDS : lbox;
  unsigned long long* my_rbox = next_far ? a.gbox + ((size_t)t * 2 + 1) * MB_WORDS : lbox + MB_WORDS;
  unsigned long long* nx_vbox;   // reflector box of position t+1
  if (!next_cta) nx_vbox = reinterpret_cast<unsigned long long*>(base + POS_DOUBLES + BOXOFF);
  else if (next_far) nx_vbox = a.gbox + ((size_t)(t + 1) * 2 + 0) * MB_WORDS;
  else nx_vbox = reinterpret_cast<unsigned long long*>(cluster.map_shared_rank(sm + BOXOFF, crank + 1));
  unsigned long long* pv_rbox;   // row box of position t-1
  if (!prev_cta) pv_rbox = reinterpret_cast<unsigned long long*>(base - POS_DOUBLES + BOXOFF) + MB_WORDS;
  else if (prev_far) pv_rbox = a.gbox + ((size_t)(t > 0 ? t - 1 : 0) * 2 + 1) * MB_WORDS;
  else pv_rbox = reinterpret_cast<unsigned long long*>(cluster.map_shared_rank(sm + (size_t)(a.W - 1) * POS_DOUBLES + BOXOFF, crank - 1)) + MB_WORDS;
  const int my_sweeps = min(n - 2, n - 1 - CB * t);
  const int nx_sweeps = t + 1 < a.NP ? min(n - 2, n - 1 - CB * (t + 1)) : 0;

  // ---- initial window (sweep 0): rows p .. p+31, p = 1 + 32 t; zero beyond the matrix ----
  {
    const int p = 1 + CB * t;
    const int i = p + lane;   // my row
#pragma unroll 4
    for (int k = 0; k < CB; ++k) {
      double ev = 0.0, dv = 0.0;
      if (i < n) {
        const int je = p - CB + k;           // E column
        if (t >= 1) {
          if (i - je <= CB) ev = a.AB[(i - je) + (long long)je * a.ldab];
        } else if (k == CB - 1) {
          ev = a.AB[(i - 0) + 0];            // position 0: column 0, rows 1 .. 32
        }
        const int jd = p + k;                // D column
        if (jd < n) dv = (i >= jd) ? a.AB[(i - jd) + (long long)jd * a.ldab] : a.AB[(jd - i) + (long long)i * a.ldab];
      }
      Ew[lane * WLD + k] = ev;
      Dw[lane * WLD + k] = dv;
    }
    if (t == 0 && lane == 0) a.d[0] = a.AB[0];
    __syncwarp();
  }

  double Dlast0 = 0.0, Dlast1 = 0.0;   // D[1][0], D[1][1] of position 0 after its last hop
  long long pc[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = a.prof != nullptr;
  const long long tstart = prof ? clock64() : 0;
  for (int s = 0; s < my_sweeps; ++s) {
    const unsigned seq = (unsigned)s + 1u;
    long long tk = prof ? clock64() : 0;
#define CH_TICK(slot)                     \
    if (prof) {                           \
      const long long _n = clock64();     \
      pc[slot] += _n - tk;                \
      tk = _n;                            \
    }
    // ---- entering row (from position t+1, produced at sweep s-1) ----
    const bool have_row = s > 0;
    if (have_row) {
      if (s - 1 < nx_sweeps) {
        if (!mb_recv(my_rbox, (unsigned)s, srow, lane, a.err)) break;
      } else {
        srow[lane] = 0.0;
        if (lane == 0) srow[CB] = 0.0;
        __syncwarp();
      }
    }
    CH_TICK(0)
    // ---- E phase: everything the neighbours wait for comes first ----
    double v, tau, beta;
    double Er[CB];
    if (t == 0) {
      double x = Ew[lane * WLD + (CB - 1)];
      if (have_row && lane == CB - 1) x = srow[0];
      v = warp_house(x, lane, tau, beta);
      if (lane == 0) a.e[s] = beta;
    } else {
#pragma unroll
      for (int k = 0; k < CB; ++k) Er[k] = Ew[lane * WLD + k];
      if (have_row && lane == CB - 1) {
#pragma unroll
        for (int k = 0; k < CB - 1; ++k) Er[k] = 0.0;
        Er[CB - 1] = srow[0];
      }
      // (a) right-apply the previous reflector of this sweep
      CH_TICK(2)
      if (!mb_recv(my_vbox, seq, svp, lane, a.err)) break;
      CH_TICK(1)
      {
        const double taup = svp[CB];
        double dot0 = 0.0, dot1 = 0.0, dot2 = 0.0, dot3 = 0.0;
#pragma unroll
        for (int k = 0; k < CB; k += 4) {
          dot0 = fma(Er[k], svp[k], dot0);
          dot1 = fma(Er[k + 1], svp[k + 1], dot1);
          dot2 = fma(Er[k + 2], svp[k + 2], dot2);
          dot3 = fma(Er[k + 3], svp[k + 3], dot3);
        }
        const double f = taup * ((dot0 + dot1) + (dot2 + dot3));
#pragma unroll
        for (int k = 0; k < CB; ++k) Er[k] = fma(-f, svp[k], Er[k]);
      }
      // (b) reflector that annihilates the first column of the bulge
      v = warp_house(Er[0], lane, tau, beta);
    }
    // reflector out at once: position t+1 is waiting for it
    sv[lane] = v;
    if (s < nx_sweeps) {
      mb_send(nx_vbox, seq, lane, v);
      if (lane == 0) mb_send(nx_vbox, seq, CB, tau);
    }
    if (t > 0) {
      // (c) column sums c_k = sum_rows v_r E[r][k] by recursive halving: lane k ends up with c_k
      double val[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) val[k] = v * Er[k];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
          const double snd = up ? val[k] : val[k + o];
          const double keep = up ? val[k + o] : val[k];
          val[k] = keep + __shfl_xor_sync(0xffffffffu, snd, o);
        }
      }
      sw[lane] = lane == 0 ? 0.0 : tau * val[0];
    }
    __syncwarp();
    CH_TICK(2)
    // ---- D phase, first half: w = tau D v - (tau^2 / 2)(v' D v) v ----
    double Dc[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) Dc[i] = Dw[lane * WLD + i];
    if (have_row) {
      if (lane == CB - 1) {
#pragma unroll
        for (int i = 0; i < CB; ++i) Dc[i] = srow[1 + i];
      } else {
        Dc[CB - 1] = srow[1 + lane];
      }
    }
    double wd;
    {
      double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
      for (int i = 0; i < CB; i += 4) {
        y0 = fma(Dc[i], sv[i], y0);
        y1 = fma(Dc[i + 1], sv[i + 1], y1);
        y2 = fma(Dc[i + 2], sv[i + 2], y2);
        y3 = fma(Dc[i + 3], sv[i + 3], y3);
      }
      wd = tau * ((y0 + y1) + (y2 + y3));
      const double gamma = wsum(wd * v);
      wd = fma(-0.5 * tau * gamma, v, wd);
      sd[lane] = wd;
    }
    // the row that leaves to position t-1 (v_0 = 1): E[0][k] - w_k, D[0][0] - 2 wd_0
    if (lane == 0) {
      if (t > 0) {
        stop[0] = beta;
#pragma unroll
        for (int k = 1; k < CB; ++k) stop[k] = Er[k] - sw[k];
      }
      const double d00 = Dc[0] - 2.0 * wd;
      stop[CB] = d00;
      if (t == 0) a.d[s + 1] = d00;
    }
    __syncwarp();
    CH_TICK(3)
    if (t > 0) {
      mb_send(pv_rbox, seq, lane, stop[lane]);
      if (lane == 0) mb_send(pv_rbox, seq, CB, stop[CB]);
    }
    CH_TICK(4)
    // ---- off the critical path: reflector store, bulk updates, windows of the next sweep (shifted by (1,1)) ----
    a.V2[(long long)s * a.ldv + CB * t + lane] = v;
    if (lane == 0) a.tau2[(long long)s * a.NP + t] = tau;
    if (t > 0 && lane >= 1) {
      double* dst = Ew + (lane - 1) * WLD - 1;
#pragma unroll
      for (int k = 1; k < CB; ++k) dst[k] = fma(-v, sw[k], Er[k]);     // column 0 is annihilated: not stored
    }
#pragma unroll
    for (int i = 0; i < CB; ++i) Dc[i] -= sv[i] * wd + sd[i] * v;
    if (lane >= 1) {
      double* dd = Dw + (lane - 1) * WLD - 1;
#pragma unroll
      for (int i = 1; i < CB; ++i) dd[i] = Dc[i];
      Ew[(lane - 1) * WLD + (CB - 1)] = Dc[0];      // the new last column of E is the old first column of D
    }
    if (t == 0 && s == my_sweeps - 1 && lane == 1) {
      Dlast0 = Dc[0];
      Dlast1 = Dc[1];
    }
    __syncwarp();
    CH_TICK(3)
  }
  if (prof && lane == 0) {
    pc[5] = clock64() - tstart;
    for (int i = 0; i < 6; ++i) a.prof[(long long)t * 6 + i] = pc[i];
  }
  if (t == 0 && lane == 1 && n >= 3) {
    a.d[n - 1] = Dlast1;
    a.e[n - 2] = Dlast0;
  }
  }   // t < NP
  cluster.sync();   // no CTA may exit while a neighbour can still push into its mailboxes
}

// =====================================================================================================================
// back-transformation with Q2:  X <- Q2 X, register-resident systolic pass
// =====================================================================================================================
// One launch = one stage of QS slots (slot q holds rows s+1+32q .. s+32+32q at sweep s).  Thread = (slot, 2 columns):
// 64 matrix elements live in registers for the whole pass; per sweep the window slides up by one row (circular register
// buffer, the loop over sweeps is unrolled 32 times so that every index is static).  The row that enters at the top comes
// from slot q-1 of the same CTA through shared memory, for the first slot of the stage from the stream the previous
// stage wrote (for stage 0: from X itself); the bottom row leaves to slot q+1 / to the stream of the next stage.
constexpr int QS = 32;          // slots per stage (CTA)
constexpr int QC = 16;          // columns per CTA (one per thread, 16 threads per slot)
constexpr int Q2_NT = QS * QC;  // 512 threads

struct Q2Args {
  double* X;                    // n_rows x ncols, leading dimension ldx (rows >= n are not touched)
  long long ldx;
  int n, ncols;
  const double* V2;
  long long ldv;
  const double* tau2;
  int NP;
  int q0;                       // first slot of this stage
  const double* sin;            // stream from the previous stage ((s_top + 1) x lds), null for stage 0
  double* sout;                 // stream to the next stage, null for the last stage
  long long lds;
  int s_top;                    // first (largest) sweep index processed; s_top + 1 is a multiple of 32
};

__global__ void __launch_bounds__(Q2_NT, 1) k_q2_stage(const Q2Args a) {
  constexpr int NB = 4;                           // ring of reflector buffers: prefetch distance 3 sweeps
  constexpr int UNR = 4;                          // sweeps per physical register shift
  __shared__ double vbuf[NB][QS * CB];            // the reflectors of the stage's slots for one sweep (8 KB each)
  __shared__ double tbuf[NB][QS];                 // their taus
  __shared__ double xfer[2][QS][QC];              // bottom rows handed to the next slot
  const int tid = threadIdx.x;
  const int sl = tid >> 4;                        // slot within the stage (a warp holds two slots)
  const int cl = tid & 15;                        // column within the CTA
  const int q = a.q0 + sl;
  const int c0 = blockIdx.x * QC + cl;
  const bool ok0 = c0 < a.ncols;
  const int n = a.n;
  // window registers: at sub-step r of a block of UNR sweeps the logical row k lives in x[k + UNR - 1 - r]; every UNR
  // sweeps the registers are shifted up by UNR (a fully static circular buffer would need the loop unrolled 32 times,
  // which does not fit the instruction cache)
  double x0[CB + UNR];
#pragma unroll
  for (int k = 0; k < CB + UNR; ++k) x0[k] = 0.0;
  for (int i = tid; i < 2 * QS * QC; i += Q2_NT) (&xfer[0][0][0])[i] = 0.0;

  auto stage_v = [&](int s, int buf) {   // cp.async the 32 x 32 reflector block of sweep s (nothing for sweeps without reflectors)
    double* dst = vbuf[buf];
    if (s >= 0 && s <= n - 3) {
      const double* src = a.V2 + (long long)s * a.ldv + (long long)CB * a.q0;
      const int avail = (int)min((long long)QS * CB, a.ldv - (long long)CB * a.q0);   // doubles available in this row
      const int i = tid * 2;                       // 512 threads x 2 doubles = the whole block
      if (i + 1 < avail) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
      } else {
        dst[i] = i < avail ? src[i] : 0.0;
        dst[i + 1] = 0.0;
      }
    }
    if (tid < QS) {
      const int qq = a.q0 + tid;
      if (s >= 0 && s <= n - 3 && qq < a.NP) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(&tbuf[buf][tid]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(a.tau2 + (long long)s * a.NP + qq));
      } else {
        tbuf[buf][tid] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  // entering rows of the first slot of the stage come from global memory: fetched three steps ahead
  auto fetch_top = [&](int s) -> double {
    if (sl != 0 || s < 0 || !ok0) return 0.0;
    if (a.sin) return s <= n - 2 ? a.sin[(long long)s * a.lds + c0] : 0.0;
    return s + 1 < n ? a.X[(s + 1) + (long long)c0 * a.ldx] : 0.0;
  };

  double pre0[NB];
#pragma unroll
  for (int i = 0; i < NB - 1; ++i) {
    stage_v(a.s_top - i, i);
    pre0[i] = fetch_top(a.s_top - i);
  }
  pre0[NB - 1] = 0.0;
  int par = 0;
  for (int sb = a.s_top; sb >= 0; sb -= UNR) {
#pragma unroll
    for (int r = 0; r < UNR; ++r) {
      const int s = sb - r;
      const int buf = r;                             // (s_top - s) mod NB == r because s_top + 1 and UNR are multiples of NB
      asm volatile("cp.async.wait_group %0;\n" ::"n"(NB - 2));
      __syncthreads();                               // vbuf[buf], tbuf[buf] and xfer[par] (written in the previous step) are visible;
      stage_v(s - (NB - 1), (r + NB - 1) % NB);      // every warp has left the previous step: its buffer may be refilled
      // entering row p = s + 1 + 32 q
      double e0;
      if (sl == 0) {
        e0 = pre0[r];
        pre0[(r + NB - 1) % NB] = fetch_top(s - (NB - 1));
      } else {
        e0 = xfer[par][sl - 1][cl];
      }
      const int off = UNR - 1 - r;                   // logical row k <-> register k + off
      x0[off] = e0;
      const double tau = tbuf[buf][sl];
      if (tau != 0.0) {
        const double* vv = vbuf[buf] + sl * CB;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
        for (int k = 0; k < CB; k += 4) {
          const double2 va = *reinterpret_cast<const double2*>(vv + k);
          const double2 vb = *reinterpret_cast<const double2*>(vv + k + 2);
          d0 = fma(va.x, x0[k + off], d0);
          d1 = fma(va.y, x0[k + 1 + off], d1);
          d2 = fma(vb.x, x0[k + 2 + off], d2);
          d3 = fma(vb.y, x0[k + 3 + off], d3);
        }
        const double f0 = -tau * ((d0 + d1) + (d2 + d3));
#pragma unroll
        for (int k = 0; k < CB; k += 2) {
          const double2 v2 = *reinterpret_cast<const double2*>(vv + k);
          x0[k + off] = fma(f0, v2.x, x0[k + off]);
          x0[k + 1 + off] = fma(f0, v2.y, x0[k + 1 + off]);
        }
      }
      // the bottom row (logical 31) leaves at the next slide
      const double b0 = x0[CB - 1 + off];
      xfer[par ^ 1][sl][cl] = b0;
      if (sl == QS - 1 && a.sout && s >= 1 && ok0) a.sout[(long long)(s - 1) * a.lds + c0] = b0;
      par ^= 1;
    }
    // after UNR sweeps logical row k sits in register k: move everything up by UNR for the next block
#pragma unroll
    for (int k = CB - 1; k >= 0; --k) x0[k + UNR] = x0[k];
  }
  // final windows (state after sweep 0, already shifted: logical row k in register k + UNR), rows 1 + 32 q ..
  const int p = 1 + CB * q;
  if (ok0) {
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (p + k < n) a.X[(p + k) + (long long)c0 * a.ldx] = x0[k + UNR];
  }
}

// =====================================================================================================================
// stage 1: dense -> band.  Householder QR of one 32-column sub-band panel inside a thread-block cluster
// =====================================================================================================================
// The m x 32 panel (rows r0 = j + 32 .. n-1 of columns j .. j+31) is split by rows over the CTAs of the cluster and stays
// in (distributed) shared memory; inside a CTA every row belongs to one lane of one warp for the whole factorisation.
// Per column: every lane forms the products of its rows' column-c entry with all 32 columns, a recursive-halving warp
// reduction leaves the warp's partial dot for column k in lane k, one block barrier, warp 0 pushes the CTA's partial sums
// (and, from the owner, the pivot row) into the exchange buffers of all CTAs, ONE hardware cluster barrier, and every
// warp derives beta, tau and the update coefficients redundantly and updates its own rows.  The dots with the finished
// reflectors (k < c) are what the compact-WY factor T needs; T is assembled once, after the last column.
constexpr int PQ_NT = 256;
constexpr int PQ_NW = PQ_NT / 32;
constexpr int PQ_MAXCS = 16;
struct PanelArgs {
  double* A;
  long long lda;
  int n, j, rp, ldp;  // rp: rows per CTA (multiple of 256), ldp = rp + 1: leading dimension of the shared-memory panel
  double* Y;          // reflector store (n x n, ldy): unit lower trapezoidal panel written at rows r0.., columns j..j+31
  long long ldy;
  double* tau;        // tau[j + k]
  double* T;          // 32 x 32 compact-WY factor of this panel (column-major)
  long long* prof;    // optional: 8 cycle counters (rank 0, thread 0)
};

__global__ void __launch_bounds__(PQ_NT, 1) k_panel_qr(const PanelArgs a) {
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool prof = a.prof != nullptr;
  long long tk = prof ? clock64() : 0;
#define PQ_TICK(slot)                 \
  if (prof) {                         \
    const long long _n = clock64();   \
    pc[slot] += _n - tk;              \
    tk = _n;                          \
  }
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  extern __shared__ double sm[];
  const int rp = a.rp, ldp = a.ldp;
  double* P = sm;                              // 32 columns x rp rows, column-major (ld = ldp, odd)
  double* ex = P + (size_t)CB * ldp;           // [2][PQ_MAXCS][32] partial dots
  double* piv = ex + 2 * PQ_MAXCS * CB;        // [2][32] pivot row
  double* wpart = piv + 2 * CB;                // [2][PQ_NW][32] per-warp partial dots
  double* wsw = wpart + 2 * PQ_NW * CB;        // [PQ_NW][32] update coefficients, one copy per warp
  double* Zm = wsw + PQ_NW * CB;               // [32*32] Zm[k + c*32] = y_k' y_c (k < c)
  double* sT = Zm + CB * CB;                   // [32*32] T factor
  double* stau = sT + CB * CB;                 // [32]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r0 = a.j + CB, m = a.n - r0;
  const int kb = min(CB, m - 1);
  const int row_lo = rank * rp;
  const int nr = max(0, min(m, row_lo + rp) - row_lo);
  const int nrl = rp / PQ_NT;                  // rows per lane
  for (int rb = tid; rb < rp; rb += PQ_NT) {
    const double* src = a.A + (r0 + row_lo + rb) + (long long)a.j * a.lda;
#pragma unroll
    for (int k0 = 0; k0 < CB; k0 += 8) {           // eight loads in flight per thread
      double tmp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) tmp[u] = rb < nr ? src[(long long)(k0 + u) * a.lda] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) P[(size_t)(k0 + u) * ldp + rb] = tmp[u];
    }
  }
  for (int idx = tid; idx < CB * CB; idx += PQ_NT) {
    sT[idx] = 0.0;
    Zm[idx] = 0.0;
  }
  if (tid < CB) stau[tid] = 0.0;
  __syncthreads();
  cluster.sync();
  PQ_TICK(0)
  for (int c = 0; c < kb; ++c) {
    const int par = c & 1;
    // (A) products of my rows (strictly below the pivot row) with every column, reduced over the warp
    {
      double val[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) val[k] = 0.0;
      for (int i = 0; i < nrl; ++i) {
        const int r = i * PQ_NT + tid;
        if (row_lo + r > c && r < nr) {
          const double pc = P[(size_t)c * ldp + r];
#pragma unroll
          for (int k = 0; k < CB; ++k) val[k] = fma(pc, P[(size_t)k * ldp + r], val[k]);
        }
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
          const double snd = up ? val[k] : val[k + o];
          const double keep = up ? val[k + o] : val[k];
          val[k] = keep + __shfl_xor_sync(0xffffffffu, snd, o);
        }
      }
      wpart[((size_t)par * PQ_NW + wid) * CB + lane] = val[0];
    }
    PQ_TICK(1)
    __syncthreads();
    PQ_TICK(2)
    // (B) CTA partial -> exchange buffers of every CTA; the owner of the pivot row adds that row
    if (wid == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < PQ_NW; ++w) tot += wpart[((size_t)par * PQ_NW + w) * CB + lane];
      const int prank = c / rp;
      const double pv = rank == prank ? P[(size_t)lane * ldp + (c - row_lo)] : 0.0;
      for (int dsti = 0; dsti < CS; ++dsti) {
        cluster.map_shared_rank(ex, dsti)[((size_t)par * PQ_MAXCS + rank) * CB + lane] = tot;
        if (rank == prank) cluster.map_shared_rank(piv, dsti)[par * CB + lane] = pv;
      }
    }
    PQ_TICK(3)
    cluster.sync();
    PQ_TICK(4)
    // (C) every warp: scalars of the reflector and the update coefficients
    double scale, beta;
    {
      double g = 0.0;
      for (int src = 0; src < CS; ++src) g += ex[((size_t)par * PQ_MAXCS + src) * CB + lane];
      const double pk = piv[par * CB + lane];
      const double xn2 = __shfl_sync(0xffffffffu, g, c);
      const double alpha = __shfl_sync(0xffffffffu, pk, c);
      double tau = 0.0;
      beta = alpha;
      scale = 0.0;
      if (xn2 > 0.0) {                            // beta = -sign(alpha) |x|, tau = 1 + |alpha| / |x|, scale = sign(alpha) / (|alpha| + |x|)
        const double s2 = fma(alpha, alpha, xn2);
        const double rn = rsqrt(s2);
        const double nrm = s2 * rn, aa = fabs(alpha);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = fma(aa, rn, 1.0);
        const double rc = 1.0 / (aa + nrm);
        scale = alpha >= 0.0 ? rc : -rc;
      }
      const double z = pk + scale * g;            // y_k' y_c for k < c;  v_c' P[:, k] for k > c
      wsw[wid * CB + lane] = lane > c ? tau * z : 0.0;
      if (wid == 0) {
        if (lane < c) Zm[lane + c * CB] = z;
        if (lane == c) stau[c] = tau;
      }
      __syncwarp();
    }
    PQ_TICK(5)
    // (D) apply to my rows
    {
      const double* w = wsw + wid * CB;
      for (int i = 0; i < nrl; ++i) {
        const int r = i * PQ_NT + tid;
        if (r >= nr) continue;
        const int grow = row_lo + r;
        if (grow >= c) {                           // w[k] = 0 for k <= c: the loop is unrolled with static addresses
          const double vr = grow > c ? scale * P[(size_t)c * ldp + r] : 1.0;
          double pk[CB];
#pragma unroll
          for (int k = 0; k < CB; ++k) pk[k] = P[(size_t)k * ldp + r];
#pragma unroll
          for (int k = 0; k < CB; ++k) pk[k] = fma(-vr, w[k], pk[k]);
#pragma unroll
          for (int k = 0; k < CB; ++k) P[(size_t)k * ldp + r] = pk[k];
          P[(size_t)c * ldp + r] = grow > c ? vr : beta;
        }
      }
      __syncwarp();
    }
    PQ_TICK(6)
  }
  __syncthreads();
  // T (dlarft, forward / columnwise): T[0:c, c] = -tau_c T[0:c, 0:c] z[0:c, c], T[c, c] = tau_c
  if (rank == 0 && wid == 0) {
    // lane i owns row i of T (registers): T[i][c] = -tau_c sum_{k=i}^{c-1} T[i][k] Zm[k][c] needs nothing from other lanes
    double trow[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) trow[c] = 0.0;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double tau = stau[c];
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < c; ++k) acc = fma(trow[k], Zm[k + c * CB], acc);   // trow[k] = 0 for k < lane
      trow[c] = lane < c ? -tau * acc : (lane == c ? tau : 0.0);
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) sT[lane + c * CB] = trow[c];
  }
  __syncthreads();
  // outputs: R into the band part of A, Y (unit lower trapezoidal, explicit zeros) into the reflector store
  for (int k = 0; k < CB; ++k)
    for (int r = tid; r < nr; r += PQ_NT) {
      const int grow = row_lo + r;
      const double v = P[(size_t)k * ldp + r];
      double y;
      if (k >= kb) y = 0.0;
      else y = grow > k ? v : (grow == k ? 1.0 : 0.0);
      a.Y[(r0 + grow) + (long long)(a.j + k) * a.ldy] = y;
      a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = (k >= kb || grow <= k) ? v : 0.0;
    }
  if (rank == 0) {
    for (int idx = tid; idx < CB * CB; idx += PQ_NT) a.T[idx] = sT[idx];
    if (tid < CB) a.tau[a.j + tid] = stau[tid];
  }
  cluster.sync();   // no CTA may exit while a peer can still address its shared memory
  PQ_TICK(7)
  if (prof && rank == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + i, (unsigned long long)pc[i]);
}

// Z0 = A22 Y (A22 symmetric m x m, full storage; Y m x 32): CTA (rb, ks) accumulates the 64 x 32 tile of rows
// 64 rb .. over the k range of split ks and writes it to Zpart[ks]; it also forms its share of G0 = Y' Z0 = Y' A22 Y.
// k_reduce_g sums the shares in a fixed order (deterministic).
constexpr int SY_BM = 64, SY_BK = 32;
__global__ void __launch_bounds__(256) k_symm_y(const double* __restrict__ A, long long lda, const double* __restrict__ Y, long long ldy, int m,
                                                int ksplit, double* __restrict__ Zpart, long long ldz, double* __restrict__ Gpart) {
  __shared__ double raw[SY_BK * (SY_BM + 1) + SY_BK * (CB + 1)];
  __shared__ double Yr[SY_BM][CB + 1];
  double (*As)[SY_BM + 1] = reinterpret_cast<double (*)[SY_BM + 1]>(raw);                       // As[kk][row]
  double (*Ys)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(raw + SY_BK * (SY_BM + 1));       // Ys[kk][col]
  double (*Zs)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(raw);                             // after the k loop: 64 x 33
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int rb = blockIdx.x, ks = blockIdx.y;
  const int row0 = rb * SY_BM;
  const int kchunk = ((m + ksplit - 1) / ksplit + SY_BK - 1) / SY_BK * SY_BK;
  const int k0 = ks * kchunk, k1 = min(m, k0 + kchunk);
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  for (int kk0 = k0; kk0 < k1; kk0 += SY_BK) {
    // A tile: rows row0 .. +63, columns kk0 .. +31 (column-major source: rows contiguous)
    for (int idx = tid; idx < SY_BK * SY_BM; idx += 256) {
      const int r = idx & (SY_BM - 1), kk = idx >> 6;
      const int gr = row0 + r, gk = kk0 + kk;
      As[kk][r] = (gr < m && gk < k1) ? A[gr + (long long)gk * lda] : 0.0;
    }
    for (int idx = tid; idx < SY_BK * CB; idx += 256) {
      const int kk = idx & 31, col = idx >> 5;
      const int gk = kk0 + kk;
      Ys[kk][col] = gk < k1 ? Y[gk + (long long)col * ldy] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < SY_BK; ++kk) {
      const double yv = Ys[kk][tx];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(As[kk][ty * 8 + i], yv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gr = row0 + ty * 8 + i;
    Zs[ty * 8 + i][tx] = acc[i];
    if (gr < m) Zpart[(long long)ks * ldz * CB + gr + (long long)tx * ldz] = acc[i];
  }
  for (int idx = tid; idx < SY_BM * CB; idx += 256) {
    const int r = idx & (SY_BM - 1), col = idx >> 6;
    Yr[r][col] = (row0 + r < m) ? Y[(row0 + r) + (long long)col * ldy] : 0.0;
  }
  __syncthreads();
  // G share: Gp[i][j] = sum_r Yr[r][i] Zs[r][j]; thread computes i = ty*4 .. +3, j = tx
  double* gp = Gpart + ((long long)ks * gridDim.x + rb) * (CB * CB);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ty * 4 + q;
    double g = 0.0;
#pragma unroll 8
    for (int r = 0; r < SY_BM; ++r) g = fma(Yr[r][i], Zs[r][tx], g);
    gp[i + tx * CB] = g;
  }
}

// G0[idx] = sum over the shares (fixed order: 8 interleaved partial sums per entry, combined by a shuffle tree)
__global__ void __launch_bounds__(256) k_reduce_g(const double* __restrict__ Gpart, int nshare, double* __restrict__ G0) {
  const int gid = blockIdx.x * 256 + threadIdx.x;
  const int idx = gid >> 3, sub = gid & 7;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int sidx = sub;
  for (; sidx + 24 < nshare; sidx += 32) {
    a0 += Gpart[(long long)sidx * (CB * CB) + idx];
    a1 += Gpart[(long long)(sidx + 8) * (CB * CB) + idx];
    a2 += Gpart[(long long)(sidx + 16) * (CB * CB) + idx];
    a3 += Gpart[(long long)(sidx + 24) * (CB * CB) + idx];
  }
  for (; sidx < nshare; sidx += 8) a0 += Gpart[(long long)sidx * (CB * CB) + idx];
  double g = (a0 + a1) + (a2 + a3);
  g += __shfl_xor_sync(0xffffffffu, g, 1);
  g += __shfl_xor_sync(0xffffffffu, g, 2);
  g += __shfl_xor_sync(0xffffffffu, g, 4);
  if (sub == 0) G0[idx] = g;
}

// W = Z0 T - 1/2 Y M with Z0 = sum of the k-split partials, M = T' G0 T;  P1 = [Y W], P2 = [W Y] (leading dimension ldp)
__global__ void __launch_bounds__(256) k_make_w(const double* __restrict__ Zpart, long long ldz, int ksplit, const double* __restrict__ Y,
                                                long long ldy, const double* __restrict__ T, const double* __restrict__ G0, int m,
                                                double* __restrict__ P1, double* __restrict__ P2, long long ldp) {
  __shared__ double sT[CB][CB + 1], sM[CB][CB + 1];
  __shared__ double zr[32][CB + 1], yr[32][CB + 1];
  const int tid = threadIdx.x;
  for (int i = tid; i < CB * CB; i += 256) {
    sT[i & 31][i >> 5] = T[i];
    zr[i & 31][i >> 5] = G0[i];
  }
  __syncthreads();
  for (int idx = tid; idx < CB * CB; idx += 256) {      // X = G0 T (T upper triangular)
    const int i = idx & 31, k = idx >> 5;
    double acc = 0.0;
    for (int q = 0; q <= k; ++q) acc = fma(zr[i][q], sT[q][k], acc);
    yr[i][k] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < CB * CB; idx += 256) {      // M = T' X
    const int i = idx & 31, k = idx >> 5;
    double acc = 0.0;
    for (int q = 0; q <= i; ++q) acc = fma(sT[q][i], yr[q][k], acc);
    sM[i][k] = acc;
  }
  __syncthreads();
  const int row0 = blockIdx.x * 32;
  for (int idx = tid; idx < 32 * CB; idx += 256) {
    const int r = idx & 31, col = idx >> 5;
    const int gr = row0 + r;
    double z = 0.0, y = 0.0;
    if (gr < m) {
      for (int ks = 0; ks < ksplit; ++ks) z += Zpart[(long long)ks * ldz * CB + gr + (long long)col * ldz];
      y = Y[gr + (long long)col * ldy];
    }
    zr[r][col] = z;
    yr[r][col] = y;
  }
  __syncthreads();
  // thread: row r = tid & 31, columns k = (tid >> 5) * 4 .. +3
  const int r = tid & 31, kq = (tid >> 5) * 4;
  const int gr = row0 + r;
  if (gr >= m) return;
  double acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0.0;
#pragma unroll 4
  for (int q = 0; q < CB; ++q) {
    const double zv = zr[r][q], yv = -0.5 * yr[r][q];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fma(zv, sT[q][kq + i], fma(yv, sM[q][kq + i], acc[i]));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kq + i;
    const double w = acc[i], y = yr[r][k];
    P1[gr + (long long)k * ldp] = y;
    P1[gr + (long long)(CB + k) * ldp] = w;
    P2[gr + (long long)k * ldp] = w;
    P2[gr + (long long)(CB + k) * ldp] = y;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
int64_t chase_positions(int64_t n) { return n >= 2 ? (n - 2) / CB + 1 : 1; }

// Band (lower storage AB, ldab >= 33, half bandwidth 32) -> tridiagonal (dd, ee); reflectors to V2 ((n-2) x ldv with
// ldv = 32 * chase_positions(n), zero-initialised by the caller) and tau2 ((n-2) x NP, zero-initialised).
void sb2st(tnad_ctx* c, const double* AB, int64_t ldab, int64_t n, double* dd, double* ee, double* V2, int64_t ldv, double* tau2) {
  TNAD_REQUIRE(n >= 3 && ldab >= CB + 1, "sb2st: need n >= 3 and a band store with 33 rows");
  TNAD_REQUIRE(c->coop_launch, "sb2st: the chase kernel needs cooperative (co-resident) launches");
  const int NP = (int)chase_positions(n);
  int W = opt_i(c, "TNAD_CHASE_W", 4);
  while ((NP + W - 1) / W > c->num_sms) ++W;      // all CTAs must be co-resident (one per SM)
  TNAD_REQUIRE(W <= 8, "sb2st: matrix too large for the chase kernel (n <= 32 * 8 * #SMs)");
  const int G = (NP + W - 1) / W;
  Tens gbox = t_alloc(c, {(int64_t)NP * 2 * MB_WORDS}, true);
  Tens err = t_alloc(c, {2}, true);
  ChaseArgs a;
  a.AB = AB; a.ldab = (int)ldab; a.n = (int)n; a.NP = NP; a.W = W;
  a.d = dd; a.e = ee; a.V2 = V2; a.ldv = ldv; a.tau2 = tau2;
  a.gbox = reinterpret_cast<unsigned long long*>(gbox.p);
  a.err = reinterpret_cast<int*>(err.p);
  const bool prof = opt_i(c, "TNAD_DC_DEBUG", 0) >= 2;
  Tens pbuf = t_alloc(c, {(int64_t)NP * 6 + 2}, true);
  a.prof = prof ? reinterpret_cast<long long*>(pbuf.p) : nullptr;
  const size_t smem = (size_t)W * POS_DOUBLES * sizeof(double);
  TNAD_CUDA(cudaFuncSetAttribute(k_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TNAD_CUDA(cudaFuncSetAttribute(k_chase, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // CTAs of one cluster exchange their messages through distributed shared memory (about 300 cycles one way instead of
  // about 1500 through L2: the pipeline period is set by its slowest link); clusters talk to each other through L2
  int CSZ = std::min(opt_i(c, "TNAD_CHASE_CLUSTER", 16), G);
  CSZ = std::max(1, std::min(CSZ, 16));
  const int Gp = (G + CSZ - 1) / CSZ * CSZ;
  TNAD_REQUIRE(Gp <= c->num_sms, "sb2st: matrix too large for the chase kernel");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(Gp);
  cfg.blockDim = dim3(32 * W);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CSZ;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  {
    KTimer kt(c, KF_EIG);
    cudaError_t le = cudaLaunchKernelEx(&cfg, k_chase, a);
    if (le != cudaSuccess) {   // co-residency cannot be promised together with this cluster shape: the mailbox time-out still guards
      cudaGetLastError();
      cfg.numAttrs = 1;
      le = cudaLaunchKernelEx(&cfg, k_chase, a);
    }
    TNAD_CUDA(le);
  }
  c->launches++;
  int herr = 0;
  TNAD_CUDA(cudaMemcpyAsync(&herr, a.err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  if (herr) fail(TNAD_ERR_INTERNAL, "sb2st: the chase pipeline timed out");
  if (prof) {
    std::vector<long long> ph((size_t)NP * 6);
    TNAD_CUDA(cudaMemcpy(ph.data(), pbuf.p, ph.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int t : {0, 1, 2, 3, 4, 5, NP / 2, NP / 2 + 1}) {
      if (t >= NP) continue;
      const double ns = (double)std::max<int64_t>(1, std::min<int64_t>(n - 2, n - 1 - CB * t));
      fprintf(stderr, "[tnad dc] chase position %d (W=%d) cycles/hop: wait row %.0f  wait v %.0f  E phase %.0f  D phase %.0f  sends %.0f  | total/hop %.0f\n",
              t, W, ph[(size_t)t * 6 + 0] / ns, ph[(size_t)t * 6 + 1] / ns, ph[(size_t)t * 6 + 2] / ns, ph[(size_t)t * 6 + 3] / ns,
              ph[(size_t)t * 6 + 4] / ns, ph[(size_t)t * 6 + 5] / ns);
    }
  }
}

// X[0:n, 0:ncols] <- Q2 X (reflectors of sb2st)
void apply_q2(tnad_ctx* c, const double* V2, int64_t ldv, const double* tau2, int64_t n, double* X, int64_t ldx, int64_t ncols) {
  if (n < 3 || ncols <= 0) return;
  const int NP = (int)chase_positions(n);
  const int nst = (NP + QS - 1) / QS;
  const int s_top = (int)(((n - 1) + 31) / 32 * 32 - 1);     // n - 2 rounded up: one leading no-op sweep at least
  const int64_t lds = (ncols + 1) & ~1LL;
  Tens sa, sb;
  if (nst > 1) {
    sa = t_alloc(c, {lds, (int64_t)s_top + 1});
    if (nst > 2) sb = t_alloc(c, {lds, (int64_t)s_top + 1});
  }
  const int grid = (int)((ncols + QC - 1) / QC);
  for (int k = 0; k < nst; ++k) {
    Q2Args a;
    a.X = X; a.ldx = ldx; a.n = (int)n; a.ncols = (int)ncols;
    a.V2 = V2; a.ldv = ldv; a.tau2 = tau2; a.NP = NP; a.q0 = k * QS;
    a.sin = k == 0 ? nullptr : ((k & 1) ? sa.p : sb.p);
    a.sout = k == nst - 1 ? nullptr : ((k & 1) ? sb.p : sa.p);
    a.lds = lds;
    // the first slot of the stage is active for sweeps s <= n - 2 - 32 q0 only: later (= earlier in time) sweeps are skipped
    a.s_top = (int)std::min<int64_t>(s_top, ((n - 1 - (int64_t)CB * a.q0) + 31) / 32 * 32 - 1);
    KTimer kt(c, KF_UPDATE);
    k_q2_stage<<<grid, Q2_NT, 0, c->stream>>>(a);
    LAUNCH_CHECK(c);
  }
}

static Tens bview(double* p, int64_t rows, int64_t cols, int64_t ld) {
  Tens t = t_wrap(p, {rows, cols});
  t.str[0] = 1;
  t.str[1] = ld;
  return t;
}

// Dense symmetric A (n x n, full storage, overwritten) -> band: on return the band (|i - j| <= 32) of the lower
// triangle of A holds B; Yst (n x n, ldy, zero-initialised) receives the reflectors of panel j in columns j .. j+31
// (rows >= j + 32, unit element of column c at row c + 32), tau1 (n, zero-initialised) their scalars.
void sy2sb(tnad_ctx* c, double* A, int64_t lda, int64_t n, double* Yst, int64_t ldy, double* tau1) {
  if (n <= CB + 1) return;
  cudaStream_t st = c->stream;
  const int ksplit_max = 8;
  const int64_t nrb_max = (n + SY_BM - 1) / SY_BM;
  Tens Tb = t_alloc(c, {CB, CB}), Mb = t_alloc(c, {CB, CB}), Zp = t_alloc(c, {n, CB, (int64_t)ksplit_max});
  Tens Gp = t_alloc(c, {CB * CB, nrb_max * ksplit_max});
  Tens P1 = t_alloc(c, {n, 2 * CB}), P2 = t_alloc(c, {n, 2 * CB});
  static std::atomic<unsigned long long> attr_devs{0};
  if (!((attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL)) {
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const size_t fixed = (size_t)(2 * PQ_MAXCS * CB + 2 * CB + 2 * PQ_NW * CB + PQ_NW * CB + 2 * CB * CB + CB) * sizeof(double);
  const int rp_max = (int)((((232448 - 1024) - fixed) / (CB * sizeof(double)) - 1) / PQ_NT * PQ_NT);
  const bool prof = opt_i(c, "TNAD_DC_DEBUG", 0) >= 2;
  double tph[4] = {0, 0, 0, 0};
  Tens pprof = t_alloc(c, {8}, true);
  cudaEvent_t pe[5];
  if (prof)
    for (auto& e : pe) e = get_event(c);
  for (int64_t j = 0; j + CB < n - 1; j += CB) {
    const int64_t r0 = j + CB, m = n - r0;
    int CS = (int)std::min<int64_t>(8, std::max<int64_t>(1, (m + PQ_NT - 1) / PQ_NT));
    int rp = (int)(((m + CS - 1) / CS + PQ_NT - 1) / PQ_NT * PQ_NT);
    if (rp > rp_max) {
      CS = PQ_MAXCS;
      rp = (int)(((m + CS - 1) / CS + PQ_NT - 1) / PQ_NT * PQ_NT);
    }
    TNAD_REQUIRE(rp <= rp_max, "sy2sb: panel too tall for one thread-block cluster (n <= 12000)");
    PanelArgs pa;
    pa.A = A; pa.lda = lda; pa.n = (int)n; pa.j = (int)j; pa.rp = rp; pa.ldp = rp + 1;
    pa.Y = Yst; pa.ldy = ldy; pa.tau = tau1; pa.T = Tb.p;
    pa.prof = prof ? reinterpret_cast<long long*>(pprof.p) : nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(PQ_NT);
    cfg.dynamicSmemBytes = fixed + (size_t)CB * (rp + 1) * sizeof(double);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (prof) TNAD_CUDA(cudaEventRecord(pe[0], st));
    {
      KTimer kt(c, KF_EIG);
      TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr, pa));
    }
    c->launches++;
    if (prof) TNAD_CUDA(cudaEventRecord(pe[1], st));
    // two-sided update of the trailing matrix: Z0 = A22 Y, M = T' (Y' Z0) T, W = Z0 T - Y M / 2, A22 -= Y W' + W Y'
    double* A22p = A + r0 + r0 * lda;
    const double* Yp = Yst + r0 + j * ldy;
    const int nrb = (int)((m + SY_BM - 1) / SY_BM);
    int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(ksplit_max, (2 * c->num_sms + nrb - 1) / nrb));
    ksplit = (int)std::min<int64_t>(ksplit, (m + 4 * SY_BK - 1) / (4 * SY_BK));
    {
      KTimer kt(c, KF_GEMM);
      k_symm_y<<<dim3(nrb, ksplit), 256, 0, st>>>(A22p, lda, Yp, ldy, (int)m, ksplit, Zp.p, n, Gp.p);
      LAUNCH_CHECK(c);
      k_reduce_g<<<CB * CB * 8 / 256, 256, 0, st>>>(Gp.p, nrb * ksplit, Mb.p);
    }
    LAUNCH_CHECK(c);
    if (prof) TNAD_CUDA(cudaEventRecord(pe[2], st));
    k_make_w<<<(int)((m + 31) / 32), 256, 0, st>>>(Zp.p, n, ksplit, Yp, ldy, Tb.p, Mb.p, (int)m, P1.p, P2.p, n);
    LAUNCH_CHECK(c);
    if (prof) TNAD_CUDA(cudaEventRecord(pe[3], st));
    Tens A22 = bview(A22p, m, m, lda);
    Tens L = bview(P1.p, m, 2 * CB, n), R = bview(P2.p, m, 2 * CB, n);
    contract(c, "ik,jk->ij", L, R, A22, -1.0, 1.0);
    if (prof) {
      TNAD_CUDA(cudaEventRecord(pe[4], st));
      TNAD_CUDA(cudaEventSynchronize(pe[4]));
      for (int i = 0; i < 4; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pe[i], pe[i + 1]);
        tph[i] += ms;
      }
    }
  }
  if (prof) {
    fprintf(stderr, "[tnad dc] sy2sb n=%lld (synchronised per panel): panel QR %.2f  Z0=A22*Y + G0 %.2f  make_w %.2f  update %.2f ms\n",
            (long long)n, tph[0], tph[1], tph[2], tph[3]);
    for (auto& e : pe) c->event_pool.push_back(e);
    long long ph[8];
    TNAD_CUDA(cudaMemcpy(ph, pprof.p, sizeof(ph), cudaMemcpyDeviceToHost));
    const double ncol = (double)std::max<int64_t>(1, n - CB - 1), npan = std::ceil(ncol / CB);
    fprintf(stderr, "[tnad dc] panel QR (rank 0, thread 0) cycles/column: products+halving %.0f  block barrier %.0f  push %.0f  cluster barrier %.0f  scalars %.0f  update %.0f | per panel: load %.0f  T+store %.0f\n",
            ph[1] / ncol, ph[2] / ncol, ph[3] / ncol, ph[4] / ncol, ph[5] / ncol, ph[6] / ncol, ph[0] / npan, ph[7] / npan);
  }
}

void extract_band(tnad_ctx* c, const double* A, int64_t lda, int64_t n, double* AB, int64_t ldab) {
  const long long total = (long long)n * (CB + 1);
  const int nb = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 8));
  k_extract_band<<<nb, 256, 0, c->stream>>>(A, lda, (int)n, AB, (int)ldab);
  LAUNCH_CHECK(c);
}

}  // namespace tnad
