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
  const double nrm = sqrt(alpha * alpha + xn2);
  beta = alpha >= 0.0 ? -nrm : nrm;
  tau = (beta - alpha) / beta;
  const double scale = 1.0 / (alpha - beta);
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
};

// per-position shared memory (doubles): Ew (1 + 32*33 + 1), Dw (32*33), 5 scratch vectors of 34, two mailboxes
constexpr int POS_DOUBLES = (2 + CB * WLD) + CB * WLD + 5 * 34 + 2 * MB_WORDS;

__global__ void __launch_bounds__(256, 1) k_chase(const ChaseArgs a) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // zero the local mailboxes of every position of this CTA
  for (int i = threadIdx.x; i < a.W * POS_DOUBLES; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const int t = blockIdx.x * a.W + wid;
  if (t >= a.NP) return;
  const int n = a.n;
  double* base = sm + (size_t)wid * POS_DOUBLES;
  double* Ew = base + 1;                       // one double of head room: the shifted store of row 1 touches Ew[-1]
  double* Dw = base + 2 + CB * WLD;
  double* sv = Dw + CB * WLD;                  // own reflector (34)
  double* svp = sv + 34;                       // previous reflector of the sweep + tau (34)
  double* srow = svp + 34;                     // entering row (34)
  double* sw = srow + 34;                      // broadcast scratch w (34)
  double* stop = sw + 34;                      // leaving top row (34)
  unsigned long long* lbox = reinterpret_cast<unsigned long long*>(stop + 34);   // [0] reflector box, [1] row box (local)
  // boxes I read: my own; boxes I write: reflector box of t+1, row box of t-1 -- in shared memory when that position
  // lives in this CTA, in global memory otherwise
  const bool prev_local = t > 0 && wid > 0, next_local = wid + 1 < a.W && t + 1 < a.NP;
  unsigned long long* my_vbox = (t > 0 && !prev_local) ? a.gbox + ((size_t)t * 2 + 0) * MB_WORDS : lbox;
  unsigned long long* my_rbox = (t + 1 < a.NP && !next_local) ? a.gbox + ((size_t)t * 2 + 1) * MB_WORDS : lbox + MB_WORDS;
  unsigned long long* nx_vbox = next_local ? reinterpret_cast<unsigned long long*>(base + POS_DOUBLES + (POS_DOUBLES - 2 * MB_WORDS))
                                           : a.gbox + ((size_t)(t + 1) * 2 + 0) * MB_WORDS;
  unsigned long long* pv_rbox = prev_local ? reinterpret_cast<unsigned long long*>(base - POS_DOUBLES + (POS_DOUBLES - 2 * MB_WORDS)) + MB_WORDS
                                           : a.gbox + ((size_t)(t > 0 ? t - 1 : 0) * 2 + 1) * MB_WORDS;
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
  for (int s = 0; s < my_sweeps; ++s) {
    const unsigned seq = (unsigned)s + 1u;
    // ---- entering row (from position t+1, produced at sweep s-1) ----
    const bool have_row = s > 0;
    if (have_row) {
      if (s - 1 < nx_sweeps) {
        if (!mb_recv(my_rbox, (unsigned)s, srow, lane, a.err)) return;
      } else {
        srow[lane] = 0.0;
        if (lane == 0) srow[CB] = 0.0;
        __syncwarp();
      }
    }
    // ---- E phase ----
    double v, tau, beta;
    if (t == 0) {
      double x = Ew[lane * WLD + (CB - 1)];
      if (have_row && lane == CB - 1) x = srow[0];
      v = warp_house(x, lane, tau, beta);
      if (lane == 0) a.e[s] = beta;
    } else {
      double Er[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) Er[k] = Ew[lane * WLD + k];
      if (have_row && lane == CB - 1) {
#pragma unroll
        for (int k = 0; k < CB - 1; ++k) Er[k] = 0.0;
        Er[CB - 1] = srow[0];
      }
      // (a) right-apply the previous reflector of this sweep
      if (!mb_recv(my_vbox, seq, svp, lane, a.err)) return;
      {
        const double taup = svp[CB];
        double dot0 = 0.0, dot1 = 0.0;
#pragma unroll
        for (int k = 0; k < CB; k += 2) {
          dot0 = fma(Er[k], svp[k], dot0);
          dot1 = fma(Er[k + 1], svp[k + 1], dot1);
        }
        const double f = taup * (dot0 + dot1);
#pragma unroll
        for (int k = 0; k < CB; ++k) Er[k] = fma(-f, svp[k], Er[k]);
      }
      // (b) reflector that annihilates the first column of the bulge
      v = warp_house(Er[0], lane, tau, beta);
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
      __syncwarp();
#pragma unroll
      for (int k = 1; k < CB; ++k) Er[k] = fma(-v, sw[k], Er[k]);
      Er[0] = lane == 0 ? beta : 0.0;
      // store the block shifted by (1,1): row r -> row r-1, column k -> column k-1; row 0 leaves (to position t-1)
      {
        double* dst = lane == 0 ? stop : (Ew + (lane - 1) * WLD - 1);
#pragma unroll
        for (int k = 0; k < CB; ++k) dst[k] = Er[k];
      }
    }
    // ---- reflector out: to position t+1 (needed at once) and to the store ----
    sv[lane] = v;
    if (s < nx_sweeps) {
      mb_send(nx_vbox, seq, lane, v);
      if (lane == 0) mb_send(nx_vbox, seq, CB, tau);
    }
    a.V2[(long long)s * a.ldv + CB * t + lane] = v;
    if (lane == 0) a.tau2[(long long)s * a.NP + t] = tau;
    __syncwarp();
    // ---- D phase: D <- H D H ----
    {
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
      double y0 = 0.0, y1 = 0.0;
#pragma unroll
      for (int i = 0; i < CB; i += 2) {
        y0 = fma(Dc[i], sv[i], y0);
        y1 = fma(Dc[i + 1], sv[i + 1], y1);
      }
      double w = tau * (y0 + y1);
      const double gamma = wsum(w * v);
      w = fma(-0.5 * tau * gamma, v, w);
      sw[lane] = w;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < CB; ++i) Dc[i] -= sv[i] * w + sw[i] * v;
      // shifted store; the new last column of E is the old first column of D below the diagonal
      if (lane >= 1) {
        double* dd = Dw + (lane - 1) * WLD - 1;
#pragma unroll
        for (int i = 1; i < CB; ++i) dd[i] = Dc[i];
        Ew[(lane - 1) * WLD + (CB - 1)] = Dc[0];
      } else {
        stop[CB] = Dc[0];
        if (t == 0) a.d[s + 1] = Dc[0];
      }
      if (t == 0 && s == my_sweeps - 1 && lane == 1) {
        Dlast0 = Dc[0];
        Dlast1 = Dc[1];
      }
    }
    __syncwarp();
    // ---- leaving row to position t-1 ----
    if (t > 0) {
      mb_send(pv_rbox, seq, lane, stop[lane]);
      if (lane == 0) mb_send(pv_rbox, seq, CB, stop[CB]);
    }
    __syncwarp();
  }
  if (t == 0 && lane == 1 && n >= 3) {
    a.d[n - 1] = Dlast1;
    a.e[n - 2] = Dlast0;
  }
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
constexpr int QC = 16;          // columns per CTA (2 per thread, 8 threads per slot)
constexpr int Q2_NT = QS * (QC / 2);   // 256 threads

struct Q2Args {
  double* X;                    // n_rows x ncols, leading dimension ldx (rows >= n are not touched)
  long long ldx;
  int n, ncols;
  const double* V2;
  long long ldv;
  const double* tau2;
  int NP;
  int q0;                       // first slot of this stage
  const double* sin;            // stream from the previous stage ((s_hi + 1) x lds), null for stage 0
  double* sout;                 // stream to the next stage, null for the last stage
  long long lds;
  int s_top;                    // first (largest) sweep index processed; s_top + 1 is a multiple of 32
};

__global__ void __launch_bounds__(Q2_NT, 1) k_q2_stage(const Q2Args a) {
  __shared__ double vbuf[2][QS * CB];          // the reflectors of the stage's slots for one sweep (8 KB each)
  __shared__ double tbuf[2][QS];               // their taus
  __shared__ double xfer[2][QS][QC];           // bottom rows handed to the next slot
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int sl = wid * 4 + (lane >> 3);        // slot within the stage
  const int cp = lane & 7;                     // column pair
  const int q = a.q0 + sl;
  const int c0 = blockIdx.x * QC + 2 * cp, c1 = c0 + 1;
  const bool ok0 = c0 < a.ncols, ok1 = c1 < a.ncols;
  const int n = a.n;
  double x0[CB], x1[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) x0[k] = x1[k] = 0.0;
  for (int i = tid; i < 2 * QS * QC; i += Q2_NT) (&xfer[0][0][0])[i] = 0.0;

  auto stage_v = [&](int s, int par) {   // cp.async the 32 x 32 reflector block of sweep s (zero for sweeps without reflectors)
    double* dst = vbuf[par];
    if (s >= 0 && s <= n - 3) {
      const double* src = a.V2 + (long long)s * a.ldv + (long long)CB * a.q0;
      const int avail = (int)min((long long)QS * CB, a.ldv - (long long)CB * a.q0);   // doubles available in this row
      for (int i = tid * 2; i < QS * CB; i += Q2_NT * 2) {
        if (i + 1 < avail) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
        } else {
          dst[i] = i < avail ? src[i] : 0.0;
          dst[i + 1] = 0.0;
        }
      }
    }
    if (tid < QS) {
      const int qq = a.q0 + tid;
      if (s >= 0 && s <= n - 3 && qq < a.NP) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(&tbuf[par][tid]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(a.tau2 + (long long)s * a.NP + qq));
      } else {
        tbuf[par][tid] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  // entering rows of the first slot of the stage come from global memory: fetched one step ahead
  auto fetch_top = [&](int s, double& e0, double& e1) {
    e0 = e1 = 0.0;
    if (sl != 0 || s < 0) return;
    if (a.sin) {
      if (s <= n - 2) {
        if (ok0) e0 = a.sin[(long long)s * a.lds + c0];
        if (ok1) e1 = a.sin[(long long)s * a.lds + c1];
      }
    } else {
      const int p = s + 1;
      if (p < n) {
        if (ok0) e0 = a.X[p + (long long)c0 * a.ldx];
        if (ok1) e1 = a.X[p + (long long)c1 * a.ldx];
      }
    }
  };

  stage_v(a.s_top, 0);
  double pre0, pre1;
  fetch_top(a.s_top, pre0, pre1);
  int par = 0;
  for (int sb = a.s_top; sb >= 0; sb -= 32) {
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int s = sb - r;
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();                               // vbuf[par], tbuf[par] and xfer[par] (written in the previous step) are visible;
      stage_v(s - 1, par ^ 1);                       // every warp has left the previous step: its buffers may be refilled
      // entering row p = s + 1 + 32 q
      double e0, e1;
      if (sl == 0) {
        e0 = pre0;
        e1 = pre1;
        fetch_top(s - 1, pre0, pre1);
      } else {
        e0 = xfer[par][sl - 1][2 * cp];
        e1 = xfer[par][sl - 1][2 * cp + 1];
      }
      // slide: logical row k lives in register (k - r) & 31; the register of the leaving bottom row takes the new top row
      x0[(32 - r) & 31] = e0;
      x1[(32 - r) & 31] = e1;
      const double tau = tbuf[par][sl];
      if (tau != 0.0) {
        const double* vv = vbuf[par] + sl * CB;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
        for (int k = 0; k < CB; k += 2) {
          const double2 v2 = *reinterpret_cast<const double2*>(vv + k);
          d0 = fma(v2.x, x0[(k - r) & 31], d0);
          d1 = fma(v2.x, x1[(k - r) & 31], d1);
          d2 = fma(v2.y, x0[(k + 1 - r) & 31], d2);
          d3 = fma(v2.y, x1[(k + 1 - r) & 31], d3);
        }
        const double f0 = -tau * (d0 + d2), f1 = -tau * (d1 + d3);
#pragma unroll
        for (int k = 0; k < CB; k += 2) {
          const double2 v2 = *reinterpret_cast<const double2*>(vv + k);
          x0[(k - r) & 31] = fma(f0, v2.x, x0[(k - r) & 31]);
          x1[(k - r) & 31] = fma(f1, v2.x, x1[(k - r) & 31]);
          x0[(k + 1 - r) & 31] = fma(f0, v2.y, x0[(k + 1 - r) & 31]);
          x1[(k + 1 - r) & 31] = fma(f1, v2.y, x1[(k + 1 - r) & 31]);
        }
      }
      // the bottom row (logical 31) leaves at the next slide
      const double b0 = x0[(31 - r) & 31], b1 = x1[(31 - r) & 31];
      xfer[par ^ 1][sl][2 * cp] = b0;
      xfer[par ^ 1][sl][2 * cp + 1] = b1;
      if (sl == QS - 1 && a.sout && s >= 1) {
        if (ok0) a.sout[(long long)(s - 1) * a.lds + c0] = b0;
        if (ok1) a.sout[(long long)(s - 1) * a.lds + c1] = b1;
      }
      par ^= 1;
    }
  }
  // final windows: sweep 0 (processed with r = 31: logical row k lives in register (k - 31) & 31), rows 1 + 32 q ..
  const int p = 1 + CB * q;
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    if (p + k < n) {
      if (ok0) a.X[(p + k) + (long long)c0 * a.ldx] = x0[(k + 1) & 31];
      if (ok1) a.X[(p + k) + (long long)c1 * a.ldx] = x1[(k + 1) & 31];
    }
  }
}

// =====================================================================================================================
// stage 1: dense -> band.  Householder QR of one 32-column sub-band panel inside a thread-block cluster
// =====================================================================================================================
// The m x 32 panel (rows r0 = j + 32 .. n-1 of columns j .. j+31) is split by rows over the CTAs of the cluster and stays
// in (distributed) shared memory.  Per column ONE cluster-wide exchange: every CTA pushes its partial dot products of
// column c (rows below the pivot) with all 32 columns into the exchange buffers of all CTAs, the owner of the pivot row
// adds that row, one hardware cluster barrier, and every CTA derives beta, tau, the update coefficients w_k (k > c) and
// column c of the compact-WY factor T (k < c: the dots with the finished reflectors) from the sums.
constexpr int PQ_NT = 256;
constexpr int PQ_MAXCS = 16;
struct PanelArgs {
  double* A;
  long long lda;
  int n, j, rp;       // rp: rows per CTA
  double* Y;          // reflector store (n x n, ldy): unit lower trapezoidal panel written at rows r0.., columns j..j+31
  long long ldy;
  double* tau;        // tau[j + k]
  double* T;          // 32 x 32 compact-WY factor of this panel (column-major)
};

__global__ void __launch_bounds__(PQ_NT, 1) k_panel_qr(const PanelArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  extern __shared__ double sm[];
  const int rp = a.rp;
  double* P = sm;                              // 32 columns x rp rows, column-major (ld = rp)
  double* ex = P + (size_t)CB * rp;            // [2][PQ_MAXCS][32] partial dots
  double* piv = ex + 2 * PQ_MAXCS * CB;        // [2][32] pivot row
  double* sw = piv + 2 * CB;                   // [32] update coefficients
  double* sz = sw + CB;                        // [32] dots with the finished reflectors
  double* sT = sz + CB;                        // [32*32] T factor
  double* sc = sT + CB * CB;                   // [4] scale, beta, tau
  double* stau = sc + 4;                       // [32]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r0 = a.j + CB, m = a.n - r0;
  const int kb = min(CB, m - 1);
  const int row_lo = rank * rp;
  const int nr = max(0, min(m, row_lo + rp) - row_lo);
  for (int idx = tid; idx < CB * rp; idx += PQ_NT) {
    const int k = idx / rp, r = idx - k * rp;
    P[idx] = r < nr ? a.A[(r0 + row_lo + r) + (long long)(a.j + k) * a.lda] : 0.0;
  }
  for (int idx = tid; idx < CB * CB; idx += PQ_NT) sT[idx] = 0.0;
  if (tid < CB) stau[tid] = 0.0;
  __syncthreads();
  cluster.sync();
  for (int c = 0; c < kb; ++c) {
    const int par = c & 1;
    // (1) partial dots of column c (rows strictly below the pivot row c) with every column
    {
      const double* pc = P + (size_t)c * rp;
      const int rstart = max(0, c + 1 - row_lo);
      double part[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* pk = P + (size_t)(wid + 8 * u) * rp;
        double acc = 0.0;
        for (int r = rstart + lane; r < nr; r += 32) acc = fma(pc[r], pk[r], acc);
        part[u] = wsum(acc);
      }
      if (lane < CS) {
        double* rex = cluster.map_shared_rank(ex, lane) + ((size_t)par * PQ_MAXCS + rank) * CB;
#pragma unroll
        for (int u = 0; u < 4; ++u) rex[wid + 8 * u] = part[u];
      }
      const int prank = c / rp;
      if (rank == prank && wid == 0) {
        const double pv = P[(size_t)lane * rp + (c - row_lo)];
        for (int dsti = 0; dsti < CS; ++dsti) cluster.map_shared_rank(piv, dsti)[par * CB + lane] = pv;
      }
    }
    cluster.sync();
    // (2) scalars of the reflector, update coefficients, column c of T  (warp 0 of every CTA, redundantly)
    if (wid == 0) {
      double g = 0.0;
      for (int src = 0; src < CS; ++src) g += ex[((size_t)par * PQ_MAXCS + src) * CB + lane];
      const double pk = piv[par * CB + lane];
      const double xn2 = __shfl_sync(0xffffffffu, g, c);
      const double alpha = __shfl_sync(0xffffffffu, pk, c);
      double tau = 0.0, beta = alpha, scale = 0.0;
      if (xn2 > 0.0) {
        const double nrm = sqrt(alpha * alpha + xn2);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      const double z = pk + scale * g;            // y_k' y_c for k < c;  v_c' P[:, k] for k > c
      sw[lane] = lane > c ? tau * z : 0.0;
      sz[lane] = lane < c ? z : 0.0;
      __syncwarp();
      // T[0:c, c] = -tau T[0:c, 0:c] z[0:c];  T[c, c] = tau
      if (lane < c) {
        double acc = 0.0;
        for (int k = lane; k < c; ++k) acc = fma(sT[lane + k * CB], sz[k], acc);
        sT[lane + c * CB] = -tau * acc;
      }
      if (lane == c) {
        sT[c + c * CB] = tau;
        stau[c] = tau;
        sc[0] = scale;
        sc[1] = beta;
      }
    }
    __syncthreads();
    // (3) apply to the local rows
    {
      const double scale = sc[0], beta = sc[1];
      double* pc = P + (size_t)c * rp;
      for (int r = tid; r < nr; r += PQ_NT) {
        const int grow = row_lo + r;
        if (grow > c) {
          const double vr = scale * pc[r];
          pc[r] = vr;
          if (vr != 0.0)
            for (int k = c + 1; k < CB; ++k) P[(size_t)k * rp + r] = fma(-vr, sw[k], P[(size_t)k * rp + r]);
        } else if (grow == c) {
          pc[r] = beta;
          for (int k = c + 1; k < CB; ++k) P[(size_t)k * rp + r] -= sw[k];
        }
      }
    }
    __syncthreads();
  }
  // outputs: R into the band part of A, Y (unit lower trapezoidal, explicit zeros) into the reflector store
  for (int idx = tid; idx < CB * rp; idx += PQ_NT) {
    const int k = idx / rp, r = idx - k * rp;
    if (r >= nr) continue;
    const int grow = row_lo + r;
    const double v = P[idx];
    double y;
    if (k >= kb) y = 0.0;
    else y = grow > k ? v : (grow == k ? 1.0 : 0.0);
    a.Y[(r0 + grow) + (long long)(a.j + k) * a.ldy] = y;
    if (k >= kb || grow <= k) a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = v;
    else a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = 0.0;
  }
  if (rank == 0) {
    for (int idx = tid; idx < CB * CB; idx += PQ_NT) a.T[idx] = sT[idx];
    if (tid < CB) a.tau[a.j + tid] = stau[tid];
  }
  cluster.sync();   // no CTA may exit while a peer can still address its shared memory
}

// W = Z0 T - 1/2 Y (T' G0 T);  P1 = [Y W], P2 = [W Y]  (m x 64 each, leading dimension ldp)
__global__ void __launch_bounds__(256) k_make_w(const double* __restrict__ Z0, long long ldz, const double* __restrict__ Y, long long ldy,
                                                const double* __restrict__ T, const double* __restrict__ G0, int m,
                                                double* __restrict__ P1, double* __restrict__ P2, long long ldp) {
  __shared__ double sT[CB * CB], sM[CB * CB], sX[CB * CB];
  __shared__ double zr[8][CB], yr[8][CB];
  const int tid = threadIdx.x;
  for (int i = tid; i < CB * CB; i += 256) sT[i] = T[i];
  __syncthreads();
  // X = G0 T, M = T' X
  for (int idx = tid; idx < CB * CB; idx += 256) {
    const int i = idx & 31, k = idx >> 5;
    double acc = 0.0;
    for (int q = 0; q <= k; ++q) acc = fma(G0[i + q * CB], sT[q + k * CB], acc);   // T upper triangular
    sX[idx] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < CB * CB; idx += 256) {
    const int i = idx & 31, k = idx >> 5;
    double acc = 0.0;
    for (int q = 0; q <= i; ++q) acc = fma(sT[q + i * CB], sX[q + k * CB], acc);
    sM[idx] = acc;
  }
  __syncthreads();
  const int rr = tid >> 5, k = tid & 31;
  for (int rb = blockIdx.x * 8; rb < m; rb += gridDim.x * 8) {
    const int r = rb + rr;
    if (r < m) {
      zr[rr][k] = Z0[r + (long long)k * ldz];
      yr[rr][k] = Y[r + (long long)k * ldy];
    }
    __syncthreads();
    if (r < m) {
      double acc = 0.0, acm = 0.0;
#pragma unroll 8
      for (int q = 0; q < CB; ++q) {
        acc = fma(zr[rr][q], sT[q + k * CB], acc);
        acm = fma(yr[rr][q], sM[q + k * CB], acm);
      }
      const double w = acc - 0.5 * acm, y = yr[rr][k];
      P1[r + (long long)k * ldp] = y;
      P1[r + (long long)(CB + k) * ldp] = w;
      P2[r + (long long)k * ldp] = w;
      P2[r + (long long)(CB + k) * ldp] = y;
    }
    __syncthreads();
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
  const size_t smem = (size_t)W * POS_DOUBLES * sizeof(double);
  TNAD_CUDA(cudaFuncSetAttribute(k_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&a};
  {
    KTimer kt(c, KF_EIG);
    TNAD_CUDA(cudaLaunchCooperativeKernel((void*)k_chase, dim3(G), dim3(32 * W), args, smem, c->stream));
  }
  c->launches++;
  int herr = 0;
  TNAD_CUDA(cudaMemcpyAsync(&herr, a.err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  if (herr) fail(TNAD_ERR_INTERNAL, "sb2st: the chase pipeline timed out");
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
    a.lds = lds; a.s_top = s_top;
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
  Tens Tb = t_alloc(c, {CB, CB}), Z0 = t_alloc(c, {n, CB}), G0 = t_alloc(c, {CB, CB});
  Tens P1 = t_alloc(c, {n, 2 * CB}), P2 = t_alloc(c, {n, 2 * CB});
  static std::atomic<unsigned long long> attr_devs{0};
  if (!((attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL)) {
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const size_t fixed = (size_t)(2 * PQ_MAXCS * CB + 2 * CB + 2 * CB + CB * CB + 4 + CB) * sizeof(double);
  const int rp_max = (int)(((232448 - 1024) - fixed) / (CB * sizeof(double)));
  for (int64_t j = 0; j + CB < n - 1; j += CB) {
    const int64_t r0 = j + CB, m = n - r0;
    int CS = (int)std::min<int64_t>(8, std::max<int64_t>(1, (m + 255) / 256));
    int rp = (int)((m + CS - 1) / CS);
    if (rp > rp_max) {
      CS = PQ_MAXCS;
      rp = (int)((m + CS - 1) / CS);
    }
    TNAD_REQUIRE(rp <= rp_max, "sy2sb: panel too tall for one thread-block cluster (n <= 12800)");
    rp = (rp + 1) & ~1;
    PanelArgs pa;
    pa.A = A; pa.lda = lda; pa.n = (int)n; pa.j = (int)j; pa.rp = rp;
    pa.Y = Yst; pa.ldy = ldy; pa.tau = tau1; pa.T = Tb.p;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(PQ_NT);
    cfg.dynamicSmemBytes = fixed + (size_t)CB * rp * sizeof(double);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    {
      KTimer kt(c, KF_EIG);
      TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr, pa));
    }
    c->launches++;
    // two-sided update of the trailing matrix: Z0 = A22 Y, G0 = Y' Z0, W = Z0 T - Y (T' G0 T) / 2, A22 -= Y W' + W Y'
    Tens A22 = bview(A + r0 + r0 * lda, m, m, lda);
    Tens Yp = bview(Yst + r0 + j * ldy, m, CB, ldy);
    Tens Zv = bview(Z0.p, m, CB, n);
    contract(c, "ik,kj->ij", A22, Yp, Zv);
    contract(c, "ki,kj->ij", Yp, Zv, G0);
    const int nb = (int)std::min<int64_t>((m + 7) / 8, 2 * c->num_sms);
    k_make_w<<<nb, 256, 0, st>>>(Z0.p, n, Yp.p, ldy, Tb.p, G0.p, (int)m, P1.p, P2.p, n);
    LAUNCH_CHECK(c);
    Tens L = bview(P1.p, m, 2 * CB, n), R = bview(P2.p, m, 2 * CB, n);
    contract(c, "ik,jk->ij", L, R, A22, -1.0, 1.0);
  }
}

void extract_band(tnad_ctx* c, const double* A, int64_t lda, int64_t n, double* AB, int64_t ldab) {
  const long long total = (long long)n * (CB + 1);
  const int nb = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 8));
  k_extract_band<<<nb, 256, 0, c->stream>>>(A, lda, (int)n, AB, (int)ldab);
  LAUNCH_CHECK(c);
}

}  // namespace tnad
