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
#include <functional>
#include <cooperative_groups.h>
#include <algorithm>
#include <chrono>

namespace cg = cooperative_groups;

namespace tnad {

namespace {

constexpr int CB = 32;     // half bandwidth = reflector length of the chase
constexpr int WLD = 33;    // leading dimension of the window arrays in shared memory

// Programmatic dependent launch (stage 1 is a chain of six short kernels per panel: their launch latencies and set-up
// overlap the predecessor's tail): nothing before the wait touches global memory, so the semantics are those of a plain
// stream-ordered launch; the trigger comes first so that the successor is resident when this grid retires.
#define PDL_ENTRY()                                              \
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  \
  asm volatile("griddepcontrol.wait;" ::: "memory")

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
// 1 / sqrt(x) and 1 / x for normal positive x without the library's special-case paths: hardware seed (MUFU.RSQ64H /
// MUFU.RCP64H, about 20 bits) + Newton steps; the reflector scalars sit on the critical path of the chase pipeline
__device__ __forceinline__ double fast_rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double fast_rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  return y;
}

// reflector scalars from alpha = x_0 and xn2 = sum_{r >= 1} x_r^2: v_r = x_r * scale (v_0 = 1), H = I - tau v v'
__device__ __forceinline__ void house_scalars(double alpha, double xn2, double& tau, double& beta, double& scale) {
  const double s2 = fma(alpha, alpha, xn2);
  if (xn2 == 0.0 || !(s2 > 1e-290 && s2 < 1e290)) {          // nothing to annihilate (or a degenerate scale): H = I
    tau = 0.0;
    beta = alpha;
    scale = 0.0;
    return;
  }
  // one reciprocal square root and one reciprocal instead of a square root and two divisions:
  // beta = -sign(alpha) |x|, tau = (beta - alpha) / beta = 1 + |alpha| / |x|, v = x / (alpha - beta) = sign(alpha) x / (|alpha| + |x|)
  const double rn = fast_rsqrt_pos(s2);
  const double nrm = s2 * rn;
  const double aa = fabs(alpha);
  beta = alpha >= 0.0 ? -nrm : nrm;
  tau = fma(aa, rn, 1.0);
  const double rc = fast_rcp_pos(aa + nrm);
  scale = alpha >= 0.0 ? rc : -rc;
}
__device__ __forceinline__ double warp_house(double x, int lane, double& tau, double& beta) {
  const double xn2 = wsum(lane >= 1 ? x * x : 0.0);
  const double alpha = __shfl_sync(0xffffffffu, x, 0);
  double scale;
  house_scalars(alpha, xn2, tau, beta, scale);
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
  unsigned long long* gbox;      // NP x 3 x MB_WORDS, zero-initialised: [t][0] = reflector box, [t][1..2] = row boxes
  int* err;
  long long* prof;               // optional: per position 6 cycle counters (wait row, wait v, E phase, D phase, sends, total)
};

// Two warps per position: the E warp owns the bulge block (right-apply the previous reflector, new reflector, column
// sums, left-apply), the D warp the diagonal block (D <- H D H).  The reflector goes from the E warp to the D warp
// through a local box; the D warp hands back the new last column of E (the old first column of D) and raises a flag.
// per-position shared memory (doubles): Ew (1 + 32*33 + 1), Dw (32*33), 7 scratch vectors of 34, mailboxes:
//   vbox (reflector from position t-1), rbox[2] (row from position t+1, double buffered by sweep parity: the D warp may
//   lag the E warp by one hop), dbox (reflector E warp -> D warp), and a word the D warp uses to publish its progress
constexpr int POS_BOXES = 4 * MB_WORDS + 4;   // 4 mailboxes + the progress words of the D warp, of E1 and of E2 (two)
constexpr int POS_DOUBLES = (2 + CB * WLD) + CB * WLD + CB * WLD + 7 * 34 + POS_BOXES;   // Ew, Dw, transpose scratch, vectors, boxes
constexpr int BOXOFF = POS_DOUBLES - POS_BOXES;      // offset of a position's boxes inside its shared-memory slice

// the E warp needs double 0 of the row message only, the D warp doubles 1 .. 32 (lane l <-> double 1 + l)
__device__ __forceinline__ bool mb_recv_one(unsigned long long* box, unsigned seq, int idx, double& out, int* err) {
  volatile unsigned long long* vb = box;
  unsigned long long w0, w1;
  long long t0 = 0;
  unsigned spins = 0;
  for (;;) {
    w0 = vb[2 * idx];
    w1 = vb[2 * idx + 1];
    const bool ok = (unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq;
    if (__all_sync(0xffffffffu, ok)) break;
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      const bool to = (now - t0 > 4000000000ll) || (*(volatile int*)err != 0);
      if (__any_sync(0xffffffffu, to)) {
        if ((threadIdx.x & 31) == 0) atomicExch(err, 1);
        return false;
      }
    }
  }
  out = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}
__device__ __forceinline__ bool flag_wait(unsigned long long* word, unsigned long long want, int* err) {
  volatile unsigned long long* vw = word;
  long long t0 = 0;
  unsigned spins = 0;
  while (*vw < want) {
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if ((now - t0 > 4000000000ll) || (*(volatile int*)err != 0)) {
        atomicExch(err, 1);
        return false;
      }
    }
  }
  return true;
}

__global__ void __launch_bounds__(384, 1) k_chase(const ChaseArgs a) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pw = wid / 3;                         // position within the CTA
  const int role = wid - 3 * pw;                  // 0: E1 (reflector: the critical path), 1: D (diagonal block), 2: E2 (column sums, left-apply)
  const bool is_d = role == 1;
  // zero the local mailboxes of every position of this CTA before any neighbour may write into them
  for (int i = threadIdx.x; i < a.W * POS_DOUBLES; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  cluster.sync();
  const int t = blockIdx.x * a.W + pw;
  const int n = a.n;
  if (t < a.NP) {
  double* base = sm + (size_t)pw * POS_DOUBLES;
  double* Ew = base + 1;                       // one double of head room: the shifted store of row 1 touches Ew[-1]
  double* Dw = base + 2 + CB * WLD;
  double* Tw = Dw + CB * WLD;                  // 32 x 33 scratch of the E warp (column sums by transposition)
  double* sv = Tw + CB * WLD;                  // own reflector, E warp's copy (34)
  double* svp = sv + 34;                       // previous reflector of the sweep + tau (34)
  double* sw = svp + 34;                       // E-phase coefficients w_k (34)
  double* stop = sw + 34;                      // leaving top row, E part (34)
  double* svd = stop + 34;                     // own reflector + tau, D warp's copy (34)
  double* sd = svd + 34;                       // D-phase vector w (34)
  double* srow = sd + 34;                      // entering row, D part (34)
  unsigned long long* lbox = reinterpret_cast<unsigned long long*>(base + BOXOFF);
  unsigned long long* my_vbox_l = lbox;                         // reflector from t-1
  unsigned long long* my_rbox_l = lbox + MB_WORDS;              // [2] row from t+1
  unsigned long long* dbox = lbox + 3 * MB_WORDS;               // reflector E -> D
  unsigned long long* dflag = lbox + 4 * MB_WORDS;              // D warp: number of hops whose E column is handed back
  unsigned long long* e1flag = dflag + 1;                       // E1: number of hops whose right-applied block sits in Tw
  unsigned long long* wflag = dflag + 2;                        // E2: number of hops whose left-apply coefficients w sit in sw
  unsigned long long* e2flag = dflag + 3;                       // E2: number of hops it has finished with Tw
  // A mailbox lives with its READER: in the reader's shared memory when the writer sits in the same CTA or in the same
  // cluster (then the writer pushes through distributed shared memory), in global memory (L2) between clusters.
  const bool prev_cta = t > 0 && pw == 0, next_cta = t + 1 < a.NP && pw + 1 == a.W;
  const bool prev_far = prev_cta && crank == 0, next_far = next_cta && crank + 1 == csize;   // neighbour in another cluster
  unsigned long long* my_vbox = prev_far ? a.gbox + ((size_t)t * 3 + 0) * MB_WORDS : my_vbox_l;
  unsigned long long* my_rbox = next_far ? a.gbox + ((size_t)t * 3 + 1) * MB_WORDS : my_rbox_l;     // two consecutive boxes
  unsigned long long* nx_vbox;   // reflector box of position t+1
  if (!next_cta) nx_vbox = reinterpret_cast<unsigned long long*>(base + POS_DOUBLES + BOXOFF);
  else if (next_far) nx_vbox = a.gbox + ((size_t)(t + 1) * 3 + 0) * MB_WORDS;
  else nx_vbox = reinterpret_cast<unsigned long long*>(cluster.map_shared_rank(sm + BOXOFF, crank + 1));
  unsigned long long* pv_rbox;   // row boxes of position t-1
  if (!prev_cta) pv_rbox = reinterpret_cast<unsigned long long*>(base - POS_DOUBLES + BOXOFF) + MB_WORDS;
  else if (prev_far) pv_rbox = a.gbox + ((size_t)(t > 0 ? t - 1 : 0) * 3 + 1) * MB_WORDS;
  else pv_rbox = reinterpret_cast<unsigned long long*>(cluster.map_shared_rank(sm + (size_t)(a.W - 1) * POS_DOUBLES + BOXOFF, crank - 1)) + MB_WORDS;
  const int my_sweeps = min(n - 2, n - 1 - CB * t);
  const int nx_sweeps = t + 1 < a.NP ? min(n - 2, n - 1 - CB * (t + 1)) : 0;

  // ---- initial window (sweep 0): rows p .. p+31, p = 1 + 32 t; zero beyond the matrix (each warp loads its block) ----
  {
    const int p = 1 + CB * t;
    const int i = p + lane;   // my row
#pragma unroll 4
    for (int k = 0; k < CB && role != 2; ++k) {
      double val = 0.0;
      if (i < n) {
        if (!is_d) {
          const int je = p - CB + k;           // E column
          if (t >= 1) {
            if (i - je <= CB) val = a.AB[(i - je) + (long long)je * a.ldab];
          } else if (k == CB - 1) {
            val = a.AB[i];                     // position 0: column 0, rows 1 .. 32
          }
        } else {
          const int jd = p + k;                // D column
          if (jd < n) val = (i >= jd) ? a.AB[(i - jd) + (long long)jd * a.ldab] : a.AB[(jd - i) + (long long)i * a.ldab];
        }
      }
      (is_d ? Dw : Ew)[lane * WLD + k] = val;
    }
    if (t == 0 && lane == 0 && role == 0) a.d[0] = a.AB[0];
  }
  // the warps of the position have written their blocks (named barrier: 96 threads)
  asm volatile("bar.sync %0, 96;" ::"r"(1 + pw) : "memory");

  long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool prof = a.prof != nullptr;
  const long long tstart = prof ? clock64() : 0;
#define CH_TICK(slot)                     \
    if (prof) {                           \
      const long long _n = clock64();     \
      pc[slot] += _n - tk;                \
      tk = _n;                            \
    }
  if (role == 0) {
    // =============================================== E1 warp ==============================================
    // right-apply of the previous reflector, the new reflector, its sends: everything position t+1 and t-1 wait for.
    // The right-applied block goes to Tw for the E2 warp, which owns the rest of the hop.
    for (int s = 0; s < my_sweeps; ++s) {
      const unsigned seq = (unsigned)s + 1u;
      long long tk = prof ? clock64() : 0;
      // the D warp has handed back column 31 of my block for this sweep (its hop s-1), E2 the left-apply coefficients
      if (s > 0 && !flag_wait(dflag, (unsigned long long)s, a.err)) break;
      if (s > 0 && t > 0 && !flag_wait(wflag, (unsigned long long)s, a.err)) break;
      __threadfence_block();                     // acquire: what the two warps stored before raising their flags
      CH_TICK(0)
      double v, tau, beta;
      double Er[CB];
      double fr = 0.0;                           // coefficient of my row in the right-apply
      const bool row_enters = s > 0 && s - 1 < nx_sweeps;
      if (t == 0) {
        double x = Ew[lane * WLD + (CB - 1)];
        double ent = 0.0;
        if (row_enters && !mb_recv_one(my_rbox + ((s - 1) & 1) * MB_WORDS, (unsigned)s, 0, ent, a.err)) break;
        if (s > 0 && lane == CB - 1) x = ent;
        v = warp_house(x, lane, tau, beta);
        if (lane == 0) a.e[s] = beta;
      } else {
        // Everything that does not depend on beta of position t+1 (= E'[31][31], the only non-zero of the entering row 31)
        // happens BEFORE the wait for it: the block, the previous reflector, the dot products and reflector inputs of
        // rows 0 .. 30 and their sum of squares.  After the wait only row 31's entry and the scalars are left, so this
        // position's share of the loop t -> t+1 -> t is a dozen dependent operations instead of a 32-term dot product
        // and a five-level shuffle reduction.
        // The block of this sweep, straight into registers: row r of it is row r + 1 of the previous sweep's right-applied
        // block (Tw, written by this warp) after the left-apply with the previous reflector (sv) and the coefficients w
        // (sw, from E2), shifted by one column; column 31 is what the D warp handed back; row 31 enters as zeros + beta.
        if (s == 0) {
#pragma unroll
          for (int k = 0; k < CB; ++k) Er[k] = Ew[lane * WLD + k];
        } else if (lane == CB - 1) {
#pragma unroll
          for (int k = 0; k < CB; ++k) Er[k] = 0.0;
        } else {
          const double vr = sv[lane + 1];
          double tw[CB], wk[CB];                   // all loads first
#pragma unroll
          for (int k = 0; k < CB - 1; ++k) {
            tw[k] = Tw[(lane + 1) * WLD + k + 1];
            wk[k] = sw[k + 1];
          }
#pragma unroll
          for (int k = 0; k < CB - 1; ++k) Er[k] = fma(-vr, wk[k], tw[k]);
          Er[CB - 1] = Ew[lane * WLD + (CB - 1)];
        }
        CH_TICK(2)
        if (!mb_recv(my_vbox, seq, svp, lane, a.err)) break;
        CH_TICK(1)
        const double taup = svp[CB];
        double dot0 = 0.0, dot1 = 0.0, dot2 = 0.0, dot3 = 0.0;
#pragma unroll
        for (int k = 0; k < CB; k += 4) {
          dot0 = fma(Er[k], svp[k], dot0);
          dot1 = fma(Er[k + 1], svp[k + 1], dot1);
          dot2 = fma(Er[k + 2], svp[k + 2], dot2);
          dot3 = fma(Er[k + 3], svp[k + 3], dot3);
        }
        fr = taup * ((dot0 + dot1) + (dot2 + dot3));
        double x = fma(-fr, svp[0], Er[0]);        // column 0 after the right-apply (rows 0 .. 30; row 31 follows)
        const double pre = wsum(lane >= 1 ? x * x : 0.0);
        const double alpha = __shfl_sync(0xffffffffu, x, 0);
        double xn2 = pre;
        CH_TICK(6)
        if (s > 0) {
          double ent = 0.0;
          if (row_enters && !mb_recv_one(my_rbox + ((s - 1) & 1) * MB_WORDS, (unsigned)s, 0, ent, a.err)) break;
          CH_TICK(7)
          const double f31 = taup * ent * svp[CB - 1], x31 = -f31 * svp[0];
          xn2 = fma(x31, x31, pre);
          if (lane == CB - 1) {
            Er[CB - 1] = ent;
            fr = f31;
            x = x31;
          }
        }
        double scale;
        house_scalars(alpha, xn2, tau, beta, scale);
        Er[0] = x;
        v = lane == 0 ? 1.0 : x * scale;
        // beta is E'[31][31] of position t-1's next sweep -- all its E1 warp needs of the leaving row (the loop
        // t -> t+1 -> t closes here)
        if (lane == 0) mb_send(pv_rbox + (s & 1) * MB_WORDS, seq, 0, beta);
      }
      // reflector out at once: to position t+1 and to my D warp
      if (s < nx_sweeps) {
        mb_send(nx_vbox, seq, lane, v);
        if (lane == 0) mb_send(nx_vbox, seq, CB, tau);
      }
      mb_send(dbox, seq, lane, v);
      if (lane == 0) mb_send(dbox, seq, CB, tau);
      CH_TICK(8)
      if (t > 0) {
        // the rest of the right-apply, then the block, the reflector and tau to the E2 warp
#pragma unroll
        for (int k = 1; k < CB; ++k) Er[k] = fma(-fr, svp[k], Er[k]);
        if (s > 0 && !flag_wait(e2flag, (unsigned long long)s, a.err)) break;      // E2 has read row 0 of Tw (the leaving row)
#pragma unroll
        for (int k = 0; k < CB; ++k) Tw[lane * WLD + k] = Er[k];
        sv[lane] = v;
        if (lane == 0) sv[CB] = tau;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *(volatile unsigned long long*)e1flag = (unsigned long long)(s + 1);
        }
      }
      CH_TICK(4)
      a.V2[(long long)s * a.ldv + CB * t + lane] = v;
      if (lane == 0) a.tau2[(long long)s * a.NP + t] = tau;
      CH_TICK(3)
    }
  } else if (role == 2) {
    // =============================================== E2 warp ==============================================
    // column sums c_k = sum_r v_r E[r][k] (lane k walks down column k of Tw) -> the left-apply coefficients w for E1, and
    // the E part of the row that leaves to position t-1
    if (t > 0)
      for (int s = 0; s < my_sweeps; ++s) {
        const unsigned seq = (unsigned)s + 1u;
        long long tk = prof ? clock64() : 0;
        if (!flag_wait(e1flag, (unsigned long long)(s + 1), a.err)) break;
        __threadfence_block();
        CH_TICK(0)
        const double tau = sv[CB];
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
        for (int r = 0; r < CB; r += 4) {
          c0 = fma(sv[r], Tw[r * WLD + lane], c0);
          c1 = fma(sv[r + 1], Tw[(r + 1) * WLD + lane], c1);
          c2 = fma(sv[r + 2], Tw[(r + 2) * WLD + lane], c2);
          c3 = fma(sv[r + 3], Tw[(r + 3) * WLD + lane], c3);
        }
        const double wl = lane == 0 ? 0.0 : tau * ((c0 + c1) + (c2 + c3));
        sw[lane] = wl;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *(volatile unsigned long long*)wflag = (unsigned long long)(s + 1);      // E1 may build the block of sweep s + 1
        }
        CH_TICK(2)
        // v_0 = 1: the leaving row is E[0][k] - w_k, straight from lane k (double 0, beta, went out from E1)
        if (lane >= 1) mb_send(pv_rbox + (s & 1) * MB_WORDS, seq, lane, Tw[lane] - wl);
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *(volatile unsigned long long*)e2flag = (unsigned long long)(s + 1);
        }
        CH_TICK(3)
      }
  } else {
    // =============================================== D warp ===============================================
    double Dlast0 = 0.0, Dlast1 = 0.0;   // D[1][0], D[1][1] of position 0 after its last hop
    for (int s = 0; s < my_sweeps; ++s) {
      const unsigned seq = (unsigned)s + 1u;
      long long tk = prof ? clock64() : 0;
      double Dc[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i) Dc[i] = Dw[lane * WLD + i];
      if (s > 0) {
        double mine = 0.0;                       // double 1 + lane of the entering row
        if (s - 1 < nx_sweeps) {
          if (!mb_recv_one(my_rbox + ((s - 1) & 1) * MB_WORDS, (unsigned)s, 1 + lane, mine, a.err)) break;
        }
        srow[lane] = mine;
        __syncwarp();
        if (lane == CB - 1) {
#pragma unroll
          for (int i = 0; i < CB; ++i) Dc[i] = srow[i];
        } else {
          Dc[CB - 1] = mine;
        }
      }
      CH_TICK(0)
      if (!mb_recv(dbox, seq, svd, lane, a.err)) break;
      CH_TICK(1)
      const double v = svd[lane], tau = svd[CB];
      double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
      for (int i = 0; i < CB; i += 4) {
        y0 = fma(Dc[i], svd[i], y0);
        y1 = fma(Dc[i + 1], svd[i + 1], y1);
        y2 = fma(Dc[i + 2], svd[i + 2], y2);
        y3 = fma(Dc[i + 3], svd[i + 3], y3);
      }
      double wd = tau * ((y0 + y1) + (y2 + y3));
      const double gamma = wsum(wd * v);
      wd = fma(-0.5 * tau * gamma, v, wd);
      sd[lane] = wd;
      __syncwarp();
      // row 0 / column 0 of the updated block first: D[0][0] leaves to position t-1, D[0][1..31] becomes the new last
      // column of E (handed back to the E warp)
      Dc[0] -= svd[0] * wd + sd[0] * v;
      if (lane == 0) {
        if (t > 0) mb_send(pv_rbox + (s & 1) * MB_WORDS, seq, CB, Dc[0]);
        else a.d[s + 1] = Dc[0];
      } else {
        Ew[(lane - 1) * WLD + (CB - 1)] = Dc[0];
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        *(volatile unsigned long long*)dflag = (unsigned long long)(s + 1);
      }
      CH_TICK(3)
#pragma unroll
      for (int i = 1; i < CB; ++i) Dc[i] -= svd[i] * wd + sd[i] * v;
      if (lane >= 1) {
        double* dd = Dw + (lane - 1) * WLD - 1;
#pragma unroll
        for (int i = 1; i < CB; ++i) dd[i] = Dc[i];
      }
      if (t == 0 && s == my_sweeps - 1 && lane == 1) {
        Dlast0 = Dc[0];
        Dlast1 = Dc[1];
      }
      __syncwarp();
      CH_TICK(4)
    }
    if (t == 0 && lane == 1 && n >= 3) {
      a.d[n - 1] = Dlast1;
      a.e[n - 2] = Dlast0;
    }
  }
  if (prof && lane == 0) {
    pc[5] = clock64() - tstart;
    for (int i = 0; i < 10; ++i) a.prof[((long long)t * 3 + role) * 10 + i] = pc[i];
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
constexpr int QC = 16;          // columns per CTA
constexpr int QH = 16;          // rows per thread: a slot's 32-row window is split over two threads (upper / lower half)
constexpr int Q2_NT = QS * 2 * (QC / 2);   // 512 threads: (slot, half, column pair)

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
  int ident;                    // X is the identity on entry: a column block is untouched by the sweeps below its last unit row
};

// Thread = (slot, half, column pair): 16 rows of 2 columns in registers, its half of the reflector cached in registers for
// both passes (dot, update): 2 bytes of shared-memory traffic per FMA (the load/store unit moves 128 B/clk/SM, the FP64
// pipe executes 64 FMA/clk/SM).  The two halves of a slot sit in lanes l and l ^ 8: one shuffle combines the dots, one
// hands the row that crosses the middle of the window.
__global__ void __launch_bounds__(Q2_NT, 1) k_q2_stage(const Q2Args a) {
  constexpr int NB = 4;                           // ring of reflector buffers: prefetch distance 3 sweeps
  constexpr int UNR = 4;                          // sweeps per physical register shift
  __shared__ double vbuf[NB][QS * CB];            // the reflectors of the stage's slots for one sweep (8 KB each)
  __shared__ double tbuf[NB][QS];                 // their taus
  __shared__ double xfer[2][QS][QC];              // bottom rows handed to the next slot
  __shared__ double topbuf[NB][QC];               // rows entering the first slot of the stage (from the stream / the matrix)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int sl = wid * 2 + (lane >> 4);           // slot within the stage
  const int h = (lane >> 3) & 1;                  // 0: rows 0..15 of the window, 1: rows 16..31
  const int cp = lane & 7;                        // column pair
  const int q = a.q0 + sl;
  const int c0 = blockIdx.x * QC + 2 * cp, c1 = c0 + 1;
  const bool ok0 = c0 < a.ncols, ok1 = c1 < a.ncols;
  const int n = a.n;
  // at sub-step r of a block of UNR sweeps the logical row k of my half lives in x[k + UNR - 1 - r]; every UNR sweeps the
  // registers are shifted up by UNR (a fully static circular buffer would need the sweep loop unrolled 32 times)
  double x0[QH + UNR], x1[QH + UNR];
#pragma unroll
  for (int k = 0; k < QH + UNR; ++k) x0[k] = x1[k] = 0.0;
  for (int i = tid; i < 2 * QS * QC; i += Q2_NT) (&xfer[0][0][0])[i] = 0.0;

  // Staging of sweep s: the 32 x 32 reflector block (16-byte cp.async per thread), the 32 taus, the 16 rows that enter the
  // first slot.  Everything that does not depend on s is hoisted: per thread one source pointer per stream, moved by a
  // constant per sweep.
  const int v_avail = (int)min((long long)QS * CB, a.ldv - (long long)CB * a.q0);
  const bool v_vec = 2 * tid + 1 < v_avail, v_one = 2 * tid < v_avail;
  const double* v_src0 = a.V2 + (long long)CB * a.q0 + 2 * tid;                        // + s * ldv
  const bool t_thr = tid < QS && a.q0 + tid < a.NP;
  const double* t_src0 = a.tau2 + a.q0 + (tid < QS ? tid : 0);                          // + s * NP
  const bool top_thr = tid >= 64 && tid < 64 + QC && blockIdx.x * QC + (tid - 64) < a.ncols;
  const int top_c = tid - 64;
  const double* top_src0 = a.sin ? a.sin + blockIdx.x * QC + top_c : a.X + 1 + (long long)(blockIdx.x * QC + top_c) * a.ldx;
  const long long top_step = a.sin ? a.lds : 1;                                         // + s * top_step
  const int top_smax = a.sin ? n - 2 : n - 2;                                           // row s + 1 <= n - 1 / stream rows 0 .. n-2
  auto stage_v = [&](int s, int buf) {
    if (s >= 0) {
      if (s <= n - 3) {
        double* dst = vbuf[buf] + 2 * tid;
        const double* src = v_src0 + (long long)s * a.ldv;
        if (v_vec) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src));
        } else {
          dst[0] = v_one ? src[0] : 0.0;
          dst[1] = 0.0;
        }
        if (t_thr) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(&tbuf[buf][tid]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(t_src0 + (long long)s * a.NP));
        }
      } else if (tid < QS) {
        tbuf[buf][tid] = 0.0;
      }
      if (top_thr) {
        if (s <= top_smax) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(&topbuf[buf][top_c]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(top_src0 + (long long)s * top_step));
        } else {
          topbuf[buf][top_c] = 0.0;
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  for (int i = tid; i < NB * QS; i += Q2_NT) (&tbuf[0][0])[i] = 0.0;       // slots beyond NP keep tau = 0
  for (int i = tid; i < NB * QC; i += Q2_NT) (&topbuf[0][0])[i] = 0.0;     // columns beyond ncols stay 0
  __syncthreads();

  // X = I: columns c0 .. c0 + 15 are zero below row c0 + 15 and sweep s only touches rows >= s + 1, so this column block
  // starts at sweep c0 + 14 (rounded up to the register-shift period); every later stage starts one period earlier than
  // the stage above it, so that the stream rows it consumes have been written (half of the work of the pass is skipped)
  const int s_first = a.ident ? min(a.s_top, (blockIdx.x * QC + QC + 2) / UNR * UNR - 1 + UNR - UNR * (a.q0 / QS + 1)) : a.s_top;
#pragma unroll
  for (int i = 0; i < NB - 1; ++i) stage_v(s_first - i, i);
  int par = 0;
  for (int sb = s_first; sb >= 0; sb -= UNR) {
#pragma unroll
    for (int r = 0; r < UNR; ++r) {
      const int s = sb - r;
      const int buf = r;                             // (s_top - s) mod NB == r because s_top + 1 and UNR are multiples of NB
      asm volatile("cp.async.wait_group %0;\n" ::"n"(NB - 2));
      __syncthreads();                               // vbuf[buf], tbuf[buf] and xfer[par] (written in the previous step) are visible;
      stage_v(s - (NB - 1), (r + NB - 1) % NB);      // every warp has left the previous step: its buffer may be refilled
      const int off = UNR - 1 - r;                   // logical row k of my half <-> register k + off
      // slide the window up by one row: the row that leaves my half sits in register QH + off
      const double up0 = __shfl_xor_sync(0xffffffffu, x0[QH + off], 8);
      const double up1 = __shfl_xor_sync(0xffffffffu, x1[QH + off], 8);
      double e0, e1;
      if (h == 1) {                                  // from the upper half of the same slot
        e0 = up0;
        e1 = up1;
      } else if (sl == 0) {                          // from the previous stage / the matrix
        e0 = topbuf[buf][2 * cp];
        e1 = topbuf[buf][2 * cp + 1];
      } else {                                       // from the slot above
        e0 = xfer[par][sl - 1][2 * cp];
        e1 = xfer[par][sl - 1][2 * cp + 1];
      }
      x0[off] = e0;
      x1[off] = e1;
      const double tau = tbuf[buf][sl];
      if (tau != 0.0) {                              // uniform over the 16 lanes of a slot
        const double* vv = vbuf[buf] + sl * CB + h * QH;
        double vh[QH];
#pragma unroll
        for (int k = 0; k < QH; k += 2) {
          const double2 v2 = *reinterpret_cast<const double2*>(vv + k);
          vh[k] = v2.x;
          vh[k + 1] = v2.y;
        }
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
        for (int k = 0; k < QH; k += 2) {
          d0 = fma(vh[k], x0[k + off], d0);
          d1 = fma(vh[k], x1[k + off], d1);
          d2 = fma(vh[k + 1], x0[k + 1 + off], d2);
          d3 = fma(vh[k + 1], x1[k + 1 + off], d3);
        }
        d0 += d2;
        d1 += d3;
        d0 += __shfl_xor_sync(0x0000ffffu << (lane & 16), d0, 8);
        d1 += __shfl_xor_sync(0x0000ffffu << (lane & 16), d1, 8);
        const double f0 = -tau * d0, f1 = -tau * d1;
#pragma unroll
        for (int k = 0; k < QH; ++k) {
          x0[k + off] = fma(f0, vh[k], x0[k + off]);
          x1[k + off] = fma(f1, vh[k], x1[k + off]);
        }
      }
      // the bottom row of the window (lower half, logical row 15 of that half) leaves the slot at the next slide
      if (h == 1) {
        const double b0 = x0[QH - 1 + off], b1 = x1[QH - 1 + off];
        xfer[par ^ 1][sl][2 * cp] = b0;
        xfer[par ^ 1][sl][2 * cp + 1] = b1;
        if (sl == QS - 1 && a.sout && s >= 1) {
          if (ok0) a.sout[(long long)(s - 1) * a.lds + c0] = b0;
          if (ok1) a.sout[(long long)(s - 1) * a.lds + c1] = b1;
        }
      }
      par ^= 1;
    }
    // after UNR sweeps logical row k sits in register k: move everything up by UNR for the next block
#pragma unroll
    for (int k = QH - 1; k >= 0; --k) {
      x0[k + UNR] = x0[k];
      x1[k + UNR] = x1[k];
    }
  }
  // final windows (state after sweep 0, already shifted: logical row k in register k + UNR), rows 1 + 32 q + 16 h ..
  const int p = 1 + CB * q + QH * h;
#pragma unroll
  for (int k = 0; k < QH; ++k) {
    if (p + k < n) {
      if (ok0) a.X[(p + k) + (long long)c0 * a.ldx] = x0[k + UNR];
      if (ok1) a.X[(p + k) + (long long)c1 * a.ldx] = x1[k + UNR];
    }
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
  const int* only_if; // optional: run only when *only_if != 0 (the Gram-based kernel asked for the fallback)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(double* sdst, const double* g, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa16(double* sdst, const double* g, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------
// Fast path of the panel factorisation: Householder QR driven by the Gram matrix.
//   G = P'P (full columns) is invariant under the reflectors, so the dot products over the rows BELOW the pivot that
//   Householder needs at column c are  g_ck = G[c][k] - sum_{r<c} R[r][c] R[r][k] - P[c][c] P[c][k]  -- they follow from
//   G and from the top 32 x 32 block of the panel alone.  Every CTA holds both (one cluster-wide reduction of the 32 x 32
//   partial Gram matrices at the start), so the whole sequence of reflector scalars and update coefficients is computed
//   by warp 0 of every CTA redundantly, without any further exchange, and the rows are then transformed independently.
//   The subtraction cancels when column c lies (almost) in the span of the previous ones: if the remaining part carries
//   less than PQ_THETA of the column's squared norm the panel is handed to the exchange-per-column kernel (k_panel_qr),
//   which recomputes every dot product from the data (flag *fb = 1; nothing has been written at that point).
//   T comes from the Gram matrix of the finished reflectors (second reduction), as in dlarft.
// ---------------------------------------------------------------------------------------------------------------------
constexpr double PQ_THETA = 0.0625;
constexpr int PG_LDS = 33;
template <int NRL>
__global__ void __launch_bounds__(PQ_NT, 1) k_panel_gram(const PanelArgs a, int* fb) {
  PDL_ENTRY();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();   // CS is a power of two <= 16
  extern __shared__ double sm[];
  constexpr int ROWS = PQ_NT * NRL, LDR = ROWS + 4;
  double* Ys = sm;                               // [32][LDR] column-major copy of my rows (DMMA operand)
  double* Gl = Ys + (size_t)CB * LDR;            // [32][33] local partial Gram
  double* recv = Gl + CB * PG_LDS;               // [CS][RS * 32] slices received from the other CTAs (1024 doubles)
  double* Gs = recv + CB * CB;                   // [32][33] full Gram
  double* tops = Gs + CB * PG_LDS;               // [32][33] top block of the panel (rows 0..31), updated in place
  double* Rs = tops + CB * PG_LDS;               // [32][33] finished rows of R
  double* wtab = Rs + CB * PG_LDS;               // [32][32] update coefficients per column
  double* sctab = wtab + CB * CB;                // [32] scale
  double* betab = sctab + CB;                    // [32] beta
  double* tautab = betab + CB;                   // [32] tau
  double* sT = tautab + CB;                      // [32][32]
  int* flag = reinterpret_cast<int*>(sT + CB * CB);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int r0 = a.j + CB, m = a.n - r0;
  const int rp = a.rp;                           // == ROWS
  const int row_lo = rank * rp;
  const int nr = max(0, min(m, row_lo + rp) - row_lo);
  const int RS = CB / CS;                        // Gram rows owned by each rank in the reduction
  double pr[NRL][CB];
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    const int r = i * PQ_NT + tid;
    const double* src = a.A + (r0 + row_lo + r) + (long long)a.j * a.lda;
#pragma unroll
    for (int k = 0; k < CB; ++k) pr[i][k] = r < nr ? src[(long long)k * a.lda] : 0.0;
  }
  if (tid == 0) *flag = 0;
  long long gpc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool gprof = a.prof != nullptr;
  long long gtk = gprof ? clock64() : 0;
#define PG_TICK(slot)                  \
  if (gprof) {                         \
    const long long _n = clock64();    \
    gpc[slot] += _n - gtk;             \
    gtk = _n;                          \
  }

  // Gram matrix of the rows staged in Ys: warp w computes the 8 x 8 tiles (w >> 1, 2 (w & 1) + {0, 1}) over all rows
  auto gram_local = [&]() {
    const int I = wid >> 1, J0 = 2 * (wid & 1);
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    const double* pa = Ys + (size_t)(8 * I + g8) * LDR + t4;
    const double* pb0 = Ys + (size_t)(8 * J0 + g8) * LDR + t4;
    const double* pb1 = pb0 + (size_t)8 * LDR;
#pragma unroll 4
    for (int q = 0; q < ROWS / 4; ++q) {
      const double av = pa[4 * q], b0 = pb0[4 * q], b1 = pb1[4 * q];
      dmma884(c00, c01, av, b0);
      dmma884(c10, c11, av, b1);
    }
    Gl[(8 * I + g8) * PG_LDS + 8 * J0 + 2 * t4] = c00;
    Gl[(8 * I + g8) * PG_LDS + 8 * J0 + 2 * t4 + 1] = c01;
    Gl[(8 * I + g8) * PG_LDS + 8 * (J0 + 1) + 2 * t4] = c10;
    Gl[(8 * I + g8) * PG_LDS + 8 * (J0 + 1) + 2 * t4 + 1] = c11;
  };
  // cluster-wide sum of the local Gram matrices (fixed order): rank d reduces rows d RS .. and broadcasts them
  auto gram_reduce = [&]() {
    __syncthreads();
    if (CS == 1) {
      for (int idx = tid; idx < CB * CB; idx += PQ_NT) Gs[(idx >> 5) * PG_LDS + (idx & 31)] = Gl[(idx >> 5) * PG_LDS + (idx & 31)];
      __syncthreads();
      return;
    }
    for (int idx = tid; idx < CB * CB; idx += PQ_NT) {
      const int row = idx >> 5, col = idx & 31;
      const int dstr = row / RS;
      cluster.map_shared_rank(recv, dstr)[rank * (RS * CB) + (row - dstr * RS) * CB + col] = Gl[row * PG_LDS + col];
    }
    cluster.sync();
    for (int idx = tid; idx < RS * CB; idx += PQ_NT) {
      double sum = 0.0;
      for (int src = 0; src < CS; ++src) sum += recv[src * (RS * CB) + idx];
      const int row = rank * RS + idx / CB, col = idx % CB;
      for (int dsti = 0; dsti < CS; ++dsti) cluster.map_shared_rank(Gs, dsti)[row * PG_LDS + col] = sum;
    }
    cluster.sync();
  };

  // ---- phase 1: G = P'P, top block to everybody ----
#pragma unroll
  for (int i = 0; i < NRL; ++i)
#pragma unroll
    for (int k = 0; k < CB; ++k) Ys[(size_t)k * LDR + i * PQ_NT + tid] = pr[i][k];
  if (rank == 0 && tid < CB) {
    for (int dsti = 0; dsti < CS; ++dsti) {
      double* tp = cluster.map_shared_rank(tops, dsti);
#pragma unroll
      for (int k = 0; k < CB; ++k) tp[tid * PG_LDS + k] = pr[0][k];
    }
  }
  __syncthreads();
  PG_TICK(0)
  gram_local();
  PG_TICK(1)
  gram_reduce();   // its cluster barriers also publish the top block
  if (CS == 1) __syncthreads();
  PG_TICK(2)

  // ---- phase 2: the table of reflector scalars and update coefficients (warp 0, lane k <-> column k) ----
  if (wid == 0) {
    int bad = 0;
    double tc[CB];                                 // column `lane` of the top block, updated in registers
#pragma unroll
    for (int r = 0; r < CB; ++r) tc[r] = tops[r * PG_LDS + lane];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // sum_{r<c} R[r][c] R[r][k]
#pragma unroll
      for (int r = 0; r < c; ++r) {
        const double rc = Rs[r * PG_LDS + c], rk = Rs[r * PG_LDS + lane];
        if ((r & 3) == 0) s0 = fma(rc, rk, s0);
        else if ((r & 3) == 1) s1 = fma(rc, rk, s1);
        else if ((r & 3) == 2) s2 = fma(rc, rk, s2);
        else s3 = fma(rc, rk, s3);
      }
      const double ssum = (s0 + s1) + (s2 + s3);
      const double pk = tc[c];
      const double alpha = __shfl_sync(0xffffffffu, pk, c);
      const double gk = Gs[c * PG_LDS + lane] - ssum - alpha * pk;
      const double xn2 = __shfl_sync(0xffffffffu, gk, c);
      const double gcc = Gs[c * PG_LDS + c];
      double tau = 0.0, beta = alpha, scale = 0.0;
      if (!(xn2 >= PQ_THETA * gcc) && gcc > 0.0) bad = 1;     // cancellation (or a NaN): let the exchange kernel redo the panel
      if (xn2 > 0.0) {
        const double sq = fma(alpha, alpha, xn2);
        const double rn = rsqrt(sq);
        const double nrm = sq * rn, aa = fabs(alpha);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = fma(aa, rn, 1.0);
        const double rcp = 1.0 / (aa + nrm);
        scale = alpha >= 0.0 ? rcp : -rcp;
      }
      const double w = lane > c ? tau * (pk + scale * gk) : 0.0;
      wtab[c * CB + lane] = w;
      if (lane == c) {
        sctab[c] = scale;
        betab[c] = beta;
        tautab[c] = tau;
      }
      // finished row c of R; rows c+1.. of the top block get the reflector (v_r = scale * column c of the block)
      Rs[c * PG_LDS + lane] = lane > c ? pk - w : (lane == c ? beta : 0.0);
#pragma unroll
      for (int r = c + 1; r < CB; ++r) {
        const double vr = scale * __shfl_sync(0xffffffffu, tc[r], c);
        tc[r] = fma(-vr, w, tc[r]);
      }
      __syncwarp();
    }
    if (lane == 0) *flag = bad;
  }
  __syncthreads();
  PG_TICK(3)
  if (*flag) {                                    // uniform over the cluster: every CTA computed the same table
    if (rank == 0 && tid == 0) *fb = 1;
    cluster.sync();
    return;
  }

  // ---- phase 3: apply the 32 reflectors to my rows ----
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    const double scale = sctab[c], beta = betab[c];
#pragma unroll
    for (int i = 0; i < NRL; ++i) {
      const int grow = row_lo + i * PQ_NT + tid;
      if (grow >= c) {
        const double vr = grow > c ? scale * pr[i][c] : 1.0;
#pragma unroll
        for (int k = c + 1; k < CB; ++k) pr[i][k] = fma(-vr, wtab[c * CB + k], pr[i][k]);
        pr[i][c] = grow > c ? vr : beta;
      }
    }
  }

  PG_TICK(4)
  // ---- phase 4: T from the Gram matrix of the reflectors (dlarft) ----
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    const int grow = row_lo + i * PQ_NT + tid;
#pragma unroll
    for (int k = 0; k < CB; ++k) Ys[(size_t)k * LDR + i * PQ_NT + tid] = grow > k ? pr[i][k] : (grow == k ? 1.0 : 0.0);
  }
  __syncthreads();
  gram_local();
  gram_reduce();
  PG_TICK(5)
  if (rank == 0 && wid == 0) {
    double trow[CB];                              // lane i owns row i of T
#pragma unroll
    for (int c = 0; c < CB; ++c) trow[c] = 0.0;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double tau = tautab[c];
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int k = 0; k < c; ++k) {
        const double zz = Gs[k * PG_LDS + c];
        if ((k & 3) == 0) acc0 = fma(trow[k], zz, acc0);
        else if ((k & 3) == 1) acc1 = fma(trow[k], zz, acc1);
        else if ((k & 3) == 2) acc2 = fma(trow[k], zz, acc2);
        else acc3 = fma(trow[k], zz, acc3);
      }
      trow[c] = lane < c ? -tau * ((acc0 + acc1) + (acc2 + acc3)) : (lane == c ? tau : 0.0);
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) a.T[lane + c * CB] = trow[c];
    a.tau[a.j + lane] = tautab[lane];
  }
  PG_TICK(6)
  // ---- outputs ----
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    const int r = i * PQ_NT + tid;
    if (r < nr) {
      const int grow = row_lo + r;
#pragma unroll
      for (int k = 0; k < CB; ++k) {
        const double v = pr[i][k];
        a.Y[(r0 + grow) + (long long)(a.j + k) * a.ldy] = grow > k ? v : (grow == k ? 1.0 : 0.0);
        a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = grow <= k ? v : 0.0;
      }
    }
  }
  cluster.sync();
  PG_TICK(7)
  if (gprof && rank == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + 8 + i, (unsigned long long)gpc[i]);
}

// NRL > 0: rp = 256 NRL and the panel lives in REGISTERS (row i*256 + tid of the CTA's slice in pr[i][0..31]);
// NRL == 0: any rp (multiple of 256), the panel lives in shared memory.
template <int NRL>
__global__ void __launch_bounds__(PQ_NT, 1) k_panel_qr(const PanelArgs a) {
  PDL_ENTRY();
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool prof = a.prof != nullptr;
  long long tk = prof ? clock64() : 0;
#define PQ_TICK(slot)                 \
  if (prof) {                         \
    const long long _n = clock64();   \
    pc[slot] += _n - tk;              \
    tk = _n;                          \
  }
  constexpr bool REG = NRL > 0;
  constexpr int NR = REG ? NRL : 1;
  if (a.only_if && *a.only_if == 0) return;       // uniform over the whole cluster
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  extern __shared__ double sm[];
  const int rp = a.rp, ldp = a.ldp;
  double* ex = sm;                             // [2][PQ_MAXCS][32] partial dots
  double* piv = ex + 2 * PQ_MAXCS * CB;        // [2][32] pivot row (as exchanged)
  double* wpart = piv + 2 * CB;                // [2][PQ_NW][32] per-warp partial dots
  double* wsw = wpart + 2 * PQ_NW * CB;        // [PQ_NW][32] update coefficients, one copy per warp
  double* Zm = wsw + PQ_NW * CB;               // [32*32] Zm[k + c*32] = y_k' y_c (k < c)
  double* sT = Zm + CB * CB;                   // [32*32] T factor
  double* stau = sT + CB * CB;                 // [32]
  double* pivs = stau + CB;                    // [32] pivot row staged by its owner thread (register path)
  double* P = pivs + CB;                       // shared-memory path: 32 columns x rp rows, column-major (ld = ldp, odd)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r0 = a.j + CB, m = a.n - r0;
  const int kb = min(CB, m - 1);
  const int row_lo = rank * rp;
  const int nr = max(0, min(m, row_lo + rp) - row_lo);
  const int nrl = rp / PQ_NT;                  // rows per lane
  double pr[NR][CB];
  if (REG) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = i * PQ_NT + tid;
      const double* src = a.A + (r0 + row_lo + r) + (long long)a.j * a.lda;
#pragma unroll
      for (int k = 0; k < CB; ++k) pr[i][k] = r < nr ? src[(long long)k * a.lda] : 0.0;
    }
  } else {
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
    const int prank = c / rp, rc = c - prank * rp;     // owner CTA and local row of the pivot row
    // (A) products of my rows (strictly below the pivot row) with every column, reduced over the warp
    {
      double val[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) val[k] = 0.0;
      if (REG) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int r = i * PQ_NT + tid;
          double pcv = 0.0;
#pragma unroll
          for (int k = 0; k < CB; ++k) pcv = k == c ? pr[i][k] : pcv;
          if (!(row_lo + r > c && r < nr)) pcv = 0.0;
#pragma unroll
          for (int k = 0; k < CB; ++k) val[k] = fma(pcv, pr[i][k], val[k]);
          if (rank == prank && r == rc) {
#pragma unroll
            for (int k = 0; k < CB; ++k) pivs[k] = pr[i][k];
          }
        }
      } else {
        for (int i = 0; i < nrl; ++i) {
          const int r = i * PQ_NT + tid;
          if (row_lo + r > c && r < nr) {
            const double pcv = P[(size_t)c * ldp + r];
#pragma unroll
            for (int k = 0; k < CB; ++k) val[k] = fma(pcv, P[(size_t)k * ldp + r], val[k]);
          }
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
    // (B) CTA partial -> exchange buffers of every CTA (warp w serves the destinations w, w + 8); the owner of the
    //     pivot row adds that row
    {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < PQ_NW; ++w) tot += wpart[((size_t)par * PQ_NW + w) * CB + lane];
      double pv = 0.0;
      if (rank == prank) pv = REG ? pivs[lane] : P[(size_t)lane * ldp + rc];
      for (int dsti = wid; dsti < CS; dsti += PQ_NW) {
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
        const double rcp = 1.0 / (aa + nrm);
        scale = alpha >= 0.0 ? rcp : -rcp;
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
    // (D) apply to my rows (the coefficients of the columns <= c are zero)
    {
      const double* w = wsw + wid * CB;
      if (REG) {
        double wk[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) wk[k] = w[k];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int grow = row_lo + i * PQ_NT + tid;
          if (grow >= c) {
            double pcv = 0.0;
#pragma unroll
            for (int k = 0; k < CB; ++k) pcv = k == c ? pr[i][k] : pcv;
            const double vr = grow > c ? scale * pcv : 1.0;
#pragma unroll
            for (int k = 0; k < CB; ++k) pr[i][k] = k == c ? (grow > c ? vr : beta) : fma(-vr, wk[k], pr[i][k]);
          }
        }
      } else {
        for (int i = 0; i < nrl; ++i) {
          const int r = i * PQ_NT + tid;
          if (r >= nr) continue;
          const int grow = row_lo + r;
          if (grow >= c) {
            const double vr = grow > c ? scale * P[(size_t)c * ldp + r] : 1.0;
            for (int k0 = (c + 1) & ~7; k0 < CB; k0 += 8) {     // columns <= c inside the first chunk have w = 0
              double pk8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) pk8[u] = P[(size_t)(k0 + u) * ldp + r];
#pragma unroll
              for (int u = 0; u < 8; ++u) P[(size_t)(k0 + u) * ldp + r] = fma(-vr, w[k0 + u], pk8[u]);
            }
            P[(size_t)c * ldp + r] = grow > c ? vr : beta;
          }
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
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int k = 0; k < c; ++k) {                 // trow[k] = 0 for k < lane
        const double zz = Zm[k + c * CB];
        if ((k & 3) == 0) acc0 = fma(trow[k], zz, acc0);
        else if ((k & 3) == 1) acc1 = fma(trow[k], zz, acc1);
        else if ((k & 3) == 2) acc2 = fma(trow[k], zz, acc2);
        else acc3 = fma(trow[k], zz, acc3);
      }
      trow[c] = lane < c ? -tau * ((acc0 + acc1) + (acc2 + acc3)) : (lane == c ? tau : 0.0);
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) sT[lane + c * CB] = trow[c];
  }
  __syncthreads();
  // outputs: R into the band part of A, Y (unit lower trapezoidal, explicit zeros) into the reflector store
  if (REG) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = i * PQ_NT + tid;
      if (r < nr) {
        const int grow = row_lo + r;
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          const double v = pr[i][k];
          const double y = k >= kb ? 0.0 : (grow > k ? v : (grow == k ? 1.0 : 0.0));
          a.Y[(r0 + grow) + (long long)(a.j + k) * a.ldy] = y;
          a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = (k >= kb || grow <= k) ? v : 0.0;
        }
      }
    }
  } else {
    for (int k = 0; k < CB; ++k)
      for (int r = tid; r < nr; r += PQ_NT) {
        const int grow = row_lo + r;
        const double v = P[(size_t)k * ldp + r];
        const double y = k >= kb ? 0.0 : (grow > k ? v : (grow == k ? 1.0 : 0.0));
        a.Y[(r0 + grow) + (long long)(a.j + k) * a.ldy] = y;
        a.A[(r0 + grow) + (long long)(a.j + k) * a.lda] = (k >= kb || grow <= k) ? v : 0.0;
      }
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

// Z0 = A22 Y (A22 symmetric m x m, full storage, leading dimension even; Y m x 32) on the FP64 tensor cores (DMMA
// m8n8k4): CTA (rb, ks) accumulates the 64 x 32 tile of rows 64 rb .. over the k range of split ks (chunks of 64 through a
// two-stage cp.async ring) and writes it to Zpart[ks]; it also forms its share of G0 = Y' Z0 = Y' A22 Y.
// k_reduce_g sums the shares in a fixed order (deterministic).
constexpr int SY_BM = 64, SY_BK = 64, SY_LD = SY_BK + 4;
constexpr int SY_STAGE = SY_BK * (SY_BM + 4) + CB * SY_LD;     // doubles per stage: A tile [k][m] (ld 68) + Y tile [n][k] (ld 68)
__global__ void __launch_bounds__(128) k_symm_y(const double* __restrict__ A, long long lda, const double* __restrict__ Y, long long ldy, int m,
                                                int ksplit, double* __restrict__ Zpart, long long ldz, double* __restrict__ Gpart, int vec) {
  PDL_ENTRY();
  extern __shared__ double smy[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int rb = blockIdx.x, ks = blockIdx.y;
  const int row0 = rb * SY_BM;
  const int kchunk = ((m + ksplit - 1) / ksplit + SY_BK - 1) / SY_BK * SY_BK;
  const int k0 = ks * kchunk, k1 = min(m, k0 + kchunk);
  const int nit = k1 > k0 ? (k1 - k0 + SY_BK - 1) / SY_BK : 0;
  auto load_stage = [&](int it, int st) {
    double* As = smy + st * SY_STAGE;                 // As[kk * 68 + r]
    double* Ys = As + SY_BK * (SY_BM + 4);            // Ys[n * 68 + kk]
    const int kk0 = k0 + it * SY_BK;
    for (int idx = tid; idx < SY_BK * (SY_BM / 2); idx += 128) {      // 16-byte chunks: 32 per k column
      const int kk = idx >> 5, r = (idx & 31) * 2;
      const int gr = row0 + r, gk = kk0 + kk;
      double* dst = As + kk * (SY_BM + 4) + r;
      const double* src = A + gr + (long long)gk * lda;
      if (vec) {                                   // m, lda even: rows come in aligned pairs
        cpa16(dst, src, gr < m && gk < k1);
      } else {
        cpa8(dst, src, gr < m && gk < k1);
        cpa8(dst + 1, src + 1, gr + 1 < m && gk < k1);
      }
    }
    for (int idx = tid; idx < CB * (SY_BK / 2); idx += 128) {
      const int nn = idx >> 5, kk = (idx & 31) * 2;
      const int gk = kk0 + kk;
      double* dst = Ys + nn * SY_LD + kk;
      const double* src = Y + gk + (long long)nn * ldy;
      if (vec) {
        cpa16(dst, src, gk < k1);
      } else {
        cpa8(dst, src, gk < k1);
        cpa8(dst + 1, src + 1, gk + 1 < k1);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  if (nit > 0) load_stage(0, 0);
  for (int it = 0; it < nit; ++it) {
    if (it + 1 < nit) load_stage(it + 1, (it + 1) & 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const double* As = smy + (it & 1) * SY_STAGE;
    const double* Ys = As + SY_BK * (SY_BM + 4);
    const int kk0 = k0 + it * SY_BK;
    const int klim = min(SY_BK, k1 - kk0);            // rows of Y beyond k1 were zero-filled; columns of A beyond k1 too
    (void)klim;
#pragma unroll 4
    for (int kq = 0; kq < SY_BK / 4; ++kq) {
      const int k = kq * 4 + t;
      double af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = As[k * (SY_BM + 4) + warp * 16 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Ys[(j * 8 + g) * SY_LD + k];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // this thread holds Z[row = warp*16 + 8i + g][col = 8j + 2t + e]
  double (*Zs)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(smy);                       // 64 x 33
  double (*Yr)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(smy + SY_BM * (CB + 1));    // 64 x 33
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = warp * 16 + i * 8 + g, col = j * 8 + 2 * t + e;
        Zs[r][col] = acc[i][j][e];
      }
  for (int idx = tid; idx < SY_BM * CB; idx += 128) {
    const int r = idx & (SY_BM - 1), col = idx >> 6;
    Yr[r][col] = (row0 + r < m) ? Y[(row0 + r) + (long long)col * ldy] : 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < SY_BM * CB; idx += 128) {       // coalesced store of the tile
    const int r = idx & (SY_BM - 1), col = idx >> 6;
    if (row0 + r < m) Zpart[(long long)ks * ldz * CB + (row0 + r) + (long long)col * ldz] = Zs[r][col];
  }
  // G share: Gp[i][j] = sum_r Yr[r][i] Zs[r][j]; thread computes i = warp*8 .. +7, j = lane
  double* gp = Gpart + ((long long)ks * gridDim.x + rb) * (CB * CB);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = warp * 8 + q;
    double g0 = 0.0, g1 = 0.0;
#pragma unroll 8
    for (int r = 0; r < SY_BM; r += 2) {
      g0 = fma(Yr[r][i], Zs[r][lane], g0);
      g1 = fma(Yr[r + 1][i], Zs[r + 1][lane], g1);
    }
    gp[i + lane * CB] = g0 + g1;
  }
}

// C (m x m, column-major, ldc) -= L R' with L, R m x 64 (leading dimension ldp): the rank-64 two-sided update
// A22 -= [Y W][W Y]' of the band reduction, one 64 x 64 tile per CTA, both operand tiles resident in shared memory.
constexpr int RU_LD = 64 + 4;
__global__ void __launch_bounds__(128) k_rank64_update(double* __restrict__ C, long long ldc, const double* __restrict__ L,
                                                       const double* __restrict__ R, long long ldp, int m) {
  PDL_ENTRY();
  extern __shared__ double smu[];
  double* Ls = smu;                    // Ls[k * 68 + r]
  double* Rs = smu + 64 * RU_LD;       // Rs[k * 68 + c]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  for (int idx = tid; idx < 64 * 32; idx += 128) {
    const int kk = idx >> 5, r = (idx & 31) * 2;
    cpa16(Ls + kk * RU_LD + r, L + (m0 + r) + (long long)kk * ldp, m0 + r < m);
    cpa16(Rs + kk * RU_LD + r, R + (n0 + r) + (long long)kk * ldp, n0 + r < m);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  // prefetch the C tile entries this thread updates while the operands arrive
  double cv[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = m0 + wm + i * 8 + g, cidx = n0 + wn + j * 8 + 2 * t + e;
        cv[i][j][e] = (r < m && cidx < m) ? C[r + (long long)cidx * ldc] : 0.0;
      }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kq = 0; kq < 16; ++kq) {
    const int k = kq * 4 + t;
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = Ls[k * RU_LD + wm + i * 8 + g];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = Rs[k * RU_LD + wn + j * 8 + g];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = m0 + wm + i * 8 + g, cidx = n0 + wn + j * 8 + 2 * t + e;
        if (r < m && cidx < m) C[r + (long long)cidx * ldc] = cv[i][j][e] - acc[i][j][e];
      }
}

// G0[idx] = sum over the shares (fixed order: 8 interleaved partial sums per entry, combined by a shuffle tree)
__global__ void __launch_bounds__(256) k_reduce_g(const double* __restrict__ Gpart, int nshare, double* __restrict__ G0) {
  PDL_ENTRY();
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
  PDL_ENTRY();
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
void sb2st(tnad_ctx* c, const double* AB, int64_t ldab, int64_t n, double* dd, double* ee, double* V2, int64_t ldv, double* tau2,
           const std::function<void(int)>& overlap) {
  TNAD_REQUIRE(n >= 3 && ldab >= CB + 1, "sb2st: need n >= 3 and a band store with 33 rows");
  TNAD_REQUIRE(c->coop_launch, "sb2st: the chase kernel needs cooperative (co-resident) launches");
  const int NP = (int)chase_positions(n);
  int W = std::max(1, std::min(4, opt_i(c, "TNAD_CHASE_W", 4)));
  while ((NP + W - 1) / W > c->num_sms) ++W;      // all CTAs must be co-resident (one per SM)
  TNAD_REQUIRE(W <= 4, "sb2st: matrix too large for the chase kernel (n <= 32 * 4 * #SMs)");   // 3 W warps per CTA, 384 threads
  const int G = (NP + W - 1) / W;
  Tens gbox = t_alloc(c, {(int64_t)NP * 3 * MB_WORDS}, true);
  Tens err = t_alloc(c, {2}, true);
  ChaseArgs a;
  a.AB = AB; a.ldab = (int)ldab; a.n = (int)n; a.NP = NP; a.W = W;
  a.d = dd; a.e = ee; a.V2 = V2; a.ldv = ldv; a.tau2 = tau2;
  a.gbox = reinterpret_cast<unsigned long long*>(gbox.p);
  a.err = reinterpret_cast<int*>(err.p);
  const bool prof = opt_i(c, "TNAD_DC_DEBUG", 0) >= 2;
  Tens pbuf = t_alloc(c, {(int64_t)NP * 30 + 2}, true);
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
  cfg.blockDim = dim3(96 * W);
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
    KTimer kt(c, KF_CHASE);
    cudaError_t le = cudaLaunchKernelEx(&cfg, k_chase, a);
    if (le != cudaSuccess) {   // co-residency cannot be promised together with this cluster shape: the mailbox time-out still guards
      cudaGetLastError();
      cfg.numAttrs = 1;
      le = cudaLaunchKernelEx(&cfg, k_chase, a);
    }
    TNAD_CUDA(le);
  }
  c->launches++;
  // the chase owns Gp SMs for milliseconds and leaves the rest of the device idle: the caller may enqueue independent
  // work on another stream here, before the host waits for the pipeline's status word
  if (overlap) overlap(Gp);
  int herr = 0;
  TNAD_CUDA(cudaMemcpyAsync(&herr, a.err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  if (herr) fail(TNAD_ERR_INTERNAL, "sb2st: the chase pipeline timed out");
  if (prof) {
    std::vector<long long> ph((size_t)NP * 30);
    TNAD_CUDA(cudaMemcpy(ph.data(), pbuf.p, ph.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int t : {0, 1, NP / 2, NP / 2 + 1}) {
      if (t >= NP) continue;
      const double ns = (double)std::max<int64_t>(1, std::min<int64_t>(n - 2, n - 1 - CB * t));
      const long long* e = ph.data() + (size_t)t * 30;
      const long long* d = e + 10;
      const long long* e2 = e + 20;
      fprintf(stderr, "[tnad dc] chase position %d (W=%d) cycles/hop  E1: wait D+E2 %.0f  load %.0f  wait v %.0f  dot+sum %.0f  wait beta %.0f  scalars+sends %.0f  block to E2 %.0f  store %.0f | %.0f   D: load+row %.0f  wait v %.0f  first half %.0f  bulk %.0f | %.0f   E2: wait E1 %.0f  column sums+row %.0f  left-apply %.0f | %.0f\n",
              t, W, e[0] / ns, e[2] / ns, e[1] / ns, e[6] / ns, e[7] / ns, e[8] / ns, e[4] / ns, e[3] / ns, e[5] / ns, d[0] / ns, d[1] / ns, d[3] / ns, d[4] / ns, d[5] / ns,
              e2[0] / ns, e2[2] / ns, e2[3] / ns, e2[5] / ns);
    }
  }
}

// X[0:n, 0:ncols] <- Q2 X (reflectors of sb2st)
void apply_q2(tnad_ctx* c, const double* V2, int64_t ldv, const double* tau2, int64_t n, double* X, int64_t ldx, int64_t ncols,
              bool x_is_identity) {
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
    a.ident = x_is_identity ? 1 : 0;
    // the first slot of the stage is active for sweeps s <= n - 2 - 32 q0 only: later (= earlier in time) sweeps are skipped
    a.s_top = (int)std::min<int64_t>(s_top, ((n - 1 - (int64_t)CB * a.q0) + 31) / 32 * 32 - 1);
    KTimer kt(c, KF_Q2);
    k_q2_stage<<<grid, Q2_NT, 0, c->stream>>>(a);
    LAUNCH_CHECK(c);
  }
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
  const int64_t ldpp = (n + 1) & ~1LL;       // even leading dimensions: the tile loaders move 16-byte chunks
  Tens P1 = t_alloc(c, {ldpp, 2 * CB}), P2 = t_alloc(c, {ldpp, 2 * CB});
  static std::atomic<unsigned long long> attr_devs{0};
  if (!((attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL)) {
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_qr<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_gram<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_gram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    TNAD_CUDA(cudaFuncSetAttribute(k_panel_gram<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    TNAD_CUDA(cudaFuncSetAttribute(k_symm_y, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * SY_STAGE * sizeof(double))));
    TNAD_CUDA(cudaFuncSetAttribute(k_rank64_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 64 * RU_LD * sizeof(double))));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const size_t fixed = (size_t)(2 * PQ_MAXCS * CB + 2 * CB + 2 * PQ_NW * CB + PQ_NW * CB + 2 * CB * CB + 2 * CB) * sizeof(double);
  const bool use_regs = opt_i(c, "TNAD_PANEL_REGS", 1) != 0;
  const bool use_gram = opt_i(c, "TNAD_PANEL_GRAM", 1) != 0;
  const bool pdl = opt_i(c, "TNAD_SY2SB_PDL", 0) != 0 && opt_i(c, "TNAD_DC_DEBUG", 0) < 2;   // measured: no gain (6.28 vs 6.20 ms at n = 2048): the kernels themselves, not the gaps between them, are the 95 us per panel
  auto launch = [&](auto kern, dim3 grid, dim3 block, size_t smem, auto... args) {   // plain grid, optional dependent launch
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    la[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = la;
    lc.numAttrs = pdl ? 1 : 0;
    TNAD_CUDA(cudaLaunchKernelEx(&lc, kern, args...));
    c->launches++;
  };
  auto gram_smem = [](int nrl) {
    return (size_t)(CB * (PQ_NT * nrl + 4) + 4 * CB * PG_LDS + 3 * CB * CB + 3 * CB + 8) * sizeof(double);
  };
  Tens fbflag = t_alloc(c, {(n / CB + 2) / 2 + 2}, true);
  const int rp_max = (int)((((232448 - 1024) - fixed) / (CB * sizeof(double)) - 1) / PQ_NT * PQ_NT);
  const bool prof = opt_i(c, "TNAD_DC_DEBUG", 0) >= 2;
  double tph[4] = {0, 0, 0, 0};
  Tens pprof = t_alloc(c, {16}, true);
  cudaEvent_t pe[5];
  if (prof)
    for (auto& e : pe) e = get_event(c);
  const auto host_t0 = std::chrono::steady_clock::now();
  for (int64_t j = 0; j + CB < n - 1; j += CB) {
    const int64_t r0 = j + CB, m = n - r0;
    int CS = (int)std::min<int64_t>(8, std::max<int64_t>(1, (m + PQ_NT - 1) / PQ_NT));
    int rp = (int)(((m + CS - 1) / CS + PQ_NT - 1) / PQ_NT * PQ_NT);
    if (rp > 2 * PQ_NT) {   // more than two rows per lane: a cluster of 16 keeps the panel in registers up to m = 8192
      CS = PQ_MAXCS;
      rp = (int)(((m + CS - 1) / CS + PQ_NT - 1) / PQ_NT * PQ_NT);
    }
    TNAD_REQUIRE(rp <= rp_max, "sy2sb: panel too tall for one thread-block cluster (n <= 12000)");
    PanelArgs pa;
    pa.A = A; pa.lda = lda; pa.n = (int)n; pa.j = (int)j; pa.rp = rp; pa.ldp = rp + 1;
    pa.Y = Yst; pa.ldy = ldy; pa.tau = tau1; pa.T = Tb.p;
    pa.prof = prof ? reinterpret_cast<long long*>(pprof.p) : nullptr;
    pa.only_if = nullptr;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    const int nat = pdl ? 2 : 1;
    if (prof) TNAD_CUDA(cudaEventRecord(pe[0], st));
    // fast path: Gram-driven factorisation (no exchange per column); it raises fbflag[panel] when a column cancels
    int* fbp = reinterpret_cast<int*>(fbflag.p) + (j / CB);
    bool gram = use_gram && m >= CB + 1;
    if (gram) {
      int gnrl = 1, gcs = 1;
      while (gcs * PQ_NT < m && gcs < PQ_MAXCS) gcs *= 2;
      if ((int64_t)gcs * PQ_NT < m) {
        gnrl = 2;
        gcs = 1;
        while (gcs * 2 * PQ_NT < m && gcs < PQ_MAXCS) gcs *= 2;
      }
      if ((int64_t)gcs * gnrl * PQ_NT < m) gram = false;
      if (gram) {
        PanelArgs pg = pa;
        pg.rp = gnrl * PQ_NT;
        pg.prof = pa.prof;
        cudaLaunchConfig_t cg_ = {};
        cg_.gridDim = dim3(gcs);
        cg_.blockDim = dim3(PQ_NT);
        cg_.dynamicSmemBytes = gram_smem(gnrl);
        cg_.stream = st;
        at[0].val.clusterDim.x = gcs;
        cg_.attrs = at;
        cg_.numAttrs = nat;
        KTimer kt(c, KF_PANEL);
        if (gnrl == 1) TNAD_CUDA(cudaLaunchKernelEx(&cg_, k_panel_gram<1>, pg, fbp));
        else TNAD_CUDA(cudaLaunchKernelEx(&cg_, k_panel_gram<2>, pg, fbp));
        c->launches++;
        pa.only_if = fbp;
      }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(PQ_NT);
    const int nrl = rp / PQ_NT;
    const bool regs = use_regs && nrl <= 2;
    cfg.dynamicSmemBytes = fixed + (regs ? 0 : (size_t)CB * (rp + 1) * sizeof(double));
    cfg.stream = st;
    at[0].val.clusterDim.x = CS;
    cfg.attrs = at;
    cfg.numAttrs = nat;
    {
      KTimer kt(c, KF_PANEL);
      if (regs && nrl == 1) TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr<1>, pa));
      else if (regs) TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr<2>, pa));
      else TNAD_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr<0>, pa));
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
      KTimer kt(c, KF_SYMM);
      launch(k_symm_y, dim3(nrb, ksplit), dim3(128), 2 * SY_STAGE * sizeof(double), (const double*)A22p, (long long)lda, Yp, (long long)ldy,
             (int)m, ksplit, Zp.p, (long long)n, Gp.p, (n % 2 == 0 && lda % 2 == 0 && ldy % 2 == 0) ? 1 : 0);
      launch(k_reduce_g, dim3(CB * CB * 8 / 256), dim3(256), 0, (const double*)Gp.p, nrb * ksplit, Mb.p);
    }
    if (prof) TNAD_CUDA(cudaEventRecord(pe[2], st));
    launch(k_make_w, dim3((unsigned)((m + 31) / 32)), dim3(256), 0, (const double*)Zp.p, (long long)n, ksplit, Yp, (long long)ldy,
           (const double*)Tb.p, (const double*)Mb.p, (int)m, P1.p, P2.p, (long long)ldpp);
    if (prof) TNAD_CUDA(cudaEventRecord(pe[3], st));
    {
      KTimer kt(c, KF_RANK64);
      const int nt = (int)((m + 63) / 64);
      launch(k_rank64_update, dim3(nt, nt), dim3(128), 2 * 64 * RU_LD * sizeof(double), A22p, (long long)lda, (const double*)P1.p,
             (const double*)P2.p, (long long)ldpp, (int)m);
    }
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
  if (opt_i(c, "TNAD_DC_DEBUG", 0) == 1)
    fprintf(stderr, "[tnad dc] sy2sb host enqueue time %.2f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
  if (prof) {
    fprintf(stderr, "[tnad dc] sy2sb n=%lld (synchronised per panel): panel QR %.2f  Z0=A22*Y + G0 %.2f  make_w %.2f  update %.2f ms\n",
            (long long)n, tph[0], tph[1], tph[2], tph[3]);
    for (auto& e : pe) c->event_pool.push_back(e);
    long long ph[16];
    TNAD_CUDA(cudaMemcpy(ph, pprof.p, sizeof(ph), cudaMemcpyDeviceToHost));
    {
      const double np_ = std::ceil((double)std::max<int64_t>(1, n - CB - 1) / CB);
      fprintf(stderr, "[tnad dc] k_panel_gram (rank 0, thread 0) cycles/panel: load+stage %.0f  gram %.0f  reduce %.0f  table %.0f  apply %.0f  gram(Y)+reduce %.0f  T %.0f  store %.0f\n",
              ph[8] / np_, ph[9] / np_, ph[10] / np_, ph[11] / np_, ph[12] / np_, ph[13] / np_, ph[14] / np_, ph[15] / np_);
    }
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
