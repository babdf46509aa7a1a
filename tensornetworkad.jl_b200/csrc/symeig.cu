// Symmetric eigensolver for the CTMRG projector: two-sided block Jacobi on M = cpmat + cpmat'.
//
// ctmrg.jl:135-136 takes `svd(cpmat)` of a *symmetric* matrix, so U S V' = Q |L| (Q sign(L))' with the
// eigen-decomposition M = Q L Q'.  Working on M itself (instead of one-sided Jacobi on M^2's Gram
// matrices) halves the number of sweeps: the spectrum is not squared and convergence is judged on
// absolute off-diagonal size, which is the accuracy class of LAPACK's dgesdd that the reference uses.
// Q is a product of rotations, so it is orthogonal to rounding whatever the spectrum (no null-space
// completion is needed).
//
// One round (all N/64 disjoint block pairs of a round-robin sweep in parallel):
//   1. k_sym_eig        the 64x64 pivot block M[(I,J),(I,J)] is read straight from M (no Gram pass) and
//                       diagonalised by a cyclic two-sided Jacobi in shared memory -> W, columns sorted
//                       by decreasing |eigenvalue|.
//   2. k_sym_update_m   M <- W' M W, fused: for every pair-of-pairs (a <= b) the 64x64 block
//                       M[a,b] <- W_a' M[a,b] W_b (two DMMA products in shared memory) and its mirror
//                       M[b,a] is written transposed, so only half of M is computed and M stays
//                       exactly symmetric.
//   3. jacobi_rotate_columns   Q[:, (I,J)] <- Q[:, (I,J)] W   (the DMMA panel kernel of jacobi.cu).
#include "common.h"
#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace tnad {

void jacobi_rotate_columns(tnad_ctx* c, double* X, int64_t ld, int nchunks, int p, int round, const double* Wbuf,
                           const int* skip, cudaStream_t st);   // jacobi.cu

namespace {

constexpr int JB = 32;
constexpr int JP = 64;
constexpr int HLD = JP + 1;
constexpr int BLD = JP + 4;

#define LAUNCH_CHECK(c)            \
  do {                             \
    (c)->launches++;               \
    TNAD_CUDA(cudaGetLastError()); \
  } while (0)

__host__ __device__ inline void rr_pair(int p, int r, int k, int& a, int& b) {
  if (k == 0) {
    a = p - 1;
    b = r;
  } else {
    a = (r + k) % (p - 1);
    b = (r - k + (p - 1)) % (p - 1);
  }
  if (a > b) {
    int t = a;
    a = b;
    b = t;
  }
}

__device__ __forceinline__ void cp_async16(double* s, const double* g) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ long long pair_index(int I, int J, int r) {   // r in [0,64) -> global row/col
  return r < JB ? (long long)I * JB + r : (long long)J * JB + r - JB;
}

// ---- 1. pivot blocks ------------------------------------------------------------------------------------
// Implicit two-sided Jacobi on the 64x64 pivot P, register resident.  The iteration is the classical
// cyclic Jacobi eigenvalue method (H <- J'HJ), but H is never stored: the kernel keeps W (the accumulated
// rotations) and Y = P W, so a rotation only touches columns, and the three pivot entries it needs are
// the dots H_pq = w_p . y_q, recomputed from scratch every step (no error accumulates in H).
// Layout: lane k owns the column slot pair (top_k, bot_k); warp w owns rows 8w..8w+7 of Y and W in
// registers.  The 32 rotations of a step act on the 32 slot pairs; between steps the columns move to the
// next Brent-Luk position with two warp shuffles per value, so nothing but the 8x32x3 partial dots
// crosses shared memory.
constexpr int EW = 8;            // warps
constexpr int ER = JP / EW;      // rows per warp

// GRAM == false: P is the pivot block of the symmetric matrix M (absolute threshold tolfac * |M|_F).
// GRAM == true : P is the Gram matrix X'X of a column pair of the one-sided method (jacobi.cu), summed from
//                `nsplit` partial products; thresholds are relative (cosines) and numerically-null columns
//                (squared norm <= nullfac * |A|_F^2) are frozen.
template <bool GRAM>
__global__ void __launch_bounds__(EW * 32) k_pivot_eig(const double* __restrict__ M, long long ld, int p, int round,
                                                       int nreal, const double* __restrict__ fro2, double tolfac,
                                                       int max_inner, int cross, double* __restrict__ Wbuf,
                                                       int* __restrict__ skip,
                                                       unsigned long long* __restrict__ offmax_bits, int nsplit,
                                                       double nullfac) {
  __shared__ double Ps[JP * HLD];
  __shared__ double part[2][EW][32][3];
  __shared__ double red[EW];
  __shared__ double keys[JP];
  const int pair = blockIdx.x, tid = threadIdx.x, w = tid >> 5, k = tid & 31;
  int I, J;
  rr_pair(p, round, pair, I, J);
  const double tol = GRAM ? tolfac : tolfac * sqrt(*fro2);   // relative (cosine) / absolute threshold
  const double thr2 = GRAM ? nullfac * (*fro2) : 0.0;
  if (GRAM) {
    const double* hp = M + (long long)pair * nsplit * (JP * JP);
    for (int idx = tid; idx < JP * JP; idx += EW * 32) {
      double h = 0.0;
      for (int sidx = 0; sidx < nsplit; ++sidx) h += hp[(long long)sidx * (JP * JP) + idx];
      Ps[(idx % JP) * HLD + idx / JP] = h;
    }
  } else {
    for (int idx = tid; idx < JP * JP; idx += EW * 32) {
      const int r = idx % JP, c = idx / JP;
      Ps[r * HLD + c] = M[pair_index(I, J, r) + pair_index(I, J, c) * ld];
    }
  }
  __syncthreads();
  double off = 0.0;
  for (int idx = tid; idx < JP * JP; idx += EW * 32) {
    const int r = idx % JP, c = idx / JP;
    if (r < c) {
      if (GRAM) {
        const double hr = Ps[r * HLD + r], hc = Ps[c * HLD + c];
        if (hr > thr2 && hc > thr2) off = fmax(off, fabs(Ps[r * HLD + c]) * rsqrt(hr * hc));
      } else {
        off = fmax(off, fabs(Ps[r * HLD + c]));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o));
  if (k == 0) red[w] = off;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < EW; ++i) m = fmax(m, red[i]);
    red[0] = m;
    atomicMax(offmax_bits, (unsigned long long)__double_as_longlong(m));
    skip[pair] = (m <= tol) ? 1 : 0;
  }
  __syncthreads();
  if (red[0] <= tol) return;

  double yt[ER], yb[ER], wt[ER], wb[ER];
  int ot = k, ob = JB + k;   // original column held in the top / bottom slot
#pragma unroll
  for (int i = 0; i < ER; ++i) {
    const int r = ER * w + i;
    yt[i] = Ps[r * HLD + k];
    yb[i] = Ps[r * HLD + JB + k];
    wt[i] = (r == k) ? 1.0 : 0.0;
    wb[i] = (r == JB + k) ? 1.0 : 0.0;
  }
  const double tol2 = tol * tol;
  int buf = 0;
  for (int sw = 0; sw < max_inner; ++sw) {
    int rotated = 0;
    // cross mode: only the 32 x 32 pairs (column of block I, column of block J) are rotated -- tops stay, bottoms
    // rotate through the lanes (one shuffle per value, 32 steps).  The within-block pairs are rotated once per
    // sweep by the full schedule (63 Brent-Luk steps) of the sweep's first round: every element pair of M is
    // then visited exactly once per sweep, as in the classical cyclic-by-blocks Jacobi method.
    const int nsteps = cross ? JB : JP - 1;
    for (int step = 0; step < nsteps; ++step) {
      double app = 0.0, aqq = 0.0, apq = 0.0;
#pragma unroll
      for (int i = 0; i < ER; ++i) {
        app += wt[i] * yt[i];
        aqq += wb[i] * yb[i];
        apq += 0.5 * (wt[i] * yb[i] + wb[i] * yt[i]);
      }
      part[buf][w][k][0] = app;
      part[buf][w][k][1] = aqq;
      part[buf][w][k][2] = apq;
      __syncthreads();
      app = aqq = apq = 0.0;
#pragma unroll
      for (int v = 0; v < EW; ++v) {   // same order in every warp -> bitwise identical rotation
        app += part[buf][v][k][0];
        aqq += part[buf][v][k][1];
        apq += part[buf][v][k][2];
      }
      buf ^= 1;
      if (GRAM ? (app > thr2 && aqq > thr2 && apq * apq > tol2 * app * aqq) : (apq * apq > tol2)) {
        const double tau = aqq - app;
        const double ww = tau * tau + 4.0 * apq * apq;
        const double d = fabs(tau) + ww * rsqrt(ww);
        const double rd = rsqrt(d);
        const double tt = copysign(2.0 * apq * rd * rd, tau * apq);
        const double c = rsqrt(1.0 + tt * tt);
        const double s = c * tt;
        rotated = 1;
#pragma unroll
        for (int i = 0; i < ER; ++i) {
          const double a = yt[i], b = yb[i];
          yt[i] = c * a - s * b;
          yb[i] = s * a + c * b;
          const double e = wt[i], f = wb[i];
          wt[i] = c * e - s * f;
          wb[i] = s * e + c * f;
        }
      }
      if (cross) {
#pragma unroll
        for (int i = 0; i < ER; ++i) {
          yb[i] = __shfl_sync(0xffffffffu, yb[i], (k + 1) & 31);
          wb[i] = __shfl_sync(0xffffffffu, wb[i], (k + 1) & 31);
        }
        ob = __shfl_sync(0xffffffffu, ob, (k + 1) & 31);
        continue;
      }
      // Brent-Luk move: top_0 stays, top_1 <- bot_0, top_k <- top_{k-1}, bot_k <- bot_{k+1}, bot_31 <- top_31
#pragma unroll
      for (int i = 0; i < ER; ++i) {
        {
          const double up = __shfl_up_sync(0xffffffffu, k == 0 ? yb[i] : yt[i], 1);
          const double dn = __shfl_down_sync(0xffffffffu, yb[i], 1);
          const double nt = (k == 0) ? yt[i] : up;
          yb[i] = (k == 31) ? yt[i] : dn;
          yt[i] = nt;
        }
        {
          const double up = __shfl_up_sync(0xffffffffu, k == 0 ? wb[i] : wt[i], 1);
          const double dn = __shfl_down_sync(0xffffffffu, wb[i], 1);
          const double nt = (k == 0) ? wt[i] : up;
          wb[i] = (k == 31) ? wt[i] : dn;
          wt[i] = nt;
        }
      }
      {
        const int up = __shfl_up_sync(0xffffffffu, k == 0 ? ob : ot, 1);
        const int dn = __shfl_down_sync(0xffffffffu, ob, 1);
        const int nt = (k == 0) ? ot : up;
        ob = (k == 31) ? ot : dn;
        ot = nt;
      }
    }
    if (!__syncthreads_or(rotated)) break;
  }
  // eigenvalue of each slot column (lambda = w . y) and the sort rank by decreasing |lambda|
  {
    double lt = 0.0, lb = 0.0;
#pragma unroll
    for (int i = 0; i < ER; ++i) {
      lt += wt[i] * yt[i];
      lb += wb[i] * yb[i];
    }
    part[buf][w][k][0] = lt;
    part[buf][w][k][1] = lb;
    __syncthreads();
    lt = lb = 0.0;
#pragma unroll
    for (int v = 0; v < EW; ++v) {
      lt += part[buf][v][k][0];
      lb += part[buf][v][k][1];
    }
    if (w == 0) {   // key: |lambda|; padding columns (exactly zero rows and columns of M) stay last
      keys[k] = pair_index(I, J, ot) >= nreal ? -1.0 : (GRAM ? fmax(lt, 0.0) : fabs(lt));
      keys[JB + k] = pair_index(I, J, ob) >= nreal ? -1.0 : (GRAM ? fmax(lb, 0.0) : fabs(lb));
    }
    __syncthreads();
  }
  int rt = 0, rb = 0;
  {
    const double kt = keys[k], kb = keys[JB + k];
    for (int i = 0; i < JP; ++i) {
      const double ki = keys[i];
      rt += (ki > kt || (ki == kt && i < k)) ? 1 : 0;
      rb += (ki > kb || (ki == kb && i < JB + k)) ? 1 : 0;
    }
  }
  double* wout = Wbuf + (long long)pair * (JP * JP);
#pragma unroll
  for (int i = 0; i < ER; ++i) {
    const int r = ER * w + i;
    wout[r + JP * rt] = wt[i];
    wout[r + JP * rb] = wb[i];
  }
}

// ---- 2. fused two-sided update of M ------------------------------------------------------------------------
// mode 0: every block (a <= b); mode 1: only the blocks in `plist` (those that contain next round's pivot
// blocks, so the next k_sym_eig can start early); mode 2: every block not flagged in `prio`.
__global__ void __launch_bounds__(256) k_sym_update_m(double* __restrict__ M, long long ld, int p, int round,
                                                      const double* __restrict__ Wbuf, const int* __restrict__ skip,
                                                      int mode, const int* __restrict__ plist,
                                                      const unsigned char* __restrict__ prio, int npairs,
                                                      unsigned long long* work_counter) {
  int a, b;
  if (mode == 1) {
    a = plist[2 * blockIdx.x];
    b = plist[2 * blockIdx.x + 1];
  } else {
    a = blockIdx.x;
    b = blockIdx.y;
    if (b < a) return;
    if (mode == 2 && prio[a * npairs + b]) return;
  }
  const int ska = skip[a], skb = skip[b];
  if (ska && skb) return;
  if (work_counter && threadIdx.x == 0) atomicAdd(work_counter, 1ULL);   // executed blocks (roofline accounting)
  extern __shared__ __align__(16) double sm[];
  double* Xs = sm;                 // B, then T = Wa' B, then B' = T Wb; stored [col * BLD + row]
  double* Ws = Xs + JP * BLD;      // W_a, then W_b (two buffers = 70 KB: three CTAs per SM); Ws[i * BLD + k] = W[k][i]
  const int tid = threadIdx.x;
  int Ia, Ja, Ib, Jb;
  rr_pair(p, round, a, Ia, Ja);
  rr_pair(p, round, b, Ib, Jb);
  // loads: block rows = pair a, block cols = pair b; each column is two 32-row segments
#pragma unroll
  for (int it = 0; it < (JP * JP / 2) / 256; ++it) {
    const int q = tid + it * 256;
    const int col = q / (JP / 2), r2 = q % (JP / 2);
    const int r = 2 * r2;
    cp_async16(Xs + col * BLD + r, M + pair_index(Ia, Ja, r) + pair_index(Ib, Jb, col) * ld);
    if (!ska) cp_async16(Ws + col * BLD + r, Wbuf + (long long)a * (JP * JP) + col * JP + r);
  }
  if (ska)   // an untouched pair carries the identity (its Wbuf slot is stale)
    for (int idx = tid; idx < JP * JP; idx += 256) Ws[(idx / JP) * BLD + idx % JP] = (idx / JP == idx % JP) ? 1.0 : 0.0;
  cp_async_wait_all();
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wy = warp & 3, wx = warp >> 2;   // rows 16*wy, cols 32*wx
  double acc[2][4][2];
  // T = Wa' * B :  A[m=i][k] = W_a[k][i] = Ws[i*BLD + k],  B[k][n=c] = Xs[c*BLD + k]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < JP / 4; ++kk) {
    const int k = kk * 4 + t;
    double af[2], bf[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) af[i] = Ws[(16 * wy + 8 * i + g) * BLD + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = Xs[(32 * wx + 8 * j + g) * BLD + k];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
  __syncthreads();   // everyone is done reading B and W_a
  if (!skb) {
#pragma unroll
    for (int it = 0; it < (JP * JP / 2) / 256; ++it) {
      const int q = tid + it * 256;
      const int col = q / (JP / 2), r = 2 * (q % (JP / 2));
      cp_async16(Ws + col * BLD + r, Wbuf + (long long)b * (JP * JP) + col * JP + r);
    }
  } else {
    for (int idx = tid; idx < JP * JP; idx += 256) Ws[(idx / JP) * BLD + idx % JP] = (idx / JP == idx % JP) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) Xs[(32 * wx + 8 * j + 2 * t + e) * BLD + 16 * wy + 8 * i + g] = acc[i][j][e];
  cp_async_wait_all();
  __syncthreads();
  // B' = T * Wb :  A[m=i][k=c] = T[i][c] = Xs[c*BLD + i],  B[k=c][n=j] = W_b[c][j] = Ws[j*BLD + c]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < JP / 4; ++kk) {
    const int k = kk * 4 + t;
    double af[2], bf[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) af[i] = Xs[k * BLD + 16 * wy + 8 * i + g];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = Ws[(32 * wx + 8 * j + g) * BLD + k];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
  __syncthreads();   // everyone is done reading T
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) Xs[(32 * wx + 8 * j + 2 * t + e) * BLD + 16 * wy + 8 * i + g] = acc[i][j][e];
  __syncthreads();
  // write B' to M[a,b] and its transpose to M[b,a] (coalesced along the stored column in both cases)
  for (int idx = tid; idx < JP * JP; idx += 256) {
    const int r = idx % JP, cidx = idx / JP;
    M[pair_index(Ia, Ja, r) + pair_index(Ib, Jb, cidx) * ld] = Xs[cidx * BLD + r];
  }
  if (a != b) {
    for (int idx = tid; idx < JP * JP; idx += 256) {
      const int r = idx % JP, cidx = idx / JP;     // r: row within pair b, cidx: column within pair a
      M[pair_index(Ib, Jb, r) + pair_index(Ia, Ja, cidx) * ld] = Xs[r * BLD + cidx];
    }
  } else {
    // keep the pivot block exactly symmetric
    __syncthreads();
    for (int idx = tid; idx < JP * JP; idx += 256) {
      const int r = idx % JP, cidx = idx / JP;
      if (r > cidx) M[pair_index(Ia, Ja, r) + pair_index(Ia, Ja, cidx) * ld] = Xs[r * BLD + cidx];
    }
  }
}

// G[i,j] = A(i,j) + A(j,i) (if sym) into the zero-padded work matrix
__global__ void k_load_sym(double* G, long long ldg, const double* __restrict__ A, long long s0, long long s1,
                           long long n, int sym) {
  const long long total = n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % n, j = idx / n;
    double v = A[i * s0 + j * s1];
    if (sym) v += A[j * s0 + i * s1];
    else v = 0.5 * (v + A[j * s0 + i * s1]);
    G[i + j * ldg] = v;
  }
}

__global__ void k_diag(const double* __restrict__ M, long long ld, long long n, double* __restrict__ d) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = M[i + i * ld];
}

// U[:, r] = Q[:, perm[r]] ; V[:, r] = sign(lambda) Q[:, perm[r]]
__global__ void k_sym_finalize(const double* __restrict__ Q, long long ldq, const int* __restrict__ perm,
                               const double* __restrict__ sgn, long long n, double* __restrict__ U,
                               double* __restrict__ V) {
  const long long r = blockIdx.x;
  const long long j = perm[r];
  const double s = sgn[r];
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double q = Q[i + j * ldq];
    U[i + r * n] = q;
    V[i + r * n] = s * q;
  }
}

__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

// E = 1.5 I - 0.5 T in place (Newton-Schulz polish of a warm-start basis)
__global__ void k_ns_factor(double* T, long long n) {
  const long long total = n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % n, j = idx / n;
    T[idx] = (i == j ? 1.5 : 0.0) - 0.5 * T[idx];
  }
}


}  // namespace

// Gram-matrix pivot problem of the one-sided method (jacobi.cu): same register-resident kernel.
void launch_gram_pivot_eig(tnad_ctx* c, const double* Hpart, int nsplit, int npairs, int p, int round, int nreal,
                           const double* fro2, double tol_rel, double nullfac, int max_inner, int cross, double* Wbuf,
                           int* skip, unsigned long long* offbits, cudaStream_t st) {
  k_pivot_eig<true><<<npairs, EW * 32, 0, st ? st : c->stream>>>(Hpart, 0, p, round, nreal, fro2, tol_rel, max_inner, cross,
                                                                Wbuf, skip, offbits, nsplit, nullfac);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
}

}  // namespace tnad

// Per-context cache: work buffers, look-ahead tables and the captured one-sweep CUDA graph, keyed by n.
struct SymEigCache {
  int64_t n = -1;
  int max_inner = 0;
  tnad::Tens Mw, Q, Wbuf, Wbuf2, skipbuf, skipbuf2, tabbuf, plbuf;
  std::vector<int> pcount;
  size_t plist_stride = 0;
  cudaGraphExec_t exec = nullptr;
  int nodes = 0;
};

namespace tnad {

void symeig_cache_free(tnad_ctx* c) {
  if (!c->symcache) return;
  if (c->symcache->exec) cudaGraphExecDestroy(c->symcache->exec);
  delete c->symcache;
  c->symcache = nullptr;
}

// SVD of the symmetric matrix A (+ A^T when `sym_add_transpose`) through its eigen-decomposition.
SvdResult svd_symmetric(tnad_ctx* c, const Tens& A, bool sym_add_transpose, const Tens* Q0) {
  TNAD_REQUIRE(A.rank == 2 && A.dim[0] == A.dim[1], "svd_symmetric: need a square matrix");
  const int64_t n = A.dim[0];
  const int64_t N = (n + JP - 1) / JP * JP;
  const int p = (int)(N / JB), npairs = p / 2, nr = p - 1;
  const int chunks = (int)((N + 127) / 128);
  const int64_t ld = (int64_t)chunks * 128;   // rows padded to the 128-row chunks of the panel kernel
  const size_t smem_upd = (size_t)(2 * JP * BLD) * sizeof(double);
  static std::atomic<unsigned long long> attr_devs{0};   // kernel attributes are per device: one bit per device id (set after the attribute call: a racing thread at worst repeats it)
  const bool attr_set = (attr_devs.load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL;
  if (!attr_set) {
    TNAD_CUDA(cudaFuncSetAttribute(k_sym_update_m, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    attr_devs.fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const double eps = 2.220446049250313e-16;
  const double tolfac = 16.0 * eps;   // |m_pq| <= 16 eps |M|_F  (LAPACK-class absolute accuracy)
  const bool debug = opt_i(c, "TNAD_JACOBI_DEBUG", 0) != 0;
  const int max_inner = opt_i(c, "TNAD_SYMEIG_INNER", 1);
  const int max_sweeps = opt_i(c, "TNAD_JACOBI_SWEEPS", 60);
  const bool lookahead = opt_i(c, "TNAD_LOOKAHEAD", 1) != 0 && npairs >= 4;
  const bool use_graph = opt_i(c, "TNAD_GRAPH", 1) != 0 && !c->ktiming;
  const bool cross_mode = opt_i(c, "TNAD_SYMEIG_CROSS", 1) != 0;
  double* fro2 = c->scal + 18;
  unsigned long long* offbits = reinterpret_cast<unsigned long long*>(c->scal + 16);
  cudaStream_t S1 = c->stream, S2 = c->stream2, S3 = opt_i(c, "TNAD_S3", 0) ? c->stream3 : c->stream2;

  // ---- cached workspace ------------------------------------------------------------------------------------
  if (!c->symcache || c->symcache->n != n || c->symcache->max_inner != max_inner + (cross_mode ? 100 : 0)) {
    symeig_cache_free(c);
    SymEigCache* sc = new SymEigCache();
    c->symcache = sc;
    sc->n = n;
    sc->max_inner = max_inner + (cross_mode ? 100 : 0);
    sc->Mw = t_alloc(c, {ld, N});
    sc->Q = t_alloc(c, {ld, N});
    sc->Wbuf = t_alloc(c, {(int64_t)JP * JP, (int64_t)npairs});
    sc->Wbuf2 = t_alloc(c, {(int64_t)JP * JP, (int64_t)npairs});
    sc->skipbuf = t_alloc(c, {(int64_t)npairs + 2});
    sc->skipbuf2 = t_alloc(c, {(int64_t)npairs + 2});
    // look-ahead tables: the pivot blocks of round r+1 live in a few blocks of the round-r update ("priority"
    // blocks: the npairs diagonal blocks plus one cross block per next pair).
    std::vector<unsigned char> prio_h((size_t)nr * npairs * npairs + 8, 0);
    // measured at n = 2048 (528 blocks): 64 -> 750 ms, 296 -> 760, 420 -> 681, 480 -> 738, 528 -> 857 per 11 SVDs
    const int nblocks = npairs * (npairs + 1) / 2;
    const int fill_default = nblocks <= 4 * c->num_sms ? (4 * nblocks) / 5 : 2 * c->num_sms;
    const int prio_fill = std::min(opt_i(c, "TNAD_PRIO_FILL", fill_default), nblocks);
    const size_t plist_stride = (size_t)2 * std::max(2 * npairs, prio_fill);
    sc->plist_stride = plist_stride;
    std::vector<int> plist_h((size_t)nr * plist_stride + 8, 0);
    sc->pcount.assign((size_t)nr, 0);
    std::vector<int> pair_of((size_t)p);
    for (int r = 0; r < nr; ++r) {
      for (int k = 0; k < npairs; ++k) {
        int I, J;
        rr_pair(p, r, k, I, J);
        pair_of[I] = k;
        pair_of[J] = k;
      }
      unsigned char* tab = prio_h.data() + (size_t)r * npairs * npairs;
      for (int k = 0; k < npairs; ++k) tab[k * npairs + k] = 1;
      const int rn = (r + 1) % nr;
      for (int k = 0; k < npairs; ++k) {
        int I, J;
        rr_pair(p, rn, k, I, J);
        const int a = std::min(pair_of[I], pair_of[J]), b = std::max(pair_of[I], pair_of[J]);
        tab[a * npairs + b] = 1;
      }
      int cnt = 0;
      int* pl = plist_h.data() + (size_t)r * plist_stride;
      for (int a = 0; a < npairs; ++a)
        for (int b = a; b < npairs; ++b)
          if (tab[a * npairs + b]) {
            pl[2 * cnt] = a;
            pl[2 * cnt + 1] = b;
            ++cnt;
          }
      // The priority launch is latency bound (one short wave); top it up with ordinary blocks until it fills
      // one wave of the machine, which shortens the bulk launch on stream 2 for free.
      for (int a = 0; a < npairs && cnt < prio_fill; ++a)
        for (int b = a; b < npairs && cnt < prio_fill; ++b)
          if (!tab[a * npairs + b]) {
            tab[a * npairs + b] = 1;
            pl[2 * cnt] = a;
            pl[2 * cnt + 1] = b;
            ++cnt;
          }
      sc->pcount[r] = cnt;
    }
    sc->tabbuf = t_alloc(c, {(int64_t)(prio_h.size() / 8 + 2)});
    sc->plbuf = t_alloc(c, {(int64_t)(plist_h.size() / 2 + 2)});
    TNAD_CUDA(cudaMemcpyAsync(sc->tabbuf.p, prio_h.data(), prio_h.size(), cudaMemcpyHostToDevice, S1));
    TNAD_CUDA(cudaMemcpyAsync(sc->plbuf.p, plist_h.data(), plist_h.size() * sizeof(int), cudaMemcpyHostToDevice, S1));
    sync(c);
  }
  SymEigCache* sc = c->symcache;
  Tens& Mw = sc->Mw;
  Tens& Q = sc->Q;
  double* Wb[2] = {sc->Wbuf.p, sc->Wbuf2.p};
  int* sk[2] = {reinterpret_cast<int*>(sc->skipbuf.p), reinterpret_cast<int*>(sc->skipbuf2.p)};
  const unsigned char* prio_d = reinterpret_cast<const unsigned char*>(sc->tabbuf.p);
  const int* plist_d = reinterpret_cast<const int*>(sc->plbuf.p);

  TNAD_CUDA(cudaMemsetAsync(Mw.p, 0, (size_t)(ld * N) * sizeof(double), S1));
  const long long total_nn = n * n;
  const int nb_nn = (int)std::max<long long>(1, std::min<long long>((total_nn + 1023) / 1024, 148 * 8));
  set_identity(c, Q.p, ld, N);
  if (Q0 && Q0->rank == 2 && Q0->dim[0] == n && Q0->dim[1] == n) {
    // Warm start (successive CTMRG steps have nearly the same eigenvectors): M' = Q0' M Q0 is already close
    // to diagonal, so only the quadratically convergent tail of the Jacobi iteration is left to do.  Q0 is
    // re-orthogonalised (Newton-Schulz) first, so orthogonality errors do not accumulate across steps.
    Tens Ms = t_alloc(c, {n, n});
    k_load_sym<<<nb_nn, 256, 0, S1>>>(Ms.p, n, A.p, A.str[0], A.str[1], n, sym_add_transpose ? 1 : 0);
    LAUNCH_CHECK(c);
    Tens T = contract_new(c, "ki,kj->ij", *Q0, *Q0);
    k_ns_factor<<<nb_nn, 256, 0, S1>>>(T.p, n);
    LAUNCH_CHECK(c);
    Tens Qv = t_wrap(Q.p, {n, n});
    Qv.str[1] = ld;
    contract(c, "ik,kj->ij", *Q0, T, Qv, 1.0, 0.0);
    Tens T1 = contract_new(c, "ik,kj->ij", Ms, Qv);
    Tens Mp = contract_new(c, "ki,kj->ij", Qv, T1);
    k_load_sym<<<nb_nn, 256, 0, S1>>>(Mw.p, ld, Mp.p, 1, n, n, 0);   // exact symmetrisation (M' + M'^T)/2
    LAUNCH_CHECK(c);
  } else {
    k_load_sym<<<nb_nn, 256, 0, S1>>>(Mw.p, ld, A.p, A.str[0], A.str[1], n, sym_add_transpose ? 1 : 0);
    LAUNCH_CHECK(c);
  }
  {
    Tens Mflat = t_wrap(Mw.p, {ld * N});
    reduce(c, RED_SUMSQ, Mflat, nullptr, fro2);
  }
  double fro2h;
  d2h(c, &fro2h, fro2, 1);
  const double tol = tolfac * std::sqrt(fro2h);

  // ---- one sweep: all rounds, two streams -------------------------------------------------------------------
  // Stream 1 runs  eig_r -> priority-update_r -> eig_{r+1};  stream 2 runs the rest of update_r and the Q update
  // behind it, so the latency-bound pivot kernel (32 SMs) overlaps with the DMMA updates on the other SMs.
  unsigned long long* wc_m = c->ktiming ? reinterpret_cast<unsigned long long*>(c->scal + 20) : nullptr;
  int nodes = 0;
  auto issue_sweep = [&]() {
    nodes = 0;
    k_set_u64<<<1, 1, 0, S1>>>(offbits, 0ULL);
    ++nodes;
    if (!lookahead) {
      for (int r = 0; r < nr; ++r) {
        {
          KTimer kt(c, KF_EIG);
          k_pivot_eig<false><<<npairs, EW * 32, 0, S1>>>(Mw.p, ld, p, r, (int)n, fro2, tolfac, max_inner,
                                                         (cross_mode && r > 0) ? 1 : 0, Wb[0], sk[0], offbits, 0, 0.0);
        }
        {
          KTimer kt(c, KF_GRAM);
          k_sym_update_m<<<dim3(npairs, npairs), 256, smem_upd, S1>>>(Mw.p, ld, p, r, Wb[0], sk[0], 0, nullptr, nullptr,
                                                                     npairs, wc_m);
        }
        {
          KTimer kt(c, KF_UPDATE);
          jacobi_rotate_columns(c, Q.p, ld, chunks, p, r, Wb[0], sk[0], S1);
          c->launches--;
        }
        nodes += 3;
      }
      return;
    }
    {
      KTimer kt(c, KF_EIG);
      k_pivot_eig<false><<<npairs, EW * 32, 0, S1>>>(Mw.p, ld, p, 0, (int)n, fro2, tolfac, max_inner, 0, Wb[0], sk[0], offbits, 0,
                                                     0.0);
    }
    ++nodes;
    for (int r = 0; r < nr; ++r) {
      const int cur = r & 1;
      TNAD_CUDA(cudaEventRecord(c->ev_eig, S1));
      if (r > 0) TNAD_CUDA(cudaStreamWaitEvent(S1, c->ev_rest, 0));   // M fully updated by round r-1
      {
        KTimer kt(c, KF_GRAM);
        k_sym_update_m<<<sc->pcount[r], 256, smem_upd, S1>>>(Mw.p, ld, p, r, Wb[cur], sk[cur], 1,
                                                             plist_d + (size_t)r * sc->plist_stride,
                                                             prio_d + (size_t)r * npairs * npairs, npairs, wc_m);
      }
      TNAD_CUDA(cudaStreamWaitEvent(S2, c->ev_eig, 0));
      {
        KTimer kt(c, KF_GRAM, S2);
        k_sym_update_m<<<dim3(npairs, npairs), 256, smem_upd, S2>>>(Mw.p, ld, p, r, Wb[cur], sk[cur], 2, nullptr,
                                                                   prio_d + (size_t)r * npairs * npairs, npairs, wc_m);
      }
      TNAD_CUDA(cudaEventRecord(c->ev_rest, S2));
      // the next pivot kernel overwrites W[cur^1], which the Q update of round r-1 (stream 3) may still read:
      // wait for it here, before ev_v is re-recorded for round r
      if (r > 0) TNAD_CUDA(cudaStreamWaitEvent(S1, c->ev_v, 0));
      TNAD_CUDA(cudaStreamWaitEvent(S3, c->ev_eig, 0));
      {
        KTimer kt(c, KF_UPDATE, S3);
        jacobi_rotate_columns(c, Q.p, ld, chunks, p, r, Wb[cur], sk[cur], S3);
        c->launches--;
      }
      TNAD_CUDA(cudaEventRecord(c->ev_v, S3));
      nodes += 3;
      if (r + 1 < nr) {
        KTimer kt(c, KF_EIG);
        k_pivot_eig<false><<<npairs, EW * 32, 0, S1>>>(Mw.p, ld, p, r + 1, (int)n, fro2, tolfac, max_inner,
                                                       cross_mode ? 1 : 0, Wb[cur ^ 1], sk[cur ^ 1], offbits, 0, 0.0);
        ++nodes;
      }
    }
    TNAD_CUDA(cudaStreamWaitEvent(S1, c->ev_rest, 0));
    TNAD_CUDA(cudaStreamWaitEvent(S1, c->ev_v, 0));
  };

  if (use_graph && !sc->exec) {
    cudaGraph_t graph = nullptr;
    TNAD_CUDA(cudaStreamBeginCapture(S1, cudaStreamCaptureModeThreadLocal));
    try {
      issue_sweep();
    } catch (...) {
      cudaStreamEndCapture(S1, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    TNAD_CUDA(cudaStreamEndCapture(S1, &graph));
    TNAD_CUDA(cudaGraphInstantiate(&sc->exec, graph, 0));
    cudaGraphDestroy(graph);
    sc->nodes = nodes;
  }

  SvdResult res;
  int sweep = 0;
  bool converged = fro2h == 0.0;
  double prev_off = 1e300;
  for (; !converged && sweep < max_sweeps; ++sweep) {
    if (use_graph) {
      TNAD_CUDA(cudaGraphLaunch(sc->exec, S1));
      c->launches += sc->nodes;
    } else {
      issue_sweep();
      c->launches += nodes;
      TNAD_CUDA(cudaGetLastError());
    }
    double off;
    d2h(c, &off, c->scal + 16, 1);
    if (debug) fprintf(stderr, "[tnad symeig] n=%lld sweep %d off %.3e (tol %.1e)\n", (long long)n, sweep, off, tol);
    if (off <= tol) {
      converged = true;
      ++sweep;
      break;
    }
    if (sweep >= 6 && off < 1e3 * tol && off > 0.5 * prev_off) {   // rounding floor
      converged = true;
      ++sweep;
      break;
    }
    prev_off = off;
  }
  if (!converged) fail(TNAD_ERR_NOCONV, "svd_symmetric: block Jacobi did not converge in " + std::to_string(max_sweeps) + " sweeps");
  res.sweeps = sweep;

  Tens dg = t_alloc(c, {N});
  k_diag<<<(int)((N + 255) / 256), 256, 0, c->stream>>>(Mw.p, ld, N, dg.p);
  LAUNCH_CHECK(c);
  std::vector<double> lam((size_t)N);
  d2h(c, lam.data(), dg.p, (size_t)N);
  std::vector<int> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return std::fabs(lam[x]) > std::fabs(lam[y]); });
  std::vector<double> sval((size_t)n), sgn((size_t)n);
  for (int64_t r = 0; r < n; ++r) {
    sval[r] = std::fabs(lam[perm[r]]);
    sgn[r] = lam[perm[r]] < 0.0 ? -1.0 : 1.0;
  }
  Tens meta = t_alloc(c, {3 * n + 4});
  int* dperm = reinterpret_cast<int*>(meta.p);
  double* dsgn = meta.p + (n + 1) / 2 + 1;
  double* dsval = dsgn + n;
  TNAD_CUDA(cudaMemcpyAsync(dperm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  TNAD_CUDA(cudaMemcpyAsync(dsgn, sgn.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TNAD_CUDA(cudaMemcpyAsync(dsval, sval.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  Tens U = t_alloc(c, {n, n}), V = t_alloc(c, {n, n}), S = t_alloc(c, {n});
  k_sym_finalize<<<(int)n, 128, 0, c->stream>>>(Q.p, ld, dperm, dsgn, n, U.p, V.p);
  LAUNCH_CHECK(c);
  TNAD_CUDA(cudaMemcpyAsync(S.p, dsval, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  signfix_cols(c, U.p, n, n, V.p, n, n, n);
  sync(c);
  res.U = U;
  res.V = V;
  res.S = S;
  res.s_host = sval;
  res.null_thr = tol;
  return res;
}

}  // namespace tnad
