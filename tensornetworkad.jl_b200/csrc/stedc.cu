// Divide-and-conquer eigensolver for a symmetric tridiagonal matrix (Cuppen's method with the Gu-Eisenstat
// stabilisation; the published algorithm behind LAPACK dstedc/dlaed0-4, restated for the GPU).
//
// The order-n problem is padded to N = s 2^L (s <= 64) with decoupled diagonal entries above the spectrum, so every
// level of the merge tree is a batch of identical-size merges and all per-level work is a handful of kernels over
// (merge, element) grids plus ONE batched DMMA GEMM (contract, spec "gcik,gckj->gcij"):
//   leaves   k_dc_leaf      cyclic Jacobi on the s x s tridiagonal blocks in shared memory
//   level    k_dc_z         z = (last row of Q1, first row of Q2)/sqrt(2), rho = 2|beta|
//            k_dc_sort      rank sort of the poles
//            k_dc_deflate   tiny-z and close-pole deflation (one sequential scan per merge, Givens chain recorded)
//            k_dc_secular   one warp per root: safeguarded regula falsi (Illinois) in the variable shifted to the
//                           nearest pole, both neighbouring poles divided out (differences d_i - lambda_j come out
//                           with high relative accuracy); ~5 function evaluations per root
//            k_dc_zhat      Loewner formula for the modified z (orthogonal eigenvectors without extended precision)
//            k_dc_vectors   eigenvectors of D + rho z z' scattered into the row-major merge matrix S
//            k_dc_rot_s     the recorded Givens chain folded into S (so Q stays block diagonal for the GEMM)
//            contract       Q_new = blockdiag(Q1, Q2) S   (m^3 flops per merge instead of 2 m^3)
// Nothing synchronises with the host between levels.
#include "drivers.h"
#include "eigdc.h"
#include <algorithm>
#include <numeric>

namespace tnad {

namespace {

constexpr double DC_EPS = 2.220446049250313e-16;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wprod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- setup: padded tridiagonal, Cuppen adjustments at every leaf boundary ---------------------------------------
// dp (N): diagonal; ep (N): ep[i] couples i and i+1; beta (N): beta[i] = coupling removed between i-1 and i (i % s == 0)
__global__ void k_dc_setup(const double* __restrict__ dd, const double* __restrict__ ee, int n, int N, int s, double scale,
                           double* dp, double* ep, double* beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  // the problem is scaled to max(|d|, |e|) = 1 (LAPACK dstedc does the same with dlascl): the deflation tolerances
  // compare against the O(1) entries of z
  const double rs = 1.0 / scale;
  double d = i < n ? dd[i] * rs : (4.0 + (double)(i - n) / (double)N);
  const double el = (i > 0 && i - 1 < n - 1) ? ee[i - 1] * rs : 0.0;   // coupling (i-1, i)
  const double er = (i < n - 1) ? ee[i] * rs : 0.0;                    // coupling (i, i+1)
  double b = 0.0, e_out = er;
  if (i % s == 0 && i > 0) {
    b = el;
    d -= fabs(el);
  }
  if ((i + 1) % s == 0 && i + 1 < N) {
    d -= fabs(er);
    e_out = 0.0;
  }
  dp[i] = d;
  ep[i] = e_out;
  beta[i] = b;
}

__global__ void k_scale_vec(double* x, int n, double f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= f;
}

__global__ void k_absmax2(const double* __restrict__ a, int na, const double* __restrict__ b, int nbv, double* out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < na; i += blockDim.x) m = fmax(m, fabs(a[i]));
  for (int i = threadIdx.x; i < nbv; i += blockDim.x) m = fmax(m, fabs(b[i]));
  m = wmax(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, red[w]);
    out[0] = r;
  }
}

// ---- leaves: cyclic two-sided Jacobi in shared memory -------------------------------------------------------------
__device__ __forceinline__ void rr_pair_leaf(int se, int r, int k, int& p, int& q) {
  // round-robin tournament on se (even) players: player se-1 fixed
  const int mod = se - 1;
  if (k == 0) {
    p = se - 1;
    q = r % mod;
  } else {
    p = (r + k) % mod;
    q = (r - k + mod) % mod;
  }
  if (p > q) {
    const int tmp = p;
    p = q;
    q = tmp;
  }
}

__global__ void __launch_bounds__(256) k_dc_leaf(const double* __restrict__ dp, const double* __restrict__ ep, int s, int N,
                                                 double* lam, double* Q, long long ldq) {
  extern __shared__ double sh[];
  const int se = s + (s & 1), ld = se + 1;
  double* A = sh;                 // se x se
  double* V = A + se * ld;        // se x se
  double* cs = V + se * ld;       // 2 * (se/2)
  __shared__ double red[8];
  __shared__ int flag;
  const int t = threadIdx.x, base = blockIdx.x * s;
  for (int idx = t; idx < se * se; idx += 256) {
    const int i = idx % se, j = idx / se;
    double a = 0.0;
    if (i < s && j < s) {
      if (i == j) a = dp[base + i];
      else if (i == j + 1) a = ep[base + j];
      else if (j == i + 1) a = ep[base + i];
    }
    A[i + j * ld] = a;
    V[i + j * ld] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  double f2 = 0.0;
  for (int idx = t; idx < se * se; idx += 256) {
    const double a = A[(idx % se) + (idx / se) * ld];
    f2 += a * a;
  }
  f2 = wsum(f2);
  if ((t & 31) == 0) red[t >> 5] = f2;
  __syncthreads();
  f2 = 0.0;
  for (int w = 0; w < 8; ++w) f2 += red[w];
  const double thr = 0.25 * DC_EPS * sqrt(f2) / (double)se;   // |a_pq| below this is left alone
  const int np = se / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    if (t == 0) flag = 0;
    __syncthreads();
    for (int r = 0; r < se - 1; ++r) {
      if (t < np) {
        int p, q;
        rr_pair_leaf(se, r, t, p, q);
        double c = 1.0, sn = 0.0;
        const double apq = A[p + q * ld];
        if (p < s && q < s && fabs(apq) > thr) {
          const double theta = (A[q + q * ld] - A[p + p * ld]) / (2.0 * apq);
          const double tt = copysign(1.0, theta) / (fabs(theta) + sqrt(1.0 + theta * theta));
          c = 1.0 / sqrt(1.0 + tt * tt);
          sn = tt * c;
          flag = 1;
        }
        cs[2 * t] = c;
        cs[2 * t + 1] = sn;
      }
      __syncthreads();
      // rows: A <- J' A
      for (int idx = t; idx < np * se; idx += 256) {
        const int k = idx / se, col = idx % se;
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        if (sn != 0.0) {
          int p, q;
          rr_pair_leaf(se, r, k, p, q);
          const double x = A[p + col * ld], y = A[q + col * ld];
          A[p + col * ld] = c * x - sn * y;
          A[q + col * ld] = sn * x + c * y;
        }
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int idx = t; idx < np * se; idx += 256) {
        const int k = idx / se, row = idx % se;
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        if (sn != 0.0) {
          int p, q;
          rr_pair_leaf(se, r, k, p, q);
          double x = A[row + p * ld], y = A[row + q * ld];
          A[row + p * ld] = c * x - sn * y;
          A[row + q * ld] = sn * x + c * y;
          x = V[row + p * ld];
          y = V[row + q * ld];
          V[row + p * ld] = c * x - sn * y;
          V[row + q * ld] = sn * x + c * y;
        }
      }
      __syncthreads();
    }
    if (!flag) break;
    __syncthreads();
  }
  for (int i = t; i < s; i += 256) lam[base + i] = A[i + i * ld];
  for (int idx = t; idx < s * s; idx += 256) {
    const int i = idx % s, j = idx / s;
    Q[(base + i) + (long long)(base + j) * ldq] = V[i + j * ld];
  }
}

// ---- one level ----------------------------------------------------------------------------------------------------
// merge g covers [g*m2, (g+1)*m2), children of size m = m2/2.
__global__ void k_dc_z(const double* __restrict__ Q, long long ldq, const double* __restrict__ beta, int m, double* zu,
                       double* rho) {
  const int g = blockIdx.y, m2 = 2 * m, lo = g * m2;
  const int tt = blockIdx.x * blockDim.x + threadIdx.x;
  if (tt >= m2) return;
  const double b = beta[lo + m];
  const double sg = b < 0.0 ? -1.0 : 1.0;
  const double r = 0.7071067811865476;
  double z;
  if (tt < m) z = Q[(lo + m - 1) + (long long)(lo + tt) * ldq] * r;
  else z = sg * Q[(lo + m) + (long long)(lo + tt) * ldq] * r;
  zu[lo + tt] = z;
  if (tt == 0) rho[g] = 2.0 * fabs(b);
}

// rank sort of the m2 poles of each merge (ties by index); also max|d|, max|z| via atomics-free per-merge pass later
__global__ void __launch_bounds__(256) k_dc_sort(const double* __restrict__ du, const double* __restrict__ zu, int m2, int* order,
                                                 double* ds, double* zs) {
  __shared__ double tile[256];
  const int g = blockIdx.y, lo = g * m2;
  const int tt = blockIdx.x * 256 + threadIdx.x;
  const double mine = tt < m2 ? du[lo + tt] : 0.0;
  int rank = 0;
  for (int base = 0; base < m2; base += 256) {
    const int u = base + threadIdx.x;
    tile[threadIdx.x] = u < m2 ? du[lo + u] : 0.0;
    __syncthreads();
    const int lim = min(256, m2 - base);
    if (tt < m2) {
      for (int q = 0; q < lim; ++q) {
        const double v = tile[q];
        rank += (v < mine || (v == mine && base + q < tt)) ? 1 : 0;
      }
    }
    __syncthreads();
  }
  if (tt < m2) {
    order[lo + rank] = tt;
    ds[lo + rank] = mine;
    zs[lo + rank] = zu[lo + tt];
  }
}

// Deflation (sorted coordinates), one CTA per merge.  Tiny-z deflation is decided in parallel; the close-pole test
// is evaluated for every pair of neighbouring survivors in parallel, and only when one of them fires (rare) the
// sequential LAPACK-style scan with its Givens chain runs on thread 0.
// meta[g*4 + 0] = k (non-deflated), +1 = number of rotations, +2 = number deflated
__global__ void __launch_bounds__(256) k_dc_deflate(int m2, const double* __restrict__ rho_arr, double* ds, double* zs,
                                                    const int* __restrict__ order, int* ndl, int* dfl, int* rotp, int* rotj,
                                                    double* rotc, double* rots, double* dn, double* zn, int* meta) {
  __shared__ double red[8];
  __shared__ int scan[256];
  __shared__ int sk, snf, srot;
  const int g = blockIdx.x, lo = g * m2, t = threadIdx.x;
  double* d = ds + lo;
  double* z = zs + lo;
  const int* ord = order + lo;
  int* nd = ndl + lo;
  int* df = dfl + lo;
  const double rho = rho_arr[g];
  double md = 0.0, mz = 0.0;
  for (int i = t; i < m2; i += 256) {
    md = fmax(md, fabs(d[i]));
    mz = fmax(mz, fabs(z[i]));
  }
  md = wmax(md);
  mz = wmax(mz);
  if ((t & 31) == 0) red[t >> 5] = md;
  __syncthreads();
  md = 0.0;
  for (int w = 0; w < 8; ++w) md = fmax(md, red[w]);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = mz;
  __syncthreads();
  mz = 0.0;
  for (int w = 0; w < 8; ++w) mz = fmax(mz, red[w]);
  const double tol = 8.0 * DC_EPS * fmax(md, mz);
  if (t == 0) srot = 0;
  __syncthreads();
  if (rho * mz <= tol) {
    for (int i = t; i < m2; i += 256) df[i] = i;
    if (t == 0) {
      meta[g * 4 + 0] = 0;
      meta[g * 4 + 1] = 0;
      meta[g * 4 + 2] = m2;
    }
    return;
  }
  // ---- parallel pass: compaction of the survivors of the tiny-z test --------------------------------------------
  const int seg = (m2 + 255) / 256, i0 = t * seg, i1 = min(m2, i0 + seg);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += (rho * fabs(z[i]) > tol) ? 1 : 0;
  scan[t] = cnt;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const int v = (t >= off) ? scan[t - off] : 0;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  const int nc = scan[255];
  int kpos = scan[t] - cnt, dpos = i0 - kpos;
  for (int i = i0; i < i1; ++i) {
    if (rho * fabs(z[i]) > tol) nd[kpos++] = i;
    else df[dpos++] = i;
  }
  __syncthreads();
  for (int q = 1 + t; q < nc; q += 256) {
    const int pj = nd[q - 1], i = nd[q];
    double s = z[pj], c = z[i];
    const double tau = hypot(c, s);
    const double tt = d[i] - d[pj];
    c /= tau;
    s = -s / tau;
    if (fabs(tt * c * s) <= tol) srot = 1;
  }
  __syncthreads();
  if (!srot) {
    if (t == 0) {
      meta[g * 4 + 0] = nc;
      meta[g * 4 + 1] = 0;
      meta[g * 4 + 2] = m2 - nc;
    }
    for (int i = t; i < nc; i += 256) {
      const int p = nd[i];
      dn[lo + i] = d[p];
      zn[lo + i] = z[p];
    }
    return;
  }
  // ---- sequential scan with the Givens chain ------------------------------------------------------------------
  if (t == 0) {
    int k = 0, nf = 0, nr = 0, pj = -1;
    for (int i = 0; i < m2; ++i) {
      const double zi = z[i];
      if (rho * fabs(zi) <= tol) {
        df[nf++] = i;
        continue;
      }
      if (pj < 0) {
        pj = i;
        continue;
      }
      double s = z[pj], c = zi;
      const double tau = hypot(c, s);
      const double tt = d[i] - d[pj];
      c /= tau;
      s = -s / tau;
      if (fabs(tt * c * s) <= tol) {
        z[i] = tau;
        z[pj] = 0.0;
        rotp[lo + nr] = ord[pj];
        rotj[lo + nr] = ord[i];
        rotc[lo + nr] = c;
        rots[lo + nr] = s;
        ++nr;
        const double dpj = d[pj] * c * c + d[i] * s * s;
        d[i] = d[pj] * s * s + d[i] * c * c;
        d[pj] = dpj;
        df[nf++] = pj;
      } else {
        nd[k++] = pj;
      }
      pj = i;
    }
    if (pj >= 0) nd[k++] = pj;
    meta[g * 4 + 0] = k;
    meta[g * 4 + 1] = nr;
    meta[g * 4 + 2] = nf;
    sk = k;
    snf = nf;
  }
  __syncthreads();
  const int k = sk;
  for (int i = t; i < k; i += 256) {
    const int p = nd[i];
    dn[lo + i] = d[p];
    zn[lo + i] = z[p];
  }
}

// f(x) = 1 + rho sum_i z_i^2 / ((d_i - d_org) - sgn x)
__device__ __forceinline__ double secular_eval(const double* __restrict__ dn, const double* __restrict__ zn, int k, double rho,
                                               double dorg, double x, int lane) {
  double s = 0.0;
  for (int i = lane; i < k; i += 32) {
    const double zi = zn[i];
    s += zi * zi / ((dn[i] - dorg) - x);
  }
  return 1.0 + rho * wsum(s);
}

// One warp per root.  Dl[(lo+i) + j*ldd] = d_i - lambda_j (i, j < k), lam_out[lo+j] = lambda_j.
__global__ void __launch_bounds__(128) k_dc_secular(int m2, const double* __restrict__ rho_arr, const double* __restrict__ dn_all,
                                                    const double* __restrict__ zn_all, const int* __restrict__ meta, double* Dl,
                                                    long long ldd, double* lam_out) {
  const int g = blockIdx.y, lo = g * m2, lane = threadIdx.x & 31;
  const int k = meta[g * 4];
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= k) return;
  const double* dn = dn_all + lo;
  const double* zn = zn_all + lo;
  const double rho = rho_arr[g];
  int org;
  double hi, sgn;   // root = d_org + sgn * x, x in (0, hi]
  if (j < k - 1) {
    const double gap = dn[j + 1] - dn[j];
    const double mid = 0.5 * gap;
    const double f = secular_eval(dn, zn, k, rho, dn[j], mid, lane);
    if (f >= 0.0) {
      org = j;
      sgn = 1.0;
    } else {
      org = j + 1;
      sgn = -1.0;
    }
    hi = mid;
  } else {
    double s = 0.0;
    for (int i = lane; i < k; i += 32) s += zn[i] * zn[i];
    s = wsum(s);
    org = j;
    sgn = 1.0;
    hi = rho * s * (1.0 + 8.0 * DC_EPS) + 1e-300;
  }
  const double dorg = dn[org];
  const double zo = zn[org] * zn[org];
  const double Delta = (j < k - 1) ? 2.0 * hi : 0.0;     // distance to the pole on the other side (interior roots)
  // F(x) x and the error scale S = 1 + rho sum |terms|, with the pole term of the origin taken out analytically:
  //   F(x) x = +-(x + rho x sum_{i != org} z_i^2/(delta_i -+ x) -+ rho z_org^2)
  auto eval = [&](double x, double& Fx, double& S) {
    double st = 0.0, sa = 0.0;
    for (int i = lane; i < k; i += 32) {
      if (i != org) {
        const double zi = zn[i];
        const double tt = zi * zi / ((dn[i] - dorg) - sgn * x);
        st += tt;
        sa += fabs(tt);
      }
    }
    st = wsum(st);
    sa = wsum(sa);
    Fx = (sgn > 0.0) ? (x + rho * x * st - rho * zo) : -(x + rho * x * st + rho * zo);
    S = 1.0 + rho * (sa + zo / x);
  };
  // Illinois regula falsi on phi(x) = F(x) x (Delta - x)/Delta: both neighbouring poles are divided out, phi is
  // smooth and increasing on (0, hi], phi(0) = -rho z_org^2 is known exactly.  Stop when |F| <= 4 eps S (the LAPACK
  // dlaed4 criterion); the bracket is kept, so the worst case degrades to bisection.  Typically 5 evaluations
  // (tools/stedc_proto.py: 5-6 on random / graded / Wilkinson, 17 on heavily clustered spectra) instead of 62.
  double xa = 0.0, fa = -rho * zo, xb = hi, fb, Fb, Sb;
  eval(xb, Fb, Sb);
  fb = (Delta > 0.0) ? Fb * ((Delta - xb) / Delta) : Fb;
  double xc = xb;
  if (fb > 0.0) {
    int side = 0;
    for (int it = 0; it < 100; ++it) {
      const double den = fb - fa;
      xc = (den != 0.0) ? (xa * fb - xb * fa) / den : 0.5 * (xa + xb);
      if (!(xc > xa && xc < xb)) xc = 0.5 * (xa + xb);
      if (!(xc > xa && xc < xb)) {
        xc = xb;
        break;
      }
      double Fc, Sc;
      eval(xc, Fc, Sc);
      if (fabs(Fc) <= 4.0 * DC_EPS * Sc * xc) break;      // Fc carries the factor x
      const double fc = (Delta > 0.0) ? Fc * ((Delta - xc) / Delta) : Fc;
      if (fc < 0.0) {
        xa = xc;
        fa = fc;
        if (side == -1) fb *= 0.5;
        side = -1;
      } else {
        xb = xc;
        fb = fc;
        if (side == 1) fa *= 0.5;
        side = 1;
      }
    }
  }
  const double mu = sgn * xc;
  if (lane == 0) lam_out[lo + j] = dorg + mu;
  for (int i = lane; i < k; i += 32) Dl[(lo + i) + (long long)j * ldd] = (dn[i] - dorg) - mu;
}

// zh_i = sign(z_i) sqrt( (lambda_i - d_i) prod_{j != i} (lambda_j - d_i)/(d_j - d_i) )   (up to the common 1/sqrt(rho))
__global__ void __launch_bounds__(128) k_dc_zhat(int m2, const double* __restrict__ dn_all, const double* __restrict__ zn_all,
                                                 const int* __restrict__ meta, const double* __restrict__ Dl, long long ldd,
                                                 double* zh) {
  const int g = blockIdx.y, lo = g * m2, lane = threadIdx.x & 31;
  const int k = meta[g * 4];
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= k) return;
  const double* dn = dn_all + lo;
  const double di = dn[i];
  double p = 1.0;
  for (int j = lane; j < k; j += 32) {
    const double dl = Dl[(lo + i) + (long long)j * ldd];
    p *= (j == i) ? -dl : dl / (di - dn[j]);
  }
  p = wprod(p);
  if (lane == 0) zh[lo + i] = copysign(sqrt(fabs(p)), zn_all[lo + i]);
}

// Column j of the merge matrix (row-major S, S[(lo + r) * lds + j]): normalised zh_i / (d_i - lambda_j) at the original
// column index r = order[nd[i]]; deflated pair t sits in column k + t with a unit entry.  Also the merged eigenvalues.
__global__ void __launch_bounds__(128) k_dc_vectors(int m2, const int* __restrict__ meta, const int* __restrict__ order,
                                                    const int* __restrict__ ndl, const int* __restrict__ dfl,
                                                    const double* __restrict__ ds, const double* __restrict__ zh,
                                                    const double* __restrict__ Dl, long long ldd, double* S, long long lds,
                                                    double* lam_out) {
  __shared__ double red[4];
  const int g = blockIdx.y, lo = g * m2, t = threadIdx.x;
  const int k = meta[g * 4];
  const int j = blockIdx.x;
  if (j >= k) {
    if (t == 0) {
      const int p = dfl[lo + (j - k)];
      S[(long long)(lo + order[lo + p]) * lds + j] = 1.0;
      lam_out[lo + j] = ds[lo + p];
    }
    return;
  }
  double s = 0.0;
  for (int i = t; i < k; i += 128) {
    const double u = zh[lo + i] / Dl[(lo + i) + (long long)j * ldd];
    s += u * u;
  }
  s = wsum(s);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  const double inv = 1.0 / sqrt((red[0] + red[1]) + (red[2] + red[3]));
  for (int i = t; i < k; i += 128) {
    const double u = zh[lo + i] / Dl[(lo + i) + (long long)j * ldd];
    S[(long long)(lo + order[lo + ndl[lo + i]]) * lds + j] = u * inv;
  }
}

// Fold the Givens chain into S: S <- G_1 (G_2 ( ... (G_r S)))  (rows mix; threads run along the row-major rows)
__global__ void __launch_bounds__(256) k_dc_rot_s(int m2, const int* __restrict__ meta, const int* __restrict__ rotp,
                                                  const int* __restrict__ rotj, const double* __restrict__ rotc,
                                                  const double* __restrict__ rots, double* S, long long lds) {
  const int g = blockIdx.y, lo = g * m2;
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int nr = meta[g * 4 + 1];
  if (col >= m2 || nr == 0) return;
  for (int q = nr - 1; q >= 0; --q) {
    const int P = rotp[lo + q], J = rotj[lo + q];
    const double c = rotc[lo + q], s = rots[lo + q];
    double* xp = S + (long long)(lo + P) * lds + col;
    double* yp = S + (long long)(lo + J) * lds + col;
    const double x = *xp, y = *yp;
    *xp = c * x - s * y;
    *yp = s * x + c * y;
  }
}

void leaf_plan(const tnad_ctx* c, int64_t n, int& s, int& L) {
  const char* ev = opt_s(c, "TNAD_DC_LEAF");
  const int smax = ev ? std::max(4, std::min(64, atoi(ev))) : 16;   // measured at n = 2048: 16 -> 2.2 ms, 32 -> 2.5, 64 -> 4.9
  L = 0;
  while ((n + (1LL << L) - 1) / (1LL << L) > smax) ++L;
  s = (int)((n + (1LL << L) - 1) / (1LL << L));
  if (s < 2) s = 2;
}

}  // namespace

void stedc(tnad_ctx* c, const double* dd, const double* ee, int64_t n, Tens& lam, Tens& Z, int64_t& Nout) {
  TNAD_REQUIRE(n >= 1, "stedc: empty problem");
  int s, L;
  leaf_plan(c, n, s, L);
  const int64_t N = (int64_t)s << L;
  Nout = N;
  cudaStream_t st = c->stream;
  Tens dp = t_alloc(c, {N}), ep = t_alloc(c, {N}), beta = t_alloc(c, {N});
  Tens lamA = t_alloc(c, {N}), lamB = t_alloc(c, {N});
  Tens Qa = t_alloc(c, {N, N}, true), Qb = t_alloc(c, {N, N}, true);
  Tens scal = t_alloc(c, {4});
  k_absmax2<<<1, 256, 0, st>>>(dd, (int)n, ee, (int)std::max<int64_t>(n - 1, 0), scal.p);
  c->launches++;
  double scale;
  d2h(c, &scale, scal.p, 1);
  if (!(scale > 0.0)) scale = 1.0;
  k_dc_setup<<<(int)((N + 255) / 256), 256, 0, st>>>(dd, ee, (int)n, (int)N, s, scale, dp.p, ep.p, beta.p);
  c->launches++;
  {
    const int se = s + (s & 1);
    const size_t smem = (size_t)(2 * se * (se + 1) + se + 8) * sizeof(double);
    TNAD_CUDA(cudaFuncSetAttribute(k_dc_leaf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KTimer kt(c, KF_STEDC);
    k_dc_leaf<<<(int)(N / s), 256, smem, st>>>(dp.p, ep.p, s, (int)N, lamA.p, Qa.p, N);
    c->launches++;
    TNAD_CUDA(cudaGetLastError());
  }
  if (L > 0) {
    Tens zu = t_alloc(c, {N}), ds = t_alloc(c, {N}), zs = t_alloc(c, {N}), dn = t_alloc(c, {N}), zn = t_alloc(c, {N}),
         zh = t_alloc(c, {N}), rotc = t_alloc(c, {N}), rots = t_alloc(c, {N}), rho = t_alloc(c, {N / (2 * s) + 1});
    Tens ibuf = t_alloc(c, {3 * N + 8});   // order, ndl, dfl, rotp, rotj as ints (5N ints <= 3N doubles)
    int* order = reinterpret_cast<int*>(ibuf.p);
    int* ndl = order + N;
    int* dfl = ndl + N;
    int* rotp = dfl + N;
    int* rotj = rotp + N;
    Tens metab = t_alloc(c, {2 * (N / (2 * s)) + 4});
    int* meta = reinterpret_cast<int*>(metab.p);
    Tens Dl = t_alloc(c, {N, N}), S = t_alloc(c, {N, N});
    double* lam_in = lamA.p;
    double* lam_out = lamB.p;
    Tens* Qin = &Qa;
    Tens* Qout = &Qb;
    const bool debug = opt_i(c, "TNAD_DC_DEBUG", 0) > 1;
    for (int lvl = 0; lvl < L; ++lvl) {
      const int m = s << lvl, m2 = 2 * m;
      cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr, e4 = nullptr;
      if (debug) {
        e0 = get_event(c); e1 = get_event(c); e2 = get_event(c); e3 = get_event(c); e4 = get_event(c);
        TNAD_CUDA(cudaEventRecord(e0, st));
      }
      const int G = (int)(N / m2);
      const int64_t ldd = N, lds = m2;
      k_dc_z<<<dim3((m2 + 255) / 256, G), 256, 0, st>>>(Qin->p, N, beta.p, m, zu.p, rho.p);
      k_dc_sort<<<dim3((m2 + 255) / 256, G), 256, 0, st>>>(lam_in, zu.p, m2, order, ds.p, zs.p);
      {
        KTimer kt(c, KF_STEDC);
        k_dc_deflate<<<G, 256, 0, st>>>(m2, rho.p, ds.p, zs.p, order, ndl, dfl, rotp, rotj, rotc.p, rots.p, dn.p, zn.p, meta);
      }
      if (debug) TNAD_CUDA(cudaEventRecord(e1, st));
      {
        KTimer kt(c, KF_STEDC);
        k_dc_secular<<<dim3((m2 + 3) / 4, G), 128, 0, st>>>(m2, rho.p, dn.p, zn.p, meta, Dl.p, ldd, lam_out);
      }
      if (debug) TNAD_CUDA(cudaEventRecord(e2, st));
      k_dc_zhat<<<dim3((m2 + 3) / 4, G), 128, 0, st>>>(m2, dn.p, zn.p, meta, Dl.p, ldd, zh.p);
      TNAD_CUDA(cudaMemsetAsync(S.p, 0, (size_t)N * m2 * sizeof(double), st));
      k_dc_vectors<<<dim3(m2, G), 128, 0, st>>>(m2, meta, order, ndl, dfl, ds.p, zh.p, Dl.p, ldd, S.p, lds, lam_out);
      k_dc_rot_s<<<dim3((m2 + 255) / 256, G), 256, 0, st>>>(m2, meta, rotp, rotj, rotc.p, rots.p, S.p, lds);
      c->launches += 7;
      TNAD_CUDA(cudaGetLastError());
      if (debug) TNAD_CUDA(cudaEventRecord(e3, st));
      // Q_out[g] = blockdiag(Q1, Q2) S_g as one batched GEMM over (merge g, child c)
      Tens Av, Bv, Cv;
      Av.p = Qin->p;
      Bv.p = S.p;
      Cv.p = Qout->p;
      Av.rank = Bv.rank = Cv.rank = 4;
      const int64_t da[4] = {G, 2, m, m}, sa[4] = {(int64_t)m2 * (1 + N), (int64_t)m * (1 + N), 1, N};
      const int64_t db[4] = {G, 2, m, m2}, sb[4] = {(int64_t)m2 * m2, (int64_t)m * m2, m2, 1};
      const int64_t dc[4] = {G, 2, m, m2}, sc[4] = {(int64_t)m2 * (1 + N), m, 1, N};
      for (int q = 0; q < 4; ++q) {
        Av.dim[q] = da[q];
        Av.str[q] = sa[q];
        Bv.dim[q] = db[q];
        Bv.str[q] = sb[q];
        Cv.dim[q] = dc[q];
        Cv.str[q] = sc[q];
      }
      contract(c, "gcik,gckj->gcij", Av, Bv, Cv);
      if (debug) {
        TNAD_CUDA(cudaEventRecord(e4, st));
        TNAD_CUDA(cudaEventSynchronize(e4));
        float a, b, d3, d4;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        cudaEventElapsedTime(&d3, e2, e3);
        cudaEventElapsedTime(&d4, e3, e4);
        std::vector<int> mh((size_t)4 * G);
        TNAD_CUDA(cudaMemcpy(mh.data(), meta, mh.size() * sizeof(int), cudaMemcpyDeviceToHost));
        long long ks = 0, rs = 0;
        for (int g = 0; g < G; ++g) {
          ks += mh[4 * g];
          rs += mh[4 * g + 1];
        }
        fprintf(stderr, "[tnad dc] level m2=%d G=%d: z/sort/deflate %.3f  secular %.3f  zhat/vectors/rot %.3f  gemm %.3f ms  (kept %lld of %lld, rotations %lld)\n",
                m2, G, a, b, d3, d4, ks, (long long)N, rs);
      }
      std::swap(Qin, Qout);
      std::swap(lam_in, lam_out);
    }
    lam = (lam_in == lamA.p) ? lamA : lamB;
    Z = *Qin;
  } else {
    lam = lamA;
    Z = Qa;
  }
  k_scale_vec<<<(int)((N + 255) / 256), 256, 0, st>>>(lam.p, (int)N, scale);
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
}

}  // namespace tnad
