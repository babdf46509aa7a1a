"""NumPy prototype of the divide-and-conquer tridiagonal eigensolver implemented in csrc/stedc.cu (development aid:
same stages, same data layout, used to settle the numerics before writing the kernels; not part of the product and
not the parity oracle).

Stages per merge (Cuppen / Gu-Eisenstat, cf. LAPACK dlaed1-4 published algorithm):
  z, rho  -> sort d -> deflate (tiny z; close d via Givens chain) -> secular roots by safeguarded regula falsi in the
  variable shifted to the nearest pole -> z-hat (Loewner) -> eigenvectors -> S (m x m) -> Q_new = Q S
"""
import numpy as np

EPS = np.finfo(float).eps


def leaf_size(n, smax=64):
    L = 0
    while -(-n // (1 << L)) > smax:
        L += 1
    s = -(-n // (1 << L))
    return s, L


def bisect_bits(fun, hi):
    """Largest positive double x in (0, hi] ... root of increasing `fun` (fun(0+) < 0 <= fun(hi)) by bit-pattern bisection."""
    lo_b, hi_b = np.int64(0), np.float64(hi).view(np.int64)
    while hi_b - lo_b > 1:
        mid_b = lo_b + (hi_b - lo_b) // 2
        mid = np.int64(mid_b).view(np.float64)
        if fun(mid) >= 0:
            hi_b = mid_b
        else:
            lo_b = mid_b
    return np.int64(hi_b).view(np.float64)


EVALS = [0, 0]     # [function evaluations, roots]


def secular_root(dn, z2, rho, j):
    """Root j of 1 + rho sum z_i^2/(d_i - lam) in the variable x = |lam - d_org| shifted to the nearest pole:
    safeguarded regula falsi (Illinois) on phi(x) = F(x) x (Delta - x)/Delta (both neighbouring poles divided out),
    stopped by the dlaed4 criterion |F| <= 4 eps (1 + rho sum |terms|).  Returns (org, mu, d - d_org)."""
    k = dn.size
    if j < k - 1:
        gap = dn[j + 1] - dn[j]
        mid = 0.5 * gap
        fmid = 1.0 + rho * np.sum(z2 / ((dn - dn[j]) - mid))
        org, sgn = (j, 1.0) if fmid >= 0 else (j + 1, -1.0)
        hi, Dlt = mid, gap
    else:
        org, sgn, hi, Dlt = j, 1.0, rho * np.sum(z2) * (1 + 8 * EPS) + 1e-300, None
    delta = dn - dn[org]
    mask = np.ones(k, bool)
    mask[org] = False
    zo = z2[org]

    def ev(x):
        t = z2[mask] / (delta[mask] - sgn * x)
        Fx = (x + rho * x * np.sum(t) - rho * zo) if sgn > 0 else -(x + rho * x * np.sum(t) + rho * zo)
        S = 1.0 + rho * (np.sum(np.abs(t)) + zo / x)
        EVALS[0] += 1
        return Fx, S

    w = (lambda x: (Dlt - x) / Dlt) if Dlt is not None else (lambda x: 1.0)
    xa, fa, xb = 0.0, -rho * zo, hi
    Fb, _ = ev(xb)
    fb = Fb * w(xb)
    xc = xb
    EVALS[1] += 1
    if fb > 0:
        side = 0
        for _ in range(100):
            den = fb - fa
            xc = (xa * fb - xb * fa) / den if den != 0 else 0.5 * (xa + xb)
            if not (xa < xc < xb):
                xc = 0.5 * (xa + xb)
            if not (xa < xc < xb):
                xc = xb
                break
            Fc, Sc = ev(xc)
            if abs(Fc) <= 4 * EPS * Sc * xc:
                break
            fc = Fc * w(xc)
            if fc < 0:
                xa, fa = xc, fc
                if side == -1:
                    fb *= 0.5
                side = -1
            else:
                xb, fb = xc, fc
                if side == 1:
                    fa *= 0.5
                side = 1
    return org, sgn * xc, delta


def merge(d, Q, z, rho):
    """Eigen-decomposition of Q (diag(d) + rho z z') Q' given as (lam, Qnew); z normalised, rho > 0."""
    m = d.size
    order = np.argsort(d, kind="stable")
    tol = 8.0 * EPS * max(np.abs(d).max(), np.abs(z).max())
    Q = Q.copy()
    d = d.copy()
    z = z.copy()
    nd, df = [], []
    if rho * np.abs(z).max() <= tol:
        return d, Q
    pj = -1
    for j in order:
        if rho * abs(z[j]) <= tol:
            df.append(j)
            continue
        if pj < 0:
            pj = j
            continue
        s, c = z[pj], z[j]
        tau = np.hypot(c, s)
        t = d[j] - d[pj]
        c, s = c / tau, -s / tau
        if abs(t * c * s) <= tol:
            z[j], z[pj] = tau, 0.0
            qp, qj = Q[:, pj].copy(), Q[:, j].copy()
            Q[:, pj] = c * qp + s * qj
            Q[:, j] = -s * qp + c * qj
            tt = d[pj] * c * c + d[j] * s * s
            d[j] = d[pj] * s * s + d[j] * c * c
            d[pj] = tt
            df.append(pj)
        else:
            nd.append(pj)
        pj = j
    nd.append(pj)
    k = len(nd)
    dn, zn = d[nd], z[nd]
    # after rotations the kept d stay ascending (LAPACK relies on the same fact)
    assert np.all(np.diff(dn) > 0), "non-deflated poles must be strictly increasing"
    lam = np.empty(k)
    Delta = np.empty((k, k))          # Delta[i, j] = dn[i] - lam[j]
    z2 = zn * zn
    for j in range(k):
        org, mu, delta = secular_root(dn, z2, rho, j)
        lam[j] = dn[org] + mu
        Delta[:, j] = delta - mu
    # Loewner: zhat_i^2 = prod_j (lam_j - d_i) / (rho prod_{j != i} (d_j - d_i))
    zh = np.empty(k)
    for i in range(k):
        p = -Delta[i, i]                         # lam_i - d_i > 0
        for j in range(k):
            if j != i:
                p *= Delta[i, j] / (dn[i] - dn[j])
        zh[i] = np.sign(zn[i]) * np.sqrt(abs(p) / rho)
    U = zh[:, None] / Delta
    U /= np.linalg.norm(U, axis=0)
    S = np.zeros((m, m))
    for i in range(k):
        S[nd[i], :k] = U[i]
    for t, j in enumerate(df):
        S[j, k + t] = 1.0
    lam_all = np.concatenate([lam, d[df]])
    return lam_all, Q @ S


def stedc(d, e, smax=64):
    """All eigenpairs of the symmetric tridiagonal (d, e).  Returns (lam, Z) unsorted."""
    n = d.size
    s, L = leaf_size(n, smax)
    N = s << L
    scale = max(np.abs(d).max(), np.abs(e).max() if e.size else 0.0, np.finfo(float).tiny)
    dd = np.empty(N)
    ee = np.zeros(N)           # ee[i] couples i and i+1
    dd[:n] = d
    ee[:n - 1] = e
    dd[n:] = scale * (4.0 + np.arange(N - n) / max(N, 1))
    beta = np.zeros(N)
    for b in range(1, N // s):
        i = b * s
        beta[i] = ee[i - 1]
        dd[i - 1] -= abs(beta[i])
        dd[i] -= abs(beta[i])
    lam = np.empty(N)
    Q = np.zeros((N, N))
    for b in range(N // s):
        sl = slice(b * s, (b + 1) * s)
        T = np.diag(dd[sl]) + np.diag(ee[b * s:(b + 1) * s - 1], 1) + np.diag(ee[b * s:(b + 1) * s - 1], -1)
        w, V = np.linalg.eigh(T)
        lam[sl] = w
        Q[sl, sl] = V
    m = s
    while m < N:
        m2 = 2 * m
        for g in range(N // m2):
            lo, mid, hi = g * m2, g * m2 + m, (g + 1) * m2
            b = beta[mid]
            z = np.concatenate([Q[mid - 1, lo:mid], np.sign(b) * Q[mid, mid:hi]]) / np.sqrt(2.0)
            rho = 2.0 * abs(b)
            if rho == 0.0:
                continue
            lam[lo:hi], Q[lo:hi, lo:hi] = merge(lam[lo:hi], Q[lo:hi, lo:hi], z, rho)
        m = m2
    keep = np.argsort(lam)[:n]          # pads are the largest values
    return lam[keep], Q[:n][:, keep]


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, kind in [(50, "rand"), (200, "rand"), (300, "graded"), (257, "cluster"), (400, "wilk")]:
        if kind == "rand":
            d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
        elif kind == "graded":
            d, e = 10.0 ** (-rng.uniform(0, 18, n)), 10.0 ** (-rng.uniform(0, 18, n - 1))
        elif kind == "cluster":
            A = rng.standard_normal((n, n)); Qr, _ = np.linalg.qr(A)
            w = np.concatenate([np.ones(n // 3), np.zeros(n // 3), -np.ones(n - 2 * (n // 3))]) + 1e-14 * rng.standard_normal(n)
            M = (Qr * w) @ Qr.T
            import scipy.linalg as sl
            H = sl.hessenberg(M)
            d, e = np.diag(H).copy(), np.diag(H, 1).copy()
        else:
            d = np.abs(np.arange(n) - n // 2).astype(float); e = np.ones(n - 1)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        lam, Z = stedc(d, e, smax=16)
        o = np.argsort(lam)
        ref = np.linalg.eigvalsh(T)
        nrm = np.abs(ref).max()
        print(f"{kind:8s} n={n}: eig err {np.abs(lam[o] - ref).max() / nrm:.2e}  resid {np.abs(T @ Z - Z * lam).max() / nrm:.2e}  orth {np.abs(Z.T @ Z - np.eye(n)).max():.2e}")
