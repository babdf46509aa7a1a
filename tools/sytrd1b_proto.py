"""NumPy prototype of the ONE-barrier-per-column blocked tridiagonalisation of csrc/tridiag.cu (development aid).

Classic dlatrd needs two global exchanges per column: the matrix-vector product y = A v, and then the next column
(whose norm defines the next reflector) once w is known.  Here the matvec is done on a vector g that every row owner
can form WITHOUT the pending global scalar of the previous column (c_p = tau_p^2 (y_p'v_p)/2):

    x = g + 2 c_p v_p                       (x: current column of the implicitly updated matrix, rows > j)
    v = s (x - beta e_jn)                   (s = 1/(alpha - beta))
    A v = s (A g + 2 c_p (A v_p restricted) - beta A[:, jn])

so one exchange per column carries: the matvec partials of A g, the panel dots with g, and a handful of scalars
(sum g^2, sum g v_p, sum v_p^2, y_p'v_p).  Everything marked GLOBAL below is what all CTAs know after the exchange;
everything else is computed on owned rows only.  The prototype keeps that discipline by only using full-vector
reductions where the kernel exchanges partial sums.
"""
import numpy as np


THETA = 1e-3
REDO = [0]


def sytrd_one_barrier(A, nb=32):
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(n - 1)
    tau = np.zeros(n)
    Vh = np.zeros((n, n))
    nref = n - 2
    for j0 in range(0, nref, nb):
        nbc = min(nb, nref - j0)
        V = np.zeros((n, nb))
        W = np.zeros((n, nb))      # stores w0 = tau*y until the fix-up with c of that column
        # state carried between columns (row-owned vectors / global scalars)
        vp = np.zeros(n); w0p = np.zeros(n); yp = np.zeros(n); Avp = np.zeros(n)
        tau_p = 0.0
        svp = np.zeros(nb); swp = np.zeros(nb)      # GLOBAL: V_k'v_p, W_k'v_p (k < i-1 valid), rows >= j
        g = np.zeros(n)
        g[j0 + 1:] = A[j0 + 1:, j0]                   # first column of the panel: fully updated by the trailing GEMM
        i = 0
        pending = False             # True while the previous column's scalar c_p is still unknown to the row owners
        while i < nbc:
            j = j0 + i
            jn = j + 1
            R = slice(jn, n)            # rows >= jn
            R1 = slice(jn + 1, n)       # rows > jn
            # ---------------- pre-barrier: partials (here: full sums) ----------------
            u = np.zeros(n)
            u[R] = A[R, R] @ g[R]                               # matvec on g (column ownership)
            Vg = V[R, :i].T @ g[R]                              # 2i panel dots with g
            Wg = W[R, :i].T @ g[R]                              # (W[:, i-1] still holds w0_p)
            s_gg = g[R1] @ g[R1]
            s_gv = g[R1] @ vp[R1]
            s_vv = vp[R1] @ vp[R1]
            gam_p = yp[j:] @ vp[j:]                             # y_p'v_p (rows >= j); zero for i == 0
            # ---------------- barrier; GLOBAL scalars ----------------
            c_p = 0.5 * tau_p * tau_p * gam_p if pending else 0.0
            alpha = g[jn] + 2.0 * c_p * vp[jn]
            sig2 = s_gg + 4.0 * c_p * s_gv + 4.0 * c_p * c_p * s_vv
            if pending and sig2 < THETA * (s_gg + 4.0 * c_p * c_p * s_vv):
                # cancellation: the expansion cannot deliver |x| accurately.  Fold c_p in (x exact on owned rows), finish the
                # previous W column and redo this column's exchange with nothing pending (costs one extra barrier).
                W[j:, i - 1] = w0p[j:] - c_p * vp[j:]
                g = g + 2.0 * c_p * vp
                g[:jn] = 0.0
                Avp = np.zeros(n)            # not used when nothing is pending
                pending = False
                REDO[0] += 1
                continue
            if sig2 == 0.0:
                beta, tj, s = alpha, 0.0, 0.0
            else:
                beta = -np.copysign(np.sqrt(alpha * alpha + sig2), alpha)
                tj = (beta - alpha) / beta
                s = 1.0 / (alpha - beta)
            # d_j from row j of the panels (W[j, i-1] fixed up with c_p; V[j, i-1] = 1)
            if i > 0:
                Wj_row = W[j, :i].copy()
                if pending:
                    Wj_row[i - 1] = w0p[j] - c_p * vp[j]
                d[j] = A[j, j] - 2.0 * (V[j, :i] @ Wj_row)
            else:
                d[j] = A[j, j]
            e[j] = beta
            tau[j] = tj
            # panel dots with v (affine in the exchanged sums)
            sv = np.zeros(nb); sw = np.zeros(nb)
            if s != 0.0:
                for k in range(i):
                    if k < i - 1 or not pending:
                        Vk_vp = svp[k] - V[j, k]                 # rows >= jn part of V_k'v_p (multiplied by c_p = 0 if nothing pends)
                        Wk_vp = swp[k] - W[j, k]
                        Wk_g = Wg[k]
                        Wk_jn = W[jn, k]
                    else:                                        # k = i-1: V_k = v_p, W_k = w0_p - c_p v_p
                        vv = s_vv + vp[jn] * vp[jn]
                        gv = s_gv + g[jn] * vp[jn]
                        Vk_vp = vv
                        Wk_vp = (tau_p * gam_p - w0p[j]) - c_p * vv
                        Wk_g = Wg[k] - c_p * gv
                        Wk_jn = w0p[jn] - c_p * vp[jn]
                    sv[k] = s * (Vg[k] + 2.0 * c_p * Vk_vp - beta * V[jn, k])
                    sw[k] = s * (Wk_g + 2.0 * c_p * Wk_vp - beta * Wk_jn)
            else:
                # v = e_jn: plain rows of the panels
                for k in range(i):
                    sv[k] = V[jn, k]
                    sw[k] = (w0p[jn] - c_p * vp[jn]) if (k == i - 1 and pending) else W[jn, k]
            # ---------------- owned rows ----------------
            if i > 0 and pending:
                W[j:, i - 1] = w0p[j:] - c_p * vp[j:]            # fix-up of the previous W column
            x = g + 2.0 * c_p * vp
            v = np.zeros(n)
            v[R1] = s * x[R1]
            v[jn] = 1.0
            Av = np.zeros(n)
            if s != 0.0:
                Av[R] = s * (u[R] + 2.0 * c_p * (Avp[R] - A[R, j]) - beta * A[R, jn])
            else:
                Av[R] = A[R, jn]
            y = np.zeros(n)
            y[R] = Av[R] - V[R, :i] @ sw[:i] - W[R, :i] @ sv[:i]
            w0 = tj * y
            V[:, i] = v
            W[:, i] = w0
            Vh[:, j] = v
            # next g (rows > jn): needs the GLOBAL scalar w0[jn] (every CTA recomputes y[jn] from the row-jn data)
            gn = np.zeros(n)
            if i + 1 < nbc:
                jq = jn                       # next column index
                Rn = slice(jq + 1, n)
                xt = A[Rn, jq] - V[Rn, :i] @ W[jq, :i] - W[Rn, :i] @ V[jq, :i]
                gn[Rn] = xt - v[Rn] * w0[jq] - w0[Rn] * v[jq]
            # shift state
            svp, swp = sv, sw
            svp = svp.copy(); swp = swp.copy()
            vp, w0p, yp, Avp, tau_p, g = v, w0, y, Av, tj, gn
            pending = True
            i += 1
        # ---------------- panel end: one more exchange for the last c ----------------
        jl = j0 + nbc - 1
        gam = yp[jl + 1:] @ vp[jl + 1:]
        c_l = 0.5 * tau_p * tau_p * gam
        W[:, nbc - 1] = w0p - c_l * vp
        jt = j0 + nbc
        A[jt:, jt:] -= V[jt:, :nbc] @ W[jt:, :nbc].T + W[jt:, :nbc] @ V[jt:, :nbc].T
    d[n - 2] = A[n - 2, n - 2]
    d[n - 1] = A[n - 1, n - 1]
    e[n - 2] = A[n - 1, n - 2]
    return d, e, Vh, tau


def form_q(Vh, tau):
    n = Vh.shape[0]
    Q = np.eye(n)
    for j in range(n - 3, -1, -1):
        v = Vh[:, j]
        Q -= tau[j] * np.outer(v, v @ Q)
    return Q


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for n, nb in [(6, 4), (12, 4), (40, 8), (97, 32), (200, 32), (150, 32), (160, 32)]:
        A = rng.standard_normal((n, n)); A = A + A.T
        if n == 40:
            A[:, 5:9] = 0; A[5:9, :] = 0      # exact rank deficiency
        if n == 150:                          # numerically low rank with a decaying spectrum (CTMRG-like)
            Qr, _ = np.linalg.qr(A)
            A = (Qr * (10.0 ** (-np.arange(n) / 4.0))) @ Qr.T
            A = A + A.T
        if n == 160:                          # exact low rank
            B = rng.standard_normal((n, 20)); A = B @ B.T
        REDO[0] = 0
        d, e, Vh, tau = sytrd_one_barrier(A, nb)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        Q = form_q(Vh, tau)
        print(f"n={n} nb={nb}: |A - Q T Q'| {np.abs(Q @ T @ Q.T - A).max():.2e}  orth {np.abs(Q.T @ Q - np.eye(n)).max():.2e}  "
              f"eig {np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A)).max() / np.abs(A).max():.2e}  redo {REDO[0]}")
