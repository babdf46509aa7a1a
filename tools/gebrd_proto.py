"""NumPy statement of the planned round-2 route for the TRG splits (development aid, not product code):

    A (m x n, m >= n)  --Golub-Kahan-->  U1 B V1'   (B upper bidiagonal: d_1..d_n, e_1..e_{n-1})
    T_GK = perfect shuffle of [0 B'; B 0]: symmetric tridiagonal of order 2n with ZERO diagonal and off-diagonals
           d_1, e_1, d_2, e_2, ..., d_n   ->  eigenpairs (+-sigma_i, interleaved (v_i, +-u_i)/sqrt(2))
    divide and conquer on T_GK (tools/stedc_proto.py = csrc/stedc.cu), de-interleave, back-transform with U1, V1.

Compared with the Jordan-Wielandt embedding that csrc/tridiag.cu uses today (tridiagonalise the dense (m+n) x (m+n)
matrix), the reduction touches A itself: 1/4 of the matrix traffic and of the back-transform flops.

`gebrd_columns` is written in the data flow a column-ownership kernel would use (every column of A lives with one CTA):
  left reflector  : the owner of column j forms u; every owner applies I - tau u u' to ITS columns in one
                    read+write pass (z_c = u'a_c is column-local, no exchange between the dot and the update)
  right reflector : row j is gathered (one entry per column), v formed redundantly, w = A v is a reduction over the
                    owners' partial sums, then every owner updates its columns a_c -= tau_v w v_c (fused with the next
                    left pass in a kernel).
"""
import importlib.util
import os

import numpy as np


def _house(x):
    """LAPACK dlarfg: H = I - tau v v', v[0] = 1, H x = beta e_1."""
    alpha = x[0]
    xn2 = float(x[1:] @ x[1:])
    if xn2 == 0.0:
        return np.concatenate([[1.0], np.zeros(x.size - 1)]), 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, (beta - alpha) / beta, beta


def gebrd_columns(A):
    """Unblocked Golub-Kahan bidiagonalisation, column-ownership formulation.  Returns d, e, the left / right reflectors."""
    A = A.copy()
    m, n = A.shape
    assert m >= n
    d, e = np.zeros(n), np.zeros(max(n - 1, 0))
    UL, tauL = np.zeros((m, n)), np.zeros(n)
    VR, tauR = np.zeros((n, n)), np.zeros(n)
    for j in range(n):
        # ---- left: column j (rows j..m-1), owned by one CTA ----
        u, tu, beta = _house(A[j:, j].copy())
        d[j] = beta
        UL[j:, j], tauL[j] = u, tu
        # every owner, its own columns c > j: one pass, the dot is column-local
        for c in range(j + 1, n):
            z = u @ A[j:, c]
            A[j:, c] -= tu * z * u
        if j < n - 2:
            # ---- right: row j (cols j+1..n-1): one entry per column -> gathered, v formed redundantly ----
            v, tv, beta = _house(A[j, j + 1:].copy())
            e[j] = beta
            VR[j + 1:, j], tauR[j] = v, tv
            # w = A[j+1:, j+1:] v : partial sums over the owners' columns, reduced in one exchange
            w = np.zeros(m - j - 1)
            for c in range(j + 1, n):
                w += A[j + 1:, c] * v[c - j - 1]
            for c in range(j + 1, n):                      # column-local update (fused with the next left pass)
                A[j + 1:, c] -= tv * w * v[c - j - 1]
        elif j == n - 2:
            e[j] = A[j, j + 1]
    return d, e, (UL, tauL), (VR, tauR)


def _apply_reflectors(V, tau, X):
    """X <- H_0 H_1 ... H_{k-1} X."""
    for j in range(V.shape[1] - 1, -1, -1):
        if tau[j] != 0.0:
            v = V[:, j]
            X -= tau[j] * np.outer(v, v @ X)
    return X


def svd_via_gk(A, smax=16):
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("stedc_proto", os.path.join(here, "stedc_proto.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    m, n = A.shape
    d, e, (UL, tauL), (VR, tauR) = gebrd_columns(A)
    off = np.empty(2 * n - 1)
    off[0::2] = d
    off[1::2] = e
    lam, Z = P.stedc(np.zeros(2 * n), off, smax=smax)       # T_GK: zero diagonal
    order = np.argsort(-lam)[:n]                            # the n non-negative eigenvalues, descending
    s = lam[order]
    V2 = np.sqrt(2.0) * Z[0::2, order]                      # T_GK z = sigma z with z = interleave(v, u)/sqrt(2)
    U2 = np.sqrt(2.0) * Z[1::2, order]
    U = np.zeros((m, n))
    U[:n] = U2
    U = _apply_reflectors(UL, tauL, U)
    V = _apply_reflectors(VR, tauR, V2.copy())
    return U, s, V


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for m, n, kind in [(12, 12, "rand"), (60, 60, "rand"), (80, 50, "rand"), (64, 64, "lowrank"), (100, 100, "graded")]:
        A = rng.standard_normal((m, n))
        if kind == "lowrank":
            A = rng.standard_normal((m, 9)) @ rng.standard_normal((9, n))
        if kind == "graded":
            A = A * 10.0 ** (-rng.uniform(0, 10, n))[None, :]
        U, s, V = svd_via_gk(A)
        ref = np.linalg.svd(A, compute_uv=False)
        r = int(np.sum(ref > 1e-13 * ref[0]))
        print(f"{kind:8s} {m}x{n}: sigma {np.abs(s - ref).max() / ref[0]:.2e}  recon {np.abs((U * s) @ V.T - A).max() / ref[0]:.2e}  "
              f"orthU_r {np.abs(U[:, :r].T @ U[:, :r] - np.eye(r)).max():.2e}  orthV_r {np.abs(V[:, :r].T @ V[:, :r] - np.eye(r)).max():.2e}")
