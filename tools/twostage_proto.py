"""NumPy statement of the two-stage tridiagonalisation the CUDA kernels implement (csrc/band.cu):

  stage 1  sy2sb : A = Q1 B Q1'   dense symmetric -> symmetric band (half bandwidth b), blocked Householder QR of the
                                   sub-band panels, two-sided compact-WY update of the trailing matrix (all GEMMs)
  stage 2  sb2st : B = Q2 T Q2'   band -> tridiagonal by bulge chasing (one Householder reflector of length <= b per
                                   hop; sweep s annihilates column s; hop t acts on rows s+1+t b .. s+(t+1) b)
  back     U = Q1 (Q2 E)          E = eigenvectors of T; Q2 applied sweep by sweep (last sweep first: within a sweep
                                   the hops act on disjoint rows), Q1 panel by panel in compact-WY form

The data layouts (lower band storage with 2b diagonals for the bulge, reflector store V2[s] = the concatenated hops of
sweep s, i.e. one vector of length n-1-s per sweep) are the ones the kernels use.  Tested on the CPU by
tests/test_host_logic.py::test_twostage_proto.
"""
import numpy as np


def house(x):
    """LAPACK dlarfg: H = I - tau v v', v[0] = 1, H x = beta e_1."""
    x = np.asarray(x, dtype=float)
    alpha = x[0]
    xn = np.linalg.norm(x[1:])
    v = x.copy()
    if xn == 0.0:
        v[:] = 0.0
        v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def sy2sb(A, b):
    """Returns the band matrix (dense storage, for checking), and the panels [(row offset, Y, T)]."""
    A = A.copy()
    n = A.shape[0]
    panels = []
    j = 0
    while j + b < n - 1:
        r0 = j + b
        m = n - r0
        P = A[r0:, j:j + b].copy()
        kb = min(b, m - 1) if m > 1 else 0
        Y = np.zeros((m, b))
        taus = np.zeros(b)
        for c in range(kb):                      # Householder QR of the panel (one reflector per column)
            v, tau, beta = house(P[c:, c])
            Y[c:, c] = v
            taus[c] = tau
            P[c:, c] = 0.0
            P[c, c] = beta
            if c + 1 < b:
                w = tau * (v @ P[c:, c + 1:])
                P[c:, c + 1:] -= np.outer(v, w)
        for c in range(kb, b):                   # columns without a reflector (short last panel): H = I
            Y[:, c] = 0.0
        T = np.zeros((b, b))                     # dlarft, forward / columnwise
        G = Y.T @ Y
        for c in range(b):
            T[:c, c] = -taus[c] * (T[:c, :c] @ G[:c, c])
            T[c, c] = taus[c]
        A[r0:, j:j + b] = P
        A[j:j + b, r0:] = P.T
        A22 = A[r0:, r0:]
        Z = A22 @ (Y @ T)
        W = Z - 0.5 * Y @ (T.T @ (Y.T @ Z))
        A22 -= Y @ W.T + W @ Y.T
        panels.append((r0, Y, T))
        j += b
    return A, panels


def apply_q1(panels, X):
    """X <- Q1 X."""
    X = X.copy()
    for r0, Y, T in reversed(panels):
        X[r0:] -= Y @ (T @ (Y.T @ X[r0:]))
    return X


def to_band(Bd, b):
    """Lower band storage AB[i-j, j], 2b diagonals (0 .. 2b-1): room for the bulge."""
    n = Bd.shape[0]
    AB = np.zeros((2 * b, n))
    for j in range(n):
        hi = min(n, j + b + 1)
        AB[:hi - j, j] = Bd[j:hi, j]
    return AB


def sb2st(AB, b):
    """Bulge chasing on the lower band storage (in place).  Returns d, e, V2 (list per sweep: vector of length n-1-s,
    the hops' reflectors back to back), tau2 (list per sweep: one tau per hop)."""
    LD, n = AB.shape
    assert LD == 2 * b

    def get(i, j):                  # element (i, j), i >= j
        return AB[i - j, j]

    def blk(r0, r1, c0, c1):        # dense copy of rows r0:r1, cols c0:c1 of the symmetric band matrix
        out = np.zeros((r1 - r0, c1 - c0))
        for i in range(r0, r1):
            for j in range(c0, c1):
                lo, hi = (j, i) if i >= j else (i, j)
                if hi - lo < LD:
                    out[i - r0, j - c0] = AB[hi - lo, lo]
        return out

    def put_lower(r0, c0, M, sym=False):   # write back the part with i >= j (all of it for an off-diagonal block)
        for i in range(M.shape[0]):
            for j in range(M.shape[1]):
                gi, gj = r0 + i, c0 + j
                if gi >= gj:
                    assert gi - gj < LD, (gi, gj)
                    AB[gi - gj, gj] = M[i, j]

    V2, tau2 = [], []
    for s in range(n - 2):
        vs = np.zeros(n - 1 - s)
        ts = []
        # hop 0: annihilate column s below the first sub-diagonal
        p = s + 1
        L = min(b, n - p)
        x = np.array([get(p + k, s) for k in range(L)])
        v, tau, beta = house(x)
        AB[1, s] = beta
        for k in range(1, L):
            AB[1 + k, s] = 0.0
        vs[:L] = v
        ts.append(tau)
        D = blk(p, p + L, p, p + L)
        w = tau * (D @ v)
        w -= 0.5 * tau * (w @ v) * v
        D -= np.outer(v, w) + np.outer(w, v)
        put_lower(p, p, D)
        t = 1
        while True:
            p = s + 1 + t * b
            if p >= n:
                break
            Lp = L                                  # length of the previous reflector (columns p-Lp .. p-1)
            L = min(b, n - p)
            E = blk(p, p + L, p - Lp, p)            # rows of this hop, columns of the previous one
            E -= tau * np.outer(E @ v, v)           # right-apply the previous reflector: creates the bulge
            vprev = v
            v, tau, beta = house(E[:, 0])           # annihilate the first column of the bulge
            E[:, 0] = 0.0
            E[0, 0] = beta
            if Lp > 1:
                wv = tau * (v @ E[:, 1:])
                E[:, 1:] -= np.outer(v, wv)
            put_lower(p, p - Lp, E)
            vs[t * b:t * b + L] = v
            ts.append(tau)
            D = blk(p, p + L, p, p + L)
            w = tau * (D @ v)
            w -= 0.5 * tau * (w @ v) * v
            D -= np.outer(v, w) + np.outer(w, v)
            put_lower(p, p, D)
            t += 1
        V2.append(vs)
        tau2.append(np.array(ts))
    d = AB[0, :].copy()
    e = AB[1, :n - 1].copy()
    return d, e, V2, tau2


def apply_q2(V2, tau2, b, X):
    """X <- Q2 X: sweeps last to first; the hops of one sweep act on disjoint row blocks."""
    X = X.copy()
    n = X.shape[0]
    for s in range(len(V2) - 1, -1, -1):
        vs, ts = V2[s], tau2[s]
        for t in range(len(ts)):
            p = s + 1 + t * b
            L = min(b, n - p)
            v = vs[t * b:t * b + L]
            X[p:p + L] -= ts[t] * np.outer(v, v @ X[p:p + L])
    return X


def eigh_twostage(A, b):
    n = A.shape[0]
    Bd, panels = sy2sb(A, b)
    AB = to_band(Bd, b)
    d, e, V2, tau2 = sb2st(AB, b)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, E = np.linalg.eigh(T)
    U = apply_q1(panels, apply_q2(V2, tau2, b, E))
    return lam, U, (Bd, d, e)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, b in [(37, 4), (64, 8), (50, 16), (9, 4), (6, 4), (5, 4), (130, 32)]:
        A = rng.standard_normal((n, n))
        A = A + A.T
        lam, U, (Bd, d, e) = eigh_twostage(A, b)
        band_ok = np.abs(np.tril(Bd, -b - 1)).max() if n > b + 1 else 0.0
        print(n, b, "band leak", band_ok, "eig err", np.abs(lam - np.linalg.eigvalsh(A)).max(),
              "res", np.abs(A @ U - U * lam).max(), "orth", np.abs(U.T @ U - np.eye(n)).max())


# ---------------------------------------------------------------------------------------------------------------------
# Systolic formulation of the chase = the data flow of k_chase (csrc/band.cu): position t owns the 32 x 64 window
# [E | D] of rows p .. p+31 (p = s + 1 + 32 t at sweep s): E = columns p-32 .. p-1 (the bulge block), D = the diagonal
# block.  Per hop a position receives the previous reflector of the sweep from position t-1 and, between sweeps, the
# row that enters its window from below from position t+1; its top row leaves to position t-1.  Everything beyond
# the matrix is zero padding, so every hop has length 32.
# ---------------------------------------------------------------------------------------------------------------------
def sb2st_systolic(AB, n, b=32):
    """AB: lower band storage with b+1 diagonals (AB[i-j, j]).  Returns d, e, V2 (n-2 x ldv, ldv = b*NP), tau2 (n-2 x NP)."""
    NP = max(1, -(-(n - 1) // b))

    def a(i, j):
        if i >= n or j >= n or i < 0 or j < 0:
            return 0.0
        lo, hi = (j, i) if i >= j else (i, j)
        return AB[hi - lo, lo] if hi - lo <= b else 0.0

    E = np.zeros((NP, b, b))
    D = np.zeros((NP, b, b))
    for t in range(NP):
        p = 1 + b * t
        for r in range(b):
            for k in range(b):
                D[t, r, k] = a(p + r, p + k)
                if t >= 1:
                    E[t, r, k] = a(p + r, p - b + k)
        if t == 0:
            for r in range(b):
                E[0, r, b - 1] = a(p + r, 0)          # position 0: only the column being annihilated
    nsweeps = np.array([min(n - 2, n - 1 - b * t) for t in range(NP)])
    row_in = np.zeros((NP, b + 1))                     # entering row (from position t+1), consumed at the next sweep
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    d[0] = AB[0, 0]
    V2 = np.zeros((max(n - 2, 0), b * NP))
    tau2 = np.zeros((max(n - 2, 0), NP))
    for s in range(n - 2):
        vmsg = None
        for t in range(NP):
            if s >= nsweeps[t]:
                break
            Et, Dt = E[t], D[t]
            if s > 0:                                  # assemble the window of this sweep: shift by (1,1) + entering row
                En, Dn = np.zeros((b, b)), np.zeros((b, b))
                En[:b - 1, :b - 1] = Et[1:, 1:]
                En[:b - 1, b - 1] = Dt[0, 1:]          # new last column = old D[1:, 0] (symmetry: D[0, 1:])
                Dn[:b - 1, :b - 1] = Dt[1:, 1:]
                m = row_in[t]
                En[b - 1, b - 1] = m[0]
                Dn[b - 1, :b - 1] = m[1:b]
                Dn[:b - 1, b - 1] = m[1:b]
                Dn[b - 1, b - 1] = m[b]
                Et, Dt = En, Dn
            if t == 0:
                v, tau, beta = house(Et[:, b - 1])
                e[s] = beta
            else:
                vp, taup = vmsg
                Et = Et - taup * np.outer(Et @ vp, vp)            # (a) right-apply the previous reflector of the sweep
                v, tau, beta = house(Et[:, 0])                    # (b)
                cdot = v @ Et                                     # (c)
                Et = Et - tau * np.outer(v, cdot)
                Et[:, 0] = 0.0
                Et[0, 0] = beta
            w = tau * (Dt @ v)                                    # (d)
            w -= 0.5 * tau * (w @ v) * v
            Dt = Dt - np.outer(v, w) - np.outer(w, v)
            V2[s, b * t:b * t + b] = v
            tau2[s, t] = tau
            vmsg = (v, tau)
            if t == 0:
                d[s + 1] = Dt[0, 0]
            else:
                row_in[t - 1, :b] = Et[0, :]                      # leaving top row -> position t-1
                row_in[t - 1, b] = Dt[0, 0]
            E[t], D[t] = Et, Dt
            if t + 1 < NP and s + 1 >= nsweeps[t + 1] and s + 1 < nsweeps[t]:
                row_in[t] = 0.0                                   # nothing below any more: zero padding enters
    if n >= 2:                                                    # what is left in position 0 after the last sweep
        if n == 2:
            d[1] = a(1, 1)
            e[0] = a(1, 0)
        else:
            d[n - 1] = D[0][1, 1]
            e[n - 2] = D[0][1, 0]
    return d, e, V2, tau2


def apply_q2_systolic(V2, tau2, n, X, b=32, W=4):
    """X <- Q2 X as k_q2_stage does it: slot q holds rows p .. p+31 (p = s+1+32q) of every column; sweeps run from
    the last one down to 0, the window slides up by one row per sweep: the row entering at the top comes from slot
    q-1 (from X itself for slot 0), the bottom row leaves to slot q+1.  Stages of W slots are separate passes coupled
    by a stream (one row per sweep)."""
    X = X.copy()
    N = X.shape[1]
    NP = V2.shape[1] // b
    nst = -(-NP // W)
    s_hi = n - 2                                       # one no-op sweep first: rows n-1 enter the pipeline
    stream_in = None
    for k in range(nst):
        q0 = k * W
        nq = min(W, NP - q0)
        win = np.zeros((nq, b, N))
        bottom = np.zeros((nq, N))                     # row that leaves slot q at the next shift
        stream_out = np.zeros((s_hi + 1, N))
        for s in range(s_hi, -1, -1):
            newbottom = np.zeros((nq, N))
            for qi in range(nq):
                q = q0 + qi
                p = s + 1 + b * q
                if qi == 0:
                    ent = (X[p] if p < n else np.zeros(N)) if k == 0 else stream_in[s]
                else:
                    ent = bottom[qi - 1]
                leaving = win[qi, b - 1].copy()
                win[qi, 1:] = win[qi, :b - 1].copy()
                win[qi, 0] = ent
                if qi == nq - 1:
                    stream_out[s] = leaving
                else:
                    pass
                if s <= n - 3:
                    v, tau = V2[s, b * q:b * q + b], tau2[s, q]
                    win[qi] -= tau * np.outer(v, v @ win[qi])
                newbottom[qi] = win[qi, b - 1]
            # hand-off of this sweep's leaving rows happens at the NEXT shift: slot q+1 takes bottom[q] of the previous step
            bottom = np.array([win[qi, b - 1] for qi in range(nq)])
        for qi in range(nq):
            p = 1 + b * (q0 + qi)
            hi = min(n, p + b)
            if p < n:
                X[p:hi] = win[qi, :hi - p]
        stream_in = stream_out
    return X


def check_systolic(n, rng, b=32):
    A = rng.standard_normal((n, n))
    A = A + A.T
    Bd = np.triu(np.tril(A, b), -b)                    # any symmetric band matrix
    AB = np.zeros((b + 1, n))
    for j in range(n):
        hi = min(n, j + b + 1)
        AB[:hi - j, j] = Bd[j:hi, j]
    d, e, V2, tau2 = sb2st_systolic(AB, n, b)
    AB2 = to_band(Bd, b)
    d0, e0, V20, tau20 = sb2st(AB2, b)
    err_t = max(np.abs(d - d0).max(), np.abs(e - e0).max() if n > 1 else 0.0)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, Ev = np.linalg.eigh(T)
    U = apply_q2_systolic(V2, tau2, n, Ev, b)
    return err_t, np.abs(Bd @ U - U * lam).max(), np.abs(U.T @ U - np.eye(n)).max()
