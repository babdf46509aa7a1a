"""GPU probe for the direct eigensolver stages (development helper, run through gpurun):
tnad_stedc vs numpy on tridiagonals, tnad_sytrd vs the input matrix, then the full SVD route on CTMRG-like input."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T

ctx = T.Context(0)
rng = np.random.default_rng(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "stedc"):
    for n in (5, 40, 64, 65, 130, 500, 1000, 2048):
        d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
        Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ctx.timer_start(); lam, Z = ctx.stedc(d, e); ms = ctx.timer_stop()
        ref = np.linalg.eigvalsh(Tm); nrm = np.abs(ref).max()
        print(f"stedc n={n}: eig {np.abs(lam - ref).max() / nrm:.2e} resid {np.abs(Tm @ Z - Z * lam).max() / nrm:.2e} "
              f"orth {np.abs(Z.T @ Z - np.eye(n)).max():.2e}  {ms:.2f} ms", flush=True)
    n = 600   # graded / clustered
    d = 10.0 ** (-rng.uniform(0, 18, n)); e = 10.0 ** (-rng.uniform(0, 18, n - 1))
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, Z = ctx.stedc(d, e); ref = np.linalg.eigvalsh(Tm)
    print(f"stedc graded: eig {np.abs(lam - ref).max():.2e} resid {np.abs(Tm @ Z - Z * lam).max():.2e} orth {np.abs(Z.T @ Z - np.eye(n)).max():.2e}", flush=True)

if which in ("all", "sytrd"):
    for n in (3, 10, 33, 64, 100, 257, 1024, 2048):
        a = rng.standard_normal((n, n)); a = a + a.T
        ctx.timer_start(); d, e, q = ctx.sytrd(a); ms = ctx.timer_stop()
        Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        nrm = np.abs(a).max() * n ** 0.5
        print(f"sytrd n={n}: |A - QTQ'| {np.abs(q @ Tm @ q.T - a).max() / nrm:.2e} orth {np.abs(q.T @ q - np.eye(n)).max():.2e}  {ms:.2f} ms (incl. explicit Q + copies)", flush=True)

if which in ("all", "svd"):
    for n in (300, 1024, 2048):
        a = rng.standard_normal((n, n)) * (10.0 ** (-rng.uniform(0, 12, n)))[None, :]
        a = a + a.T
        for mode in ("1", "2"):
            os.environ["TNAD_SYMEIG"] = mode
            ctx.svd_sym(a)
            ctx.timer_start(); u, s, v = ctx.svd_sym(a)[:3]; ms = ctx.timer_stop()
            nrm = s[0]
            print(f"svd_sym mode={mode} n={n}: recon {np.abs((u * s) @ v.T - a).max() / nrm:.2e} orthU {np.abs(u.T @ u - np.eye(n)).max():.2e} "
                  f"s-err {np.abs(s - np.linalg.svd(a, compute_uv=False)).max() / nrm:.2e}  {ms:.1f} ms (with h2d/d2h)", flush=True)
