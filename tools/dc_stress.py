"""Randomised stress of the direct symmetric eigensolver and the Jordan-Wielandt SVD route (development helper):
many sizes (odd, non multiples of 4/32, around the cluster-tail limit 416), spectra (random, low rank, graded,
clustered, +- pairs), checked against numpy."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnad_b200 as T
os.environ["TNAD_SYMEIG"] = "2"
ctx = T.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
sizes = [3, 4, 5, 7, 31, 32, 33, 63, 64, 65, 95, 97, 127, 129, 255, 257, 383, 415, 416, 417, 418, 420, 447, 448, 449, 450, 511, 513, 640, 641, 777, 1000, 1023, 1025, 1500]
worst = 0.0
t0 = time.time()
for n in sizes:
    for kind in ("rand", "lowrank", "graded", "cluster", "pm"):
        a = rng.standard_normal((n, n))
        if kind == "rand":
            m = a + a.T
        elif kind == "lowrank":
            b = rng.standard_normal((n, max(1, n // 7))); m = b @ b.T
        elif kind == "graded":
            d = 10.0 ** (-rng.uniform(0, 12, n)); m = d[:, None] * (a + a.T) * d[None, :]
        elif kind == "cluster":
            q, _ = np.linalg.qr(a); w = np.round(rng.uniform(-2, 2, n)); m = (q * w) @ q.T; m = 0.5 * (m + m.T)
        else:
            k = n // 2; b = rng.standard_normal((k, n - k)); m = np.zeros((n, n)); m[:k, k:] = b; m[k:, :k] = b.T
        u, s, v = ctx.svd_sym(m)
        ref = np.linalg.svd(m, compute_uv=False)
        sc = max(ref[0], 1e-300)
        err = max(np.abs(s - ref).max() / sc, np.abs((u * s) @ v.T - m).max() / sc, np.abs(u.T @ u - np.eye(n)).max())
        worst = max(worst, err)
        if not (err < 5e-12) or not np.all(np.isfinite(u)):
            print(f"FAIL n={n} kind={kind}: err {err:.2e}", flush=True)
print(f"symmetric: {len(sizes) * 5} matrices, worst error {worst:.2e}, {time.time() - t0:.1f} s", flush=True)
# TRG route: general rectangular / rank-deficient matrices through tnad_trg_svd
worst = 0.0
for (d1, d2, d3, d4) in [(8, 8, 8, 8), (10, 7, 5, 9), (12, 12, 12, 12), (20, 20, 20, 20), (16, 9, 9, 16)]:
    for rank in (None, 11):
        m, n = d1 * d2, d3 * d4
        if rank is None:
            t4 = rng.standard_normal((d1, d2, d3, d4))
        else:
            t4 = (rng.standard_normal((m, rank)) @ rng.standard_normal((rank, n))).reshape((d1, d2, d3, d4), order="F")
        k = 9
        us, vs = ctx.trg_svd(t4, k, 1e-12)[:2]
        M = t4.reshape((m, n), order="F")
        U, S, Vt = np.linalg.svd(M)
        kk = us.shape[-1]
        approx = us.reshape((m, kk), order="F") @ vs.reshape((kk, n), order="F")
        best = (U[:, :kk] * S[:kk]) @ Vt[:kk]
        err = np.abs(approx - best).max() / S[0]
        worst = max(worst, err)
        if not err < 1e-10:
            print(f"FAIL trg_svd {d1,d2,d3,d4} rank={rank}: {err:.2e} (k={kk})", flush=True)
print(f"trg_svd: worst truncation mismatch {worst:.2e}", flush=True)
