"""Stage timings of the symmetric eigensolver (TNAD_DC_DEBUG=1 prints them from the library).
usage: ts_time.py n [two_stage 0|1] [reps]"""
import os, sys
os.environ.setdefault("TNAD_DC_DEBUG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ts = sys.argv[2] if len(sys.argv) > 2 else "1"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = T.Context(0)
ctx.set_option("TNAD_SYMEIG", "2")
ctx.set_option("TNAD_EIG_2STAGE", ts)
a = np.random.default_rng(0).standard_normal((n, n)); a = a + a.T
for r in range(reps):
    ctx.timer_start()
    u, s, v = ctx.svd_sym(a)
    ms = ctx.timer_stop()
    print(f"n={n} two_stage={ts} rep {r}: {ms:.2f} ms (incl. H2D/D2H of the matrices)", flush=True)
ref = np.linalg.svd(a, compute_uv=False)
print("max |s - ref| / s0 =", np.abs(s - ref).max() / ref[0], " recon", np.abs((u * s) @ v.T - a).max() / ref[0],
      " orth", np.abs(u.T @ u - np.eye(n)).max())
