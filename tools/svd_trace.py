"""Bring-up helper: one symmetric SVD of size n with the per-sweep convergence trace.
usage: svd_trace.py n reps [sym|gen]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnad_b200 as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = sys.argv[3] if len(sys.argv) > 3 else "gen"
ctx = T.Context(0)
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n)); A = A + A.T
for r in range(reps):
    t0 = time.time()
    U, S, V = ctx.svd_sym(A) if mode == "sym" else ctx.svd(A)
    dt = time.time() - t0
    print(f"{mode} n={n} wall {dt:.3f}s sweeps {ctx.last_sweeps} rec {np.linalg.norm((U*S)@V.T-A)/np.linalg.norm(A):.2e} "
          f"orthU {np.abs(U.T@U-np.eye(n)).max():.1e} orthV {np.abs(V.T@V-np.eye(n)).max():.1e} "
          f"dS {np.abs(S-np.linalg.svd(A,compute_uv=False)).max()/S[0]:.1e}", flush=True)
