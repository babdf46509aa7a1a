"""Per-call device time of the headline workload (energy + gradient, d = 4, chi = 128, maxit = 10): looks for outliers.
usage: c4_calls.py [ncalls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = T.Context(0)
h = T.hamiltonian(T.Heisenberg())
A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))).bulk
for i in range(n):
    ctx.timer_start()
    e, g = ctx.energy(h, A, 128, 0.0, 10, grad=True)
    ms = ctx.timer_stop()
    t = ctx.last_timing()
    print(f"call {i}: {ms:8.1f} ms  " + "  ".join(f"{k} {v:7.1f}" for k, v in t.items()), flush=True)
