"""One energy + gradient call at a given size (profiling target). usage: c4_trace.py d chi maxit [grad 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
d = int(sys.argv[1]) if len(sys.argv) > 1 else 4
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 1
grad = (sys.argv[4] == "1") if len(sys.argv) > 4 else True
ctx = T.Context(0)
h = T.hamiltonian(T.Heisenberg())
A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((d, d, d, d, 2)))).bulk
out = ctx.energy(h, A, chi, 0.0, maxit, grad=grad)
print("energy", out[0] if grad else out, "steps", ctx.last_steps)
