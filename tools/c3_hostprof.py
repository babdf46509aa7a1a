import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, tnad_b200 as T, tnad_oracle as O
ctx = T.Context(0)
h = T.hamiltonian(T.Heisenberg())
A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))
for _ in range(3): ctx.energy(h, A, 20, 1e-6, 100, grad=True)
t0 = time.perf_counter()
for _ in range(10): ctx.energy(h, A, 20, 1e-6, 100, grad=True)
print("wall ms per call", (time.perf_counter() - t0) * 100, "launches per call", ctx.launch_count() / 13)
ctx.close()
