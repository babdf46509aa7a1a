"""Bring-up helper: C4-shaped energy(+grad) call with timing breakdown."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T
import tnad_oracle as O
d = int(sys.argv[1]) if len(sys.argv) > 1 else 4
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 5
grad = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
ctx = T.Context(0)
h = O.hamiltonian_heisenberg()
A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((d, d, d, d, 2)))
for rep in range(2):
    ctx.reset_launch_count()
    t0 = time.time()
    out = ctx.energy(h, A, chi, 0.0, maxit, grad=grad)
    dt = time.time() - t0
    e = out[0] if grad else out
    print(f"d={d} chi={chi} maxit={maxit} grad={grad}: E={e!r} steps={ctx.last_steps} wall={dt:.3f}s "
          f"launches={ctx.launch_count()} timing={ctx.last_timing()}", flush=True)
