"""Measurement helper: TRG steady-state iterations/s at bond dimension chi (BASELINE metric, third part).
usage: trg_chi64.py chi niter [grad]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnad_b200 as T
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 64
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
grad = len(sys.argv) > 3 and sys.argv[3] == "1"
ctx = T.Context(0)
a = T.model_tensor(T.Ising(), 0.44)
# steady state starts once the bond dimension has saturated (iteration ~ log2(chi)+2): time niter and niter+2
def run(n):
    ctx.timer_start()
    if grad:
        lnz, g = T.trg_value_and_grad(a, chi, n, ctx=ctx)
    else:
        lnz = T.trg(a, chi, n, ctx=ctx)
    return ctx.timer_stop(), lnz
run(2)
# every distinct iteration count grows the stream-ordered pool on its first run (first-touch allocation is slow):
# run each count twice and keep the second timing
run(niter + 2)
t2, l2 = run(niter + 2)
run(niter)
t1, l1 = run(niter)
per_iter = (t2 - t1) / 2.0
print(f"TRG chi={chi} grad={grad}: {niter} it {t1:.1f} ms, {niter+2} it {t2:.1f} ms -> steady state {per_iter:.1f} ms/iteration = {1e3/per_iter:.3f} it/s; lnZ={l2!r}", flush=True)
