import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, tnad_b200 as T, bench as B
mode = sys.argv[1]
ctx = T.Context(0)
h = B.heisenberg_h(); A = B.ipeps_tensor(0)
hp, Ap, gp = ctx.dev_alloc(h.size), ctx.dev_alloc(A.size), ctx.dev_alloc(A.size)
ctx.dev_upload(hp, h); ctx.dev_upload(Ap, A)
out = []
if mode == "outer": ctx.timer_start()
for i in range(20):
    if mode == "inner": ctx.timer_start()
    ctx.energy_device(hp, Ap, B.D_IPEPS, B.S_PHYS, B.CHI, 0.0, 10, gp)
    if mode == "inner": ctx.timer_stop()
    t = ctx.last_timing(); out.append(f"{t['svd']:.0f}/{t['backward']:.0f}")
print(mode, " ".join(out))
