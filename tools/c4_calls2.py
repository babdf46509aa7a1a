"""Per-call device time of the two bench arms (device-resident inputs / pinned host buffers): looks for outliers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
import bench as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = T.Context(0)
h = B.heisenberg_h(); A = B.ipeps_tensor(0)
hp, Ap, gp = ctx.dev_alloc(h.size), ctx.dev_alloc(A.size), ctx.dev_alloc(A.size)
ctx.dev_upload(hp, h); ctx.dev_upload(Ap, A)
out = []
for i in range(n):
    ctx.timer_start(); ctx.energy_device(hp, Ap, B.D_IPEPS, B.S_PHYS, B.CHI, 0.0, 10, gp); out.append(ctx.timer_stop())
print("device arm :", " ".join(f"{x:.0f}" for x in out))
hh, Ah = ctx.host_alloc(h.shape), ctx.host_alloc(A.shape)
hh[...] = h; Ah[...] = A
out = []
for i in range(n):
    ctx.timer_start(); ctx.energy(hh, Ah, B.CHI, 0.0, 10, grad=True); out.append(ctx.timer_stop())
print("pinned host:", " ".join(f"{x:.0f}" for x in out))
