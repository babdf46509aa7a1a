"""Per-contraction timing of one ctmrgstep's einsum chain on its real shapes (device-resident operands).
usage: contract_bench.py [chi] [D] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n = chi * D
ctx = T.Context(0)
peak = ctx.dmma_peak()
print(f"DMMA issue peak {peak:.2f} TFLOP/s")
ctx.set_pointer_mode(1)
only = os.environ.get("CB_ONLY")
specs = [  # drivers.cu: ctmrg_step (ctmrg.jl:130-140), label extents from the tensors
    ("iba,ad->ibd", (chi, D, chi), (chi, chi)),
    ("ibd,dcl->ibcl", (chi, D, chi), (chi, D, chi)),
    ("ibcl,jkcb->ijlk", (chi, D, D, chi), (D, D, D, D)),
    ("pq,qj->pj", (n, n), (n, chi)),
    ("pi,pj->ij", (n, chi), (n, chi)),
    ("abi,aed->ibed", (chi, D, chi), (chi, D, chi)),
    ("ibed,bjce->ijcd", (chi, D, D, chi), (D, D, D, D)),
    ("ijcd,dck->ijk", (chi, D, D, chi), (chi, D, chi)),
]
rng = np.random.default_rng(0)
tot_ms = tot_fl = 0.0
for spec, da, db in specs:
    if only and only not in spec: continue
    la, rest = spec.split(","); lb, lc = rest.split("->")
    ext = {}
    for l, d in zip(la, da): ext[l] = d
    for l, d in zip(lb, db): ext[l] = d
    dc = tuple(ext[l] for l in lc)
    fl = 2.0 * np.prod([ext[l] for l in set(la + lb)])
    a = rng.standard_normal(int(np.prod(da))); b = rng.standard_normal(int(np.prod(db)))
    pa, pb, pc = ctx.dev_alloc(a.size), ctx.dev_alloc(b.size), ctx.dev_alloc(int(np.prod(dc)))
    ctx.dev_upload(pa, a); ctx.dev_upload(pb, b)
    for _ in range(3): ctx.dev_contract(spec, pa, da, pb, db, pc)
    # one C call enqueues the product `reps` times (TNAD_CONTRACT_REPS) and synchronises once: device-bound timing
    ctx.set_option("TNAD_CONTRACT_REPS", str(reps))
    ctx.timer_start()
    ctx.dev_contract(spec, pa, da, pb, db, pc)
    ms = ctx.timer_stop() / reps
    ctx.set_option("TNAD_CONTRACT_REPS", "1")
    c = ctx.dev_download(pc, dc)
    ref = np.einsum(spec, a.reshape(da, order="F"), b.reshape(db, order="F"), optimize=True)
    err = np.abs(c - ref).max() / np.abs(ref).max()
    plan = T.contract_plan(spec, da, db)
    print(f"{spec:18s} M={plan['M']:6d} N={plan['N']:5d} K={plan['K']:5d} {ms*1e3:8.1f} us {fl/ms/1e9:6.2f} TFLOP/s ({fl/ms/1e9/peak*100:4.1f} %)  err {err:.1e}", flush=True)
    tot_ms += ms; tot_fl += fl
    for p in (pa, pb, pc): ctx.dev_free(p)
print(f"step forward chain: {tot_fl/1e9:.2f} GFLOP in {tot_ms*1e3:.0f} us = {tot_fl/tot_ms/1e9:.2f} TFLOP/s ({tot_fl/tot_ms/1e9/peak*100:.1f} % of DMMA peak)")
