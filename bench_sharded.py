#!/usr/bin/env python
"""chi-sharded ctmrgstep on N GPUs (BASELINE configs[4], first half: d=5 -> D=25, chi=256, n=6400).

    python bench_sharded.py [--d 5 --chi 256 --steps 3]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_sharded.py --gpus N [--check]

Each rank computes a chi/N slice of every contraction, NCCL all-gathers the slices (tensornetworkad.jl_b200/sharded.py);
the n x n eigen-decomposition is replicated.  Prints one JSON line on rank 0 with the max-over-ranks time per step
and its split into contraction / all-gather / SVD time.  --check also runs the unsharded C-ABI step on rank 0's
inputs and reports the largest gauge-invariant difference (every rank must agree with it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bench import Dist  # noqa: E402


def comm_setup(dist, ctx):
    """Communicator of the library's own sharded step: rank 0 asks NCCL for a unique id, torch.distributed carries the
    128 bytes to the other ranks, every rank joins with its context."""
    if getattr(ctx, "_comm_ready", False):
        return
    if dist.on:
        import torch
        ident = torch.zeros(128, dtype=torch.uint8, device=dist.dev)
        if dist.rank == 0:
            ident.copy_(torch.frombuffer(bytearray(ctx.nccl_unique_id()), dtype=torch.uint8))
        dist.dist.broadcast(ident, src=0)
        ctx.comm_init(bytes(ident.cpu().numpy().tobytes()), dist.rank, dist.world)
    else:
        ctx.comm_init(None, 0, 1)
    ctx._comm_ready = True


def measure_library(dist, ctx, bulk, corner, edge, steps, warmup, check):
    """tnad_ctmrgstep_sharded: the whole step (sliced contractions, ncclAllGather, shared back-transformation) inside
    the library, device-resident environment, one C call per step."""
    import torch
    chi, D = corner.shape[0], bulk.shape[0]
    comm_setup(dist, ctx)
    dev = torch.device("cuda", ctx.device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))).to(dev)
    tb, tc, te = up(bulk), up(corner), up(edge)
    tco, teo = torch.empty_like(tc), torch.empty_like(te)
    torch.cuda.synchronize(dev)
    agg = [0.0, 0.0, 0.0]
    ctx.set_pointer_mode(1)
    try:
        for _ in range(warmup):
            ctx.dev_ctmrgstep_sharded(tb.data_ptr(), D, tc.data_ptr(), te.data_ptr(), chi, tco.data_ptr(), teo.data_ptr(), timing=False)
            tc.copy_(tco); te.copy_(teo)
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vals = None
        for _ in range(steps):
            vals, ms3 = ctx.dev_ctmrgstep_sharded(tb.data_ptr(), D, tc.data_ptr(), te.data_ptr(), chi, tco.data_ptr(), teo.data_ptr())
            tc.copy_(tco); te.copy_(teo)
            agg = [a + b for a, b in zip(agg, ms3)]
        e1.record()
        torch.cuda.synchronize()
        ms = dist.max(e0.elapsed_time(e1)) / steps
        parts = [dist.max(v) / steps for v in agg]
        diff = None
        if check:
            tc.copy_(up(corner)); te.copy_(up(edge))
            vg, _ = ctx.dev_ctmrgstep_sharded(tb.data_ptr(), D, tc.data_ptr(), te.data_ptr(), chi, tco.data_ptr(), teo.data_ptr())
            cg = tco.cpu().numpy().reshape((chi, chi), order="F")
            eg = teo.cpu().numpy().reshape((chi, D, chi), order="F")
    finally:
        ctx.set_pointer_mode(0)
    if check:
        c1, e1_, v1 = ctx.ctmrgstep(bulk, corner, edge)
        diff = dist.max(float(max(np.abs(vg - v1).max(), np.abs(cg - c1).max(), np.abs(eg - e1_).max())))
    return ms, {"contract": parts[0], "gather": parts[1], "svd": parts[2]}, diff


def measure(dist, ctx, d=5, chi=256, steps=3, warmup=1, check=False, impl="library"):
    """chi-sharded forward ctmrgstep on dist.world GPUs: max-over-ranks device time per step and its split.
    impl = "library": tnad_ctmrgstep_sharded (C ABI, NCCL inside the library); "python": the same schedule driven op by
    op from tensornetworkad.jl_b200/sharded.py through torch.distributed."""
    import torch
    import tnad_b200 as T
    from tnad_b200.sharded import ShardedCTMRG
    torch.cuda.set_device(dist.local_rank)
    D = d * d
    rng = np.random.default_rng(17)                       # same inputs on every rank
    a = rng.standard_normal((d,) * 4 + (2,))
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(a))
    bulk = np.einsum("abcdx,ijklx->aibjckdl", ipeps.bulk, ipeps.bulk).reshape((D, D, D, D), order="F")
    bulk /= np.linalg.norm(bulk)
    corner = rng.standard_normal((chi, chi)); corner += corner.T
    edge = rng.standard_normal((chi, D, chi)); edge += edge.transpose(2, 1, 0)
    n = chi * D
    if impl == "library":
        ms, parts, diff = measure_library(dist, ctx, bulk, corner, edge, steps, warmup, check)
        return {"s_per_step": ms * 1e-3, "ms_contract": parts["contract"], "ms_gather": parts["gather"], "ms_svd": parts["svd"],
                "gather_bytes_per_step": 8 * (n * n + chi * chi + chi * D * chi) + (8 * n * n if dist.world > 1 else 0),
                "check_max_abs_diff_vs_unsharded": diff, "workload": f"ctmrgstep forward, d={d} (D={D}), chi={chi}, n={n}",
                "parallelism": f"tnad_ctmrgstep_sharded: chi-sharded contractions over {dist.world} GPUs, ncclAllGather inside the library, "
                               "replicated reduction + shared back-transformation of the eigen-decomposition"}
    sh = ShardedCTMRG(ctx, chi, D, dist.dist if dist.on else None)
    sh.load(bulk, corner, edge)
    for _ in range(warmup):
        sh.step()
        sh.advance()
    agg = {"contract": 0.0, "gather": 0.0, "svd": 0.0}
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sh.step(timing=True)
        sh.advance()
        for k in agg:
            agg[k] += sh.ms[k]
    e1.record()
    torch.cuda.synchronize()
    ms = dist.max(e0.elapsed_time(e1)) / steps
    parts = {k: dist.max(v) / steps for k, v in agg.items()}
    diff = None
    if check:
        sh.load(bulk, corner, edge)
        sh.step()
        cg, eg, vg = sh.result()
        c1, e1_, v1 = ctx.ctmrgstep(bulk, corner, edge)
        diff = dist.max(float(max(np.abs(vg - v1).max(), np.abs(cg - c1).max(), np.abs(eg - e1_).max())))
    n = chi * D
    return {"s_per_step": ms * 1e-3, "ms_contract": parts["contract"], "ms_gather": parts["gather"], "ms_svd": parts["svd"],
            "gather_bytes_per_step": 8 * (n * n + chi * chi + chi * D * chi), "check_max_abs_diff_vs_unsharded": diff,
            "workload": f"ctmrgstep forward, d={d} (D={D}), chi={chi}, n={n}",
            "parallelism": f"chi-sharded contractions over {dist.world} GPUs, NCCL all-gather, replicated eigen-decomposition"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--d", "--phys-d", dest="d", type=int, default=5)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--impl", default="library", choices=["library", "python"])
    ap.add_argument("--check-energy", action="store_true",
                    help="energy + gradient (d=2, chi=32, maxit=4) on the communicator context (sharded forward steps, replicated "
                         "reverse sweep) against a plain context on every rank")
    args = ap.parse_args()
    import tnad_b200 as T
    dist = Dist()
    ctx = T.Context(dist.local_rank)
    r = measure(dist, ctx, args.d, args.chi, args.steps, args.warmup, args.check, args.impl)
    if args.check_energy:
        comm_setup(dist, ctx)
        h = T.hamiltonian(T.Heisenberg())
        A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(3).standard_normal((2, 2, 2, 2, 2)))).bulk
        ctx.set_option("TNAD_SHARDED_LOOP", "1")        # collective from here on: every rank makes the same calls
        e1, g1 = ctx.energy(h, A, 32, 0.0, 4, grad=True)
        ctx.set_option("TNAD_SHARDED_LOOP", None)
        plain = T.Context(dist.local_rank)
        e0, g0 = plain.energy(h, A, 32, 0.0, 4, grad=True)
        plain.close()
        r["check_energy_rel_diff"] = dist.max(abs(e1 - e0) / abs(e0))
        r["check_gradient_rel_diff"] = dist.max(float(np.abs(g1 - g0).max() / np.abs(g0).max()))
    if dist.rank == 0:
        print(json.dumps({
            "metric": "sharded_ctmrgstep_seconds", "value": r["s_per_step"], "unit": "s/step", "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": False, "scaling": "strong", "dtype": "f64",
            "data": "synthetic", "config": {"workload": r["workload"], "parallelism": r["parallelism"]},
            "ms_contract": r["ms_contract"], "ms_gather": r["ms_gather"], "ms_svd": r["ms_svd"],
            "gather_bytes_per_step": r["gather_bytes_per_step"],
            "check_max_abs_diff_vs_unsharded": r["check_max_abs_diff_vs_unsharded"],
            "check_energy_rel_diff": r.get("check_energy_rel_diff"), "check_gradient_rel_diff": r.get("check_gradient_rel_diff"),
        }), flush=True)
    dist.barrier()
    ctx.close()
    dist.close()


if __name__ == "__main__":
    main()
