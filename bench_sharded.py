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


def measure(dist, ctx, d=5, chi=256, steps=3, warmup=1, check=False):
    """chi-sharded forward ctmrgstep on dist.world GPUs: max-over-ranks device time per step and its split."""
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
    ap.add_argument("--d", type=int, default=5)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import tnad_b200 as T
    dist = Dist()
    ctx = T.Context(dist.local_rank)
    r = measure(dist, ctx, args.d, args.chi, args.steps, args.warmup, args.check)
    if dist.rank == 0:
        print(json.dumps({
            "metric": "sharded_ctmrgstep_seconds", "value": r["s_per_step"], "unit": "s/step", "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": False, "scaling": "strong", "dtype": "f64",
            "data": "synthetic", "config": {"workload": r["workload"], "parallelism": r["parallelism"]},
            "ms_contract": r["ms_contract"], "ms_gather": r["ms_gather"], "ms_svd": r["ms_svd"],
            "gather_bytes_per_step": r["gather_bytes_per_step"],
            "check_max_abs_diff_vs_unsharded": r["check_max_abs_diff_vs_unsharded"],
        }), flush=True)
    dist.barrier()
    ctx.close()
    dist.close()


if __name__ == "__main__":
    main()
