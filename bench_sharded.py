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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--d", type=int, default=5)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import tnad_b200 as T
    from tnad_b200.sharded import ShardedCTMRG
    dist = Dist()
    torch.cuda.set_device(dist.local_rank)
    ctx = T.Context(dist.local_rank)
    D, chi = args.d * args.d, args.chi
    rng = np.random.default_rng(17)                       # same inputs on every rank
    a = rng.standard_normal((args.d,) * 4 + (2,))
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(a))
    bulk = np.einsum("abcdx,ijklx->aibjckdl", ipeps.bulk, ipeps.bulk).reshape((D, D, D, D), order="F")
    bulk /= np.linalg.norm(bulk)
    corner = rng.standard_normal((chi, chi)); corner += corner.T
    edge = rng.standard_normal((chi, D, chi)); edge += edge.transpose(2, 1, 0)
    sh = ShardedCTMRG(ctx, chi, D, dist.dist if dist.on else None)
    sh.load(bulk, corner, edge)
    for _ in range(args.warmup):
        sh.step()
        sh.advance()
    agg = {"contract": 0.0, "gather": 0.0, "svd": 0.0}
    sweeps = []
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sh.step(timing=True)
        sh.advance()
        for k in agg:
            agg[k] += sh.ms[k]
        sweeps.append(sh.sweeps)
    e1.record()
    torch.cuda.synchronize()
    ms = dist.max(e0.elapsed_time(e1)) / args.steps
    parts = {k: dist.max(v) / args.steps for k, v in agg.items()}
    diff = None
    if args.check:
        sh.load(bulk, corner, edge)
        sh.step()
        cg, eg, vg = sh.result()
        c1, e1_, v1 = ctx.ctmrgstep(bulk, corner, edge)
        diff = dist.max(float(max(np.abs(vg - v1).max(), np.abs(np.abs(cg) - np.abs(c1)).max(),
                                  np.abs(np.abs(eg) - np.abs(e1_)).max())))
    if dist.rank == 0:
        n = chi * D
        print(json.dumps({
            "metric": "sharded_ctmrgstep_seconds", "value": ms * 1e-3, "unit": "s/step", "n_gpus": dist.world,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": False, "scaling": "strong", "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"ctmrgstep forward, d={args.d} (D={D}), chi={chi}, n={n}",
                                            "parallelism": f"chi-sharded contractions over {dist.world} GPUs, NCCL all-gather, replicated SVD"},
            "ms_contract": parts["contract"], "ms_gather": parts["gather"], "ms_svd": parts["svd"], "svd_sweeps": sweeps,
            "gather_bytes_per_step": 8 * (n * n + chi * chi + chi * D * chi),
            "check_max_abs_diff_vs_unsharded": diff,
        }), flush=True)
    dist.barrier()
    ctx.close()
    dist.close()


if __name__ == "__main__":
    main()
