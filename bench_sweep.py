#!/usr/bin/env python
"""Ising beta-sweep of independent TRG instances fanned out over GPUs (BASELINE configs[4], second half;
SURVEY.md section 8e): instance i -> rank i mod N, no data-path collective, one all-reduce of the result table.

    python bench_sweep.py [--instances 64] [--chi 20] [--niter 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_sweep.py --gpus N

Prints one JSON line on rank 0: instances/s (value + gradient per instance), max-over-ranks device time.
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


def measure(dist, ctx, instances=64, chi=20, niter=20):
    """64-instance Ising beta sweep (value + d/dbeta per instance) round-robin over dist.world GPUs."""
    from tnad_b200.sweep import trg_beta_sweep, gather_results
    betas = np.linspace(0.30, 0.60, instances)
    trg_beta_sweep(betas[:1], chi, 2, 0, 1, ctx=ctx)          # warm-up (module load, pool growth)
    dist.barrier()
    ctx.timer_start()
    local = trg_beta_sweep(betas, chi, niter, dist.rank, dist.world, ctx=ctx)
    ms = dist.max(ctx.timer_stop())
    table = gather_results(local, instances, dist.dist if dist.on else None)
    i = int(np.argmin(np.abs(betas - 0.5)))
    return {"instances_per_s": instances / (ms * 1e-3), "ms_total": ms,
            "workload": f"{instances} TRG instances (value + d/dbeta), Ising beta in [0.30, 0.60], chi={chi}, niter={niter}",
            "parallelism": f"instances round-robin over {dist.world} GPUs, no collective",
            "sample": {"beta": float(betas[i]), "lnZ": float(table[i, 0]), "dlnZ_dbeta": float(table[i, 1])},
            "all_finite": bool(np.all(np.isfinite(table)))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--instances", type=int, default=64)
    ap.add_argument("--chi", type=int, default=20)
    ap.add_argument("--niter", type=int, default=20)
    args = ap.parse_args()
    import tnad_b200 as T
    dist = Dist()
    ctx = T.Context(dist.local_rank)
    r = measure(dist, ctx, args.instances, args.chi, args.niter)
    if dist.rank == 0:
        print(json.dumps({
            "metric": "trg_beta_sweep_instances_per_s", "value": r["instances_per_s"], "unit": "instances/s",
            "n_gpus": dist.world, "ms_total": r["ms_total"], "higher_is_better": True, "scaling": "strong", "dtype": "f64",
            "data": "synthetic", "config": {"workload": r["workload"], "parallelism": r["parallelism"]},
            "sample": r["sample"], "all_finite": r["all_finite"],
        }), flush=True)
    dist.barrier()
    ctx.close()
    dist.close()


if __name__ == "__main__":
    main()
