"""Independent-instance fan-out: beta-sweeps and parameter scans shard by instance,
one process per GPU, with no data-path collective; only the scalar results are gathered at the end."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def partition(n_instances: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of instance indices to ranks (instance i -> rank i mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    return list(range(rank, n_instances, world))


def trg_beta_sweep(betas: Sequence[float], chi: int, niter: int, rank: int = 0, world: int = 1, ctx=None,
                   tol: float = 1e-16) -> Dict[int, Tuple[float, float]]:
    """lnZ and d lnZ / d beta (README.md:59-70) for this rank's share of an Ising beta-sweep."""
    from . import Ising, model_tensor, dmodel_tensor, trg_value_and_grad
    out = {}
    for i in partition(len(betas), rank, world):
        b = float(betas[i])
        lnz, g = trg_value_and_grad(model_tensor(Ising(), b), chi, niter, tol=tol, ctx=ctx)
        out[i] = (lnz, float(np.sum(g * dmodel_tensor(Ising(), b))))
    return out


def gather_results(local: Dict[int, Tuple[float, float]], n_instances: int, dist=None) -> np.ndarray:
    """All ranks' (lnZ, dlnZ/dbeta) rows in instance order; `dist` is torch.distributed (or None at world 1)."""
    table = np.full((n_instances, 2), np.nan)
    for i, v in local.items():
        table[i] = v
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return table
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.nan_to_num(table, nan=0.0)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)       # instances are disjoint across ranks
    return t.cpu().numpy()
