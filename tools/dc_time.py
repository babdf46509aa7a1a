"""Device-resident timing of the symmetric SVD routes on a CTMRG corner matrix (development helper)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tnad_b200 as T
ctx = T.Context(0)
d, chi = int(sys.argv[1]), int(sys.argv[2])
from tnad_b200.sharded import ShardedCTMRG
D = d * d
rng = np.random.default_rng(17)
a = rng.standard_normal((d,) * 4 + (2,))
ipeps = T.indexperm_symmetrize(T.SquareIPEPS(a))
bulk = np.einsum("abcdx,ijklx->aibjckdl", ipeps.bulk, ipeps.bulk).reshape((D, D, D, D), order="F")
bulk /= np.linalg.norm(bulk)
corner = rng.standard_normal((chi, chi)); corner += corner.T
edge = rng.standard_normal((chi, D, chi)); edge += edge.transpose(2, 1, 0)
sh = ShardedCTMRG(ctx, chi, D)
sh.load(bulk, corner, edge)
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    sh.step(timing=True)
    print(f"step {it}: n={chi * D} svd {sh.ms['svd']:.2f} ms contract {sh.ms['contract']:.2f} ms", flush=True)
    sh.advance()
