"""chi-sharded ctmrgstep over N GPUs (BASELINE.json north_star: "at chi >= 256 the enlarged-corner GEMMs are
sharded along the chi index with an NCCL all-gather over NVLink before the SVD").

One process per GPU.  Every rank holds the full environment (corner chi x chi, edge chi x D x chi: a few MB),
computes a 1/N slice of every O(chi^3 D^3) / O(chi^3 D^4) contraction of ctmrg.jl:126-142 and the slices are
concatenated with `all_gather_into_tensor` (NCCL; gloo in the CPU tests).  Slices are always taken on the
SLOWEST (last, column-major) index of the result so the gathered buffer is the full tensor without a
re-layout:

    grow      X1[i,b,d]    = edge[i,b,a] corner[a,d]                       replicated (chi^3 D)
              X2r[i,b,c,l'] = X1[i,b,d] edge[d,c,l']         l' in this rank's chi/N slice of l
              cpk[i,j,k,l'] = X2r[i,b,c,l'] bulk[j,k,c,b]
              all-gather -> cpk[i,j,k,l];  cp = permutedims(cpk,(1,2,4,3))           (tnad_permute)
    svd       U,S,V = svd(cp + cp')                           replicated (tnad_svd_symmetrized; not sharded -
                                                              the Jacobi sweep is latency bound)
    corner    W[:,j'] = CP Z[:,j'];  c1[:,j'] = Z' W[:,j']    j' slice of the chi kept columns, all-gather
    edge      Yr[a,e,c,k'] = edge[a,e,d] z[d,c,k'];  Yb[a,b,j,k'] = Yr[a,e,c,k'] bulk[b,j,c,e]
              e1[i,j,k'] = z[a,b,i] Yb[a,b,j,k']              k' slice, all-gather
    finish    symmetrise + normalise                          replicated (tnad_ctmrg_finish)

All device work goes through the C ABI in TNAD_POINTER_DEVICE mode on buffers owned by torch (torch is the
allocator and the NCCL plumbing only).  There is no CPU fallback: `ShardedCTMRG` needs the CUDA library; the
step is expressed as data (`step_schedule`), which the gloo world-2 CPU test interprets with its own NumPy code.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ["shard_plan", "buffer_sizes", "step_schedule", "ShardedCTMRG"]


@dataclass(frozen=True)
class ShardPlan:
    chi: int
    D: int
    world: int
    rank: int
    width: int      # chi / world
    start: int      # first kept index of this rank

    @property
    def n(self):
        return self.chi * self.D


def shard_plan(chi: int, D: int, world: int, rank: int) -> ShardPlan:
    """Equal slices of the chi index (all_gather_into_tensor needs equal parts)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if chi % world:
        raise ValueError(f"chi={chi} must be divisible by the number of ranks ({world})")
    w = chi // world
    return ShardPlan(chi, D, world, rank, w, rank * w)


def buffer_sizes(p: ShardPlan) -> dict:
    """Flat (column-major) device buffers of one rank, in doubles."""
    chi, D, n, w = p.chi, p.D, p.n, p.width
    return {"bulk": D ** 4, "corner": chi * chi, "edge": chi * D * chi, "X1": chi * D * chi, "X2r": chi * D * D * w,
            "cpk_r": chi * D * D * w, "cpk": n * n, "cp": n * n, "U": n * n, "S": n, "V": n * n, "Wr": n * w,
            "c1_r": chi * w, "c1": chi * chi, "Yr": chi * D * D * w, "Yb": chi * D * D * w, "e1_r": chi * D * w,
            "e1": chi * D * chi, "corner_out": chi * chi, "edge_out": chi * D * chi}


def step_schedule(p: ShardPlan) -> list:
    """The sharded step as data: a list of operations on named flat buffers.

    ("contract", spec, (A, offset, dims), (B, offset, dims), C) | ("gather", full, part) |
    ("permute", src, dims, perm, dst) | ("svd_symmetrized", A, n, U, S, V) | ("finish", c1, e1, corner_out, edge_out)
    `ShardedCTMRG.step` executes it through the C ABI; tests/test_host_logic.py interprets the same list with
    NumPy under gloo (world_size 2) to check the slicing against the oracle."""
    chi, D, n, w, k0 = p.chi, p.D, p.n, p.width, p.start
    return [
        ("contract", "iba,ad->ibd", ("edge", 0, (chi, D, chi)), ("corner", 0, (chi, chi)), "X1"),
        ("contract", "ibd,dcl->ibcl", ("X1", 0, (chi, D, chi)), ("edge", k0 * chi * D, (chi, D, w)), "X2r"),
        ("contract", "ibcl,jkcb->ijkl", ("X2r", 0, (chi, D, D, w)), ("bulk", 0, (D, D, D, D)), "cpk_r"),
        ("gather", "cpk", "cpk_r"),
        ("permute", "cpk", (chi, D, D, chi), (0, 1, 3, 2), "cp"),
        ("svd_symmetrized", "cp", n, "U", "S", "V"),
        # corner = Z' CP Z on this rank's columns of Z = U[:, :chi]
        ("contract", "pq,qj->pj", ("cp", 0, (n, n)), ("U", k0 * n, (n, w)), "Wr"),
        ("contract", "pi,pj->ij", ("U", 0, (n, chi)), ("Wr", 0, (n, w)), "c1_r"),
        # edge = z' (edge * bulk) z on this rank's slice of the right z
        ("contract", "aed,dck->aeck", ("edge", 0, (chi, D, chi)), ("U", k0 * n, (chi, D, w)), "Yr"),
        ("contract", "aeck,bjce->abjk", ("Yr", 0, (chi, D, D, w)), ("bulk", 0, (D, D, D, D)), "Yb"),
        ("contract", "abi,abjk->ijk", ("U", 0, (chi, D, chi)), ("Yb", 0, (chi, D, D, w)), "e1_r"),
        ("gather", "c1", "c1_r"),
        ("gather", "e1", "e1_r"),
        ("finish", "c1", "e1", "corner_out", "edge_out"),
    ]


class ShardedCTMRG:
    """ctmrgstep (ctmrg.jl:126-153) with its contractions sharded over the ranks of `group`."""

    def __init__(self, ctx, chi: int, D: int, dist=None, group=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.dist = dist
        self.group = group
        world = dist.get_world_size(group) if dist is not None else 1
        rank = dist.get_rank(group) if dist is not None else 0
        self.plan = shard_plan(chi, D, world, rank)
        self.dev = torch.device("cuda", ctx.device)
        self.buf = {k: torch.empty(v, dtype=torch.float64, device=self.dev) for k, v in buffer_sizes(self.plan).items()}
        self.schedule = step_schedule(self.plan)
        self.sweeps = 0
        self.split_eig = True          # share the back-transformation of the eigen-decomposition between the ranks
        self.ms = {"contract": 0.0, "gather": 0.0, "svd": 0.0}

    # -- helpers ------------------------------------------------------------------------------------------
    def _gather(self, full, part):
        if self.dist is None or self.plan.world == 1:
            full.copy_(part)
            self.torch.cuda.synchronize(self.dev)
            return
        self.dist.all_gather_into_tensor(full, part, group=self.group)
        self.torch.cuda.synchronize(self.dev)

    def load(self, bulk, corner, edge):
        """Upload column-major host arrays (every rank passes the same values)."""
        t = self.torch
        for name, src in (("bulk", bulk), ("corner", corner), ("edge", edge)):
            self.buf[name].copy_(t.from_numpy(np.ascontiguousarray(np.asarray(src, dtype=np.float64).ravel(order="F"))))
        t.cuda.synchronize(self.dev)

    def result(self):
        p = self.plan
        c = self.buf["corner_out"].cpu().numpy().reshape((p.chi, p.chi), order="F")
        e = self.buf["edge_out"].cpu().numpy().reshape((p.chi, p.D, p.chi), order="F")
        s = self.buf["S"].cpu().numpy()
        return c, e, s / s[0]

    def step(self, timing: bool = False):
        """One ctmrgstep on the loaded (bulk, corner, edge); the result lands in corner_out / edge_out / S."""
        t, c, p, B = self.torch, self.ctx, self.plan, self.buf
        P = lambda name, off=0: B[name].data_ptr() + 8 * off
        ms = {"contract": 0.0, "gather": 0.0, "svd": 0.0}
        phase = {"contract": "contract", "permute": "contract", "finish": "contract", "gather": "gather",
                 "svd_symmetrized": "svd"}
        c.set_pointer_mode(1)
        try:
            for op in self.schedule:
                if timing:      # every C-ABI call is synchronous, so events on an idle device bracket it
                    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                    e0.record()
                kind = op[0]
                if kind == "contract":
                    _, spec, (a, oa, da), (b, ob, db), out = op
                    c.dev_contract(spec, P(a, oa), da, P(b, ob), db, P(out))
                elif kind == "gather":
                    self._gather(B[op[1]], B[op[2]])
                elif kind == "permute":
                    c.dev_permute(P(op[1]), op[2], op[3], P(op[4]))
                elif kind == "svd_symmetrized":
                    self._svd(op, P)
                elif kind == "finish":
                    c.dev_ctmrg_finish(P(op[1]), P(op[2]), p.D, p.chi, P(op[3]), P(op[4]))
                else:
                    raise RuntimeError(f"unknown schedule op {kind}")
                if timing:
                    e1.record()
                    e1.synchronize()
                    ms[phase[kind]] += e0.elapsed_time(e1)
        finally:
            c.set_pointer_mode(0)
        if timing:
            self.ms = ms

    def _svd(self, op, P):
        """svd(cp + cp') (ctmrg.jl:134-136).  One rank: one call.  Several ranks: every rank reduces the replicated
        matrix to tridiagonal form and solves the tridiagonal problem (latency-bound stages), back-transforms ITS block
        of N / world eigenvector columns (the 2 x 2 n^3 flop of the decomposition), the blocks are all-gathered and
        every rank sorts / sign-fixes the result."""
        c, p, B, t = self.ctx, self.plan, self.buf, self.torch
        n = op[2]
        if self.dist is None or p.world == 1 or not self.split_eig:
            self.sweeps = c.dev_svd_symmetrized(P(op[1]), n, P(op[3]), P(op[4]), P(op[5]))
            return
        h, N = c.dev_symeig_reduce(P(op[1]), n, True)
        try:
            if N % p.world:
                raise ValueError(f"padded order {N} is not divisible by the number of ranks")
            wcols = N // p.world
            if "Zfull" not in B or B["Zfull"].numel() != N * N:
                B["Zfull"] = t.empty(N * N, dtype=t.float64, device=self.dev)
                B["Zpart"] = t.empty(N * wcols, dtype=t.float64, device=self.dev)
            c.dev_symeig_backtransform(h, p.rank * wcols, wcols, B["Zpart"].data_ptr())
            self._gather(B["Zfull"], B["Zpart"])
            c.dev_symeig_finish(h, B["Zfull"].data_ptr(), P(op[3]), P(op[4]), P(op[5]))
        finally:
            c.symeig_free(h)

    def advance(self):
        """Feed the result back as the next environment."""
        self.buf["corner"].copy_(self.buf["corner_out"])
        self.buf["edge"].copy_(self.buf["edge_out"])
        self.torch.cuda.synchronize(self.dev)
