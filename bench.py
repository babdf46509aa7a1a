#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 CTMRG hot path (BASELINE.json `metric`).

Workload (`config.workload`): BASELINE.json configs[3] -- Heisenberg iPEPS energy + gradient, iPEPS
bond d=4 (CTMRG bulk dimension D=16), chi=128, FP64, `tol=0, maxit=10` (the reference's stop rule then
runs exactly 11 ctmrgsteps), seeded synthetic iPEPS tensor.  One bench "step" is one
`energy(h, ipeps; chi, tol, maxit)` + its gradient (forward CTMRG, expectation value, unrolled reverse
sweep).  The metric is seconds per ctmrgstep (forward + backward): value = time / (steps * 11 * n_gpus).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path (oracle port)

N > 1 is launched by torchrun (one rank per GPU); ranks run independent replicas (weak scaling, no
data-path collective -- SURVEY.md section 8e), timed as max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_IPEPS, CHI, MAXIT, S_PHYS = 4, 128, 10, 2
METRIC = "ctmrg_energy_grad_seconds_per_step_d4_chi128"
UNIT = "s/step"


def heisenberg_h():
    sx = np.array([[0.0, 1.0], [1.0, 0.0]]); sy = np.array([[0.0, -1j], [1j, 0.0]]); sz = np.array([[1.0, 0.0], [0.0, -1.0]])
    h = np.einsum("ij,kl->ijkl", sz, sz) - np.einsum("ij,kl->ijkl", sx, sx) - np.einsum("ij,kl->ijkl", sy, sy)
    h = np.einsum("ijcd,kc,ld->ijkl", h, sx, sx.conj().T)
    return np.asfortranarray(np.real(h / 2))


def ipeps_tensor(seed, d=D_IPEPS, s=S_PHYS):
    """Seeded synthetic iPEPS: symmetrised standard-normal tensor."""
    x = np.random.default_rng(seed).standard_normal((d, d, d, d, s))
    for p in [(0, 3, 2, 1, 4), (2, 1, 0, 3, 4), (1, 0, 3, 2, 4), (3, 2, 1, 0, 4)]:
        x = x + np.transpose(x, p)
    return np.asfortranarray(x / np.linalg.norm(x))


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class Dist:
    """torch.distributed plumbing (NCCL) only when launched under torchrun; nothing at N=1."""

    def __init__(self, backend="nccl"):
        self.rank, self.local_rank, self.world = dist_env()
        self.on = self.world > 1
        if self.on:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend)
            self.dev = torch.device("cuda", self.local_rank) if backend == "nccl" else torch.device("cpu")

    def barrier(self):
        if self.on:
            self.dist.barrier()
            if self.dev.type == "cuda":
                self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if not self.on:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if not self.on:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.on:
            self.dist.destroy_process_group()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
def cpu_reference_sample(sample_maxit):
    """The reference's CPU path (oracle port: NumPy/SciPy -> OpenBLAS dgemm + LAPACK dgesdd, all host threads)
    on a bounded sample of the workload: same tensor, same chi, `maxit = sample_maxit`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnad_oracle as O
    try:    # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core it can use
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    h, A = heisenberg_h(), ipeps_tensor(0)
    info = {}
    t0 = time.perf_counter()
    e, g = O.energy_value_and_grad(h, A, CHI, 0.0, sample_maxit, info=info)
    dt = time.perf_counter() - t0
    return dt, info["nsteps"], float(e)


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    try:
        from threadpoolctl import threadpool_info
        nthreads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        nthreads = os.cpu_count() or 1
    sample_maxit = args.ref_maxit
    for _ in range(args.warmup if args.warmup < 1 else 1):      # one untimed pass pages in BLAS/LAPACK
        cpu_reference_sample(0)
    times, nsteps = [], 0
    for _ in range(args.steps):
        dt, nsteps, _ = cpu_reference_sample(sample_maxit)
        times.append(dt)
    total = float(np.sum(times))
    val = total / (args.steps * nsteps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Heisenberg iPEPS energy+gradient d={D_IPEPS} (D={D_IPEPS**2}) chi={CHI} FP64, tol=0; bounded sample maxit={sample_maxit} "
                               f"({nsteps} ctmrgsteps per call) of the maxit={MAXIT} workload", "ipeps_d": D_IPEPS, "chi": CHI,
                   "maxit": sample_maxit, "parallelism": "host threads (OpenBLAS)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(nthreads), "kind": "port",
                         "sample": f"oracle port (NumPy/SciPy: OpenBLAS dgemm + LAPACK dgesdd), energy+gradient with maxit={sample_maxit}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import tnad_b200 as T
    dist = Dist()
    rank, local_rank, world = dist.rank, dist.local_rank, dist.world
    ctx = T.Context(local_rank)
    d, s = D_IPEPS, S_PHYS
    h = heisenberg_h()
    A = ipeps_tensor(rank)                      # every rank an independent replica (its own seed)
    nsteps_per_call = args.maxit + 1

    # -- device-resident arm (`value`): inputs and the gradient output live in HBM
    hp, Ap, gp = ctx.dev_alloc(h.size), ctx.dev_alloc(A.size), ctx.dev_alloc(A.size)
    ctx.dev_upload(hp, h); ctx.dev_upload(Ap, A)
    for _ in range(args.warmup):
        ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier()
    ctx.reset_launch_count()
    ctx.timer_start()
    e_last = 0.0
    for _ in range(args.steps):
        e_last = ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
    ms = ctx.timer_stop()
    launches = ctx.launch_count()
    dist.barrier()
    sampler.stop_flag = True
    ms = dist.max(ms)
    launches_total = int(dist.sum(float(launches)))
    steps_done = ctx.last_steps
    timing = ctx.last_timing()

    # -- end-to-end arm (`e2e`): host buffers through the C ABI, H2D of inputs and D2H of results inside
    hh, Ah = ctx.host_alloc(h.shape), ctx.host_alloc(A.shape)
    hh[...] = h; Ah[...] = A
    ctx.energy(hh, Ah, CHI, 0.0, args.maxit, grad=True)
    dist.barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        e_h, g_h = ctx.energy(hh, Ah, CHI, 0.0, args.maxit, grad=True)
    ms_e2e = dist.max(ctx.timer_stop())
    h2d = int(h.nbytes + A.nbytes)
    d2h = int(8 + A.nbytes)

    if rank == 0:
        # -- roofline of the dominant kernel family (extra instrumented pass, outside the timed regions)
        ctx.set_kernel_timing(True)
        ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
        kt = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        dmma_peak = ctx.dmma_peak()
        gemm_tf = ctx.gemm_bench(4096, 4096, 4096, 3)
        # executed work is counted on the device (skipped, already-converged pairs do no work), so the flop
        # numerators below are exact: one k_sym_update_m block = two 64^3 products, one k_jacobi_update slab =
        # one 128x64x64 product; the pivot kernel is FP64 vector math (dots + plane rotations).
        mu, qu, pe = kt["m_update"], kt["q_update"], kt["eig_panel"]
        fl_m = mu.get("blocks", 0) * 2 * 2.0 * 64 ** 3
        fl_q = qu.get("slabs", 0) * 2.0 * 128 * 64 * 64
        tf_m = fl_m / (mu["ms"] * 1e-3) / 1e12 if mu["ms"] > 0 else 0.0
        tf_q = fl_q / (qu["ms"] * 1e-3) / 1e12 if qu["ms"] > 0 else 0.0
        pk, pk_kind = peaks()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("k_sym_update_m_dram_bytes_per_launch")
        dom = max(("m_update", "eig_panel", "q_update", "gemm"), key=lambda k: kt[k]["ms"])
        roof_jacobi = {"bound": "tensor",
                "kernel": "k_sym_update_m (fused two-sided 64x64 block update M <- W'MW of the symmetric block-Jacobi "
                          "eigensolver, FP64 DMMA m8n8k4)",
                "achieved": tf_m, "peak": dmma_peak, "unit": "TFLOP/s", "frac": tf_m / dmma_peak if dmma_peak else None,
                "traffic": traffic,
                "flops_per_launch": fl_m / mu["launches"] if mu["launches"] else None,
                "avg_launch_us": 1e3 * mu["ms"] / mu["launches"] if mu["launches"] else None,
                "peak_source": "FP64 DMMA issue-rate microbenchmark (tnad_dmma_peak) run in this process: MEASURED_PEAKS.json "
                               f"carries no FP64 figure (its hbm_gbs={pk.get('hbm_gbs')} [{pk_kind}] bounds the elementwise kernels)",
                "kernel_ms_instrumented_pass": {k: round(v["ms"], 3) for k, v in kt.items()},
                "kernel_launches": {k: v["launches"] for k, v in kt.items()},
                "largest_family_by_device_time": dom,
                "other_kernels": {
                    "k_jacobi_update (Q <- Q W panel rotation, FP64 DMMA)": {"achieved_tflops": tf_q, "frac": tf_q / dmma_peak if dmma_peak else None},
                    "k_sym_eig (64x64 pivot eigenproblem, register-resident Jacobi, FP64 vector + shuffles; latency bound, "
                    "runs on N/64 SMs concurrently with the DMMA updates)": {"ms": round(pe["ms"], 3), "launches": pe["launches"]},
                    "gemm_dmma_kernel (einsum contractions) on 4096^3": {"achieved_tflops": gemm_tf, "frac": gemm_tf / dmma_peak if dmma_peak else None}}}
        if mu["launches"] > 0:
            roof = roof_jacobi
        else:
            # Direct eigensolver (default from n >= 256): the dominant kernel is the persistent tridiagonalisation panel
            # kernel k_sytrd_panel (timed in the `pivot_eig` slot).  Its algorithmic traffic is the trailing matrix read
            # once per reflector: 8 * sum_j (n-j-1)^2 bytes per decomposition (~ 8 n^3 / 3), DESIGN.md section 4.2.
            n_mat = CHI * D_IPEPS ** 2
            # columns [0, j_tail) run in the grid-wide panel kernel (32 per launch), the last t = n - j_tail <= 416 columns
            # in ONE launch of the cluster kernel, which loads the trailing matrix once into distributed shared memory
            j_tail = ((n_mat - 416 + 31) // 32) * 32 if n_mat > 416 else 0
            panels = j_tail // 32 + 1
            n_svd = pe["launches"] / panels if panels else 0
            bytes_svd = 8.0 * sum((n_mat - j - 1) ** 2 for j in range(j_tail)) + 8.0 * (n_mat - j_tail) ** 2
            gbs = bytes_svd * n_svd / (pe["ms"] * 1e-3) / 1e9 if pe["ms"] > 0 else 0.0
            hbm = float(pk.get("hbm_gbs", 0.0)) or None
            tr = None
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tr = json.load(f).get("k_sytrd_panel_dram_bytes_per_launch")
            roof = {"bound": "hbm",
                    "kernel": "k_sytrd_panel1 (+ k_sytrd_tail_cluster for the last 416 columns): persistent cooperative Householder "
                              "tridiagonalisation, per column one matrix-vector product over the trailing matrix and one grid-wide exchange; FP64 vector",
                    "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if hbm else None, "traffic": tr,
                    "bytes_per_launch": bytes_svd / panels if panels else None,
                    "avg_launch_us": 1e3 * pe["ms"] / pe["launches"] if pe["launches"] else None,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs [{pk_kind}]",
                    "note": "synchronisation bound at n=2048: the 33.5 MB trailing matrix is L2 resident, every column costs one grid-wide exchange "
                            "(~3 L2 round trips) plus gather and matvec; the fraction says how far the column loop is from streaming the matrix at HBM speed",
                    "kernel_ms_instrumented_pass": {("sytrd_panel" if k == "eig_panel" else k): round(v["ms"], 3) for k, v in kt.items()},
                    "kernel_launches": {("sytrd_panel" if k == "eig_panel" else k): v["launches"] for k, v in kt.items()},
                    "largest_family_by_device_time": "sytrd_panel" if dom == "eig_panel" else dom,
                    "other_kernels": {
                        "gemm_dmma_kernel (einsum contractions, trailing updates, back-transform, D&C merges) on 4096^3":
                            {"achieved_tflops": gemm_tf, "peak_tflops": dmma_peak, "frac": gemm_tf / dmma_peak if dmma_peak else None}}}
        # -- CPU baseline on a bounded sample (same box, same run)
        if args.no_cpu_baseline or world > 1:      # the CPU baseline is measured on rank 0 at N = 1 only
            cpu = None
        else:
            cdt, cns, _ = cpu_reference_sample(args.ref_maxit)
            try:
                from threadpoolctl import threadpool_info
                nthreads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
            except Exception:
                nthreads = os.cpu_count() or 1
            cpu = {"value": cdt / cns, "unit": UNIT, "cores": int(nthreads), "kind": "port",
                   "sample": f"oracle port (OpenBLAS dgemm + LAPACK dgesdd) energy+gradient, same tensor and chi, maxit={args.ref_maxit} ({cns} steps, {cdt:.1f} s)"}
        total_units = args.steps * steps_done * world
        line = {
            "metric": METRIC, "value": ms * 1e-3 / total_units, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Heisenberg iPEPS energy+gradient d={D_IPEPS} (CTMRG D={D_IPEPS**2}) chi={CHI} FP64, tol=0 maxit={args.maxit} "
                                   f"({steps_done} ctmrgsteps forward + unrolled backward per call); BASELINE configs[3]",
                       "ipeps_d": D_IPEPS, "chi": CHI, "maxit": args.maxit, "ctmrgsteps_per_call": steps_done,
                       "parallelism": f"replicas x{world} (no collective)",
                       "l2_note": "per-step working set (cp, U, V tapes: 3 x 33.5 MB x 11 steps plus SVD work buffers) exceeds the 126 MB L2; "
                                  "no explicit flush needed"},
            "e2e": {"value": ms_e2e * 1e-3 / total_units, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_total,
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "breakdown_ms_last_call": {k: round(v, 3) for k, v in timing.items()},
            "energy": e_last,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    for p in (hp, Ap, gp):
        ctx.dev_free(p)
    ctx.close()
    dist.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--maxit", type=int, default=MAXIT)
    ap.add_argument("--ref-maxit", type=int, default=1, help="bounded CPU sample: maxit of the oracle run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
