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
    on the workload's tensor and chi with `maxit = sample_maxit`."""
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
    return dt, info["nsteps"], float(e), g


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        return int(max([p.get("num_threads", 1) for p in threadpool_info()] + [1]))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm (oracle port, all host threads) on OUR arm's configuration
    (same tensor, chi and maxit: `same_config` true unless --ref-maxit overrides it)."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    sample_maxit = args.maxit if args.ref_maxit is None else args.ref_maxit
    cpu_reference_sample(0)                                      # one untimed pass pages in BLAS / LAPACK
    times, nsteps = [], 0
    for _ in range(args.steps):
        dt, nsteps, _, _ = cpu_reference_sample(sample_maxit)
        times.append(dt)
    total = float(np.sum(times))
    val = total / (args.steps * nsteps)
    nthreads = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "same_config": sample_maxit == args.maxit,
        "config": {"workload": f"Heisenberg iPEPS energy+gradient d={D_IPEPS} (CTMRG D={D_IPEPS**2}) chi={CHI} FP64, tol=0 maxit={sample_maxit} "
                               f"({nsteps} ctmrgsteps forward + unrolled backward per call); BASELINE configs[3]",
                   "ipeps_d": D_IPEPS, "chi": CHI, "maxit": sample_maxit, "ctmrgsteps_per_call": nsteps,
                   "parallelism": "host threads (OpenBLAS)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"oracle port (NumPy/SciPy: OpenBLAS dgemm + LAPACK dgesdd), energy+gradient, maxit={sample_maxit}, "
                                   f"{args.steps} calls of {total / args.steps:.1f} s"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
FWD_FLOPS_PER_STEP = None


def contraction_flops(chi, D, s=2):
    """Optimal-order algorithmic flops: forward contractions of one ctmrgstep and of expectationvalue."""
    n = chi * D
    step = (2 * chi ** 3 * D + 2 * chi ** 3 * D ** 2 + 2 * chi ** 2 * D ** 4) + (2 * n * n * chi + 2 * n * chi ** 2) + (4 * chi ** 3 * D ** 2 + 2 * chi ** 2 * D ** 4)
    expv = 4 * chi ** 3 * D + 2 * chi ** 3 * D ** 2 + 2 * chi ** 2 * D ** 4 * s * s + 2 * chi ** 3 * D ** 2 * s * s
    return float(step), float(expv)


def measure_trg_chi64(ctx, T):
    """Third part of BASELINE's metric: steady-state TRG iterations/s at chi=64 (value + gradient).  The bond
    dimension of the Ising tensor saturates at 64 in iteration 5; iterations 6.. split full 4096 x 4096 matrices:
    the steady-state cost per iteration is (t(9 iterations) - t(7 iterations)) / 2."""
    a = T.model_tensor(T.Ising(), 0.44)

    def run(niter):
        ctx.timer_start()
        lnz, g = T.trg_value_and_grad(a, 64, niter, ctx=ctx)
        return ctx.timer_stop(), lnz, g
    run(7)                                   # grows the stream-ordered pool (first touch) for these sizes
    t7, lnz7, g7 = run(7)
    run(9)
    t9, lnz9, _ = run(9)
    per_iter = (t9 - t7) / 2.0
    out = {"iters_per_s": 1e3 / per_iter if per_iter > 0 else None, "ms_per_iteration": per_iter,
           "ms_7_iterations": t7, "ms_9_iterations": t9, "lnZ_7": lnz7,
           "workload": "TRG Ising beta=0.44 chi=64, value + gradient, steady state (two 4096x4096 truncated SVDs + 2 chi^6 contraction per iteration)"}
    fx = os.path.join(ROOT, "tests", "golden", "large.npz")
    if os.path.exists(fx):
        vec = np.load(fx)
        ref, dref = float(vec["trg64_lnz"]), float(vec["trg64_dbeta"])
        db = float(np.sum(g7 * T.dmodel_tensor(T.Ising(), 0.44)))
        out["parity_vs_oracle_fixture"] = {"lnZ_rel_err": abs(lnz7 - ref) / abs(ref), "dlnZ_dbeta_rel_err": abs(db - dref) / abs(dref)}
    return out


def run_ours(args):
    import tnad_b200 as T
    dist = Dist()
    rank, local_rank, world = dist.rank, dist.local_rank, dist.world
    ctx = T.Context(local_rank)
    d, s = D_IPEPS, S_PHYS
    h = heisenberg_h()
    A = ipeps_tensor(rank)                      # every rank an independent replica (its own seed)

    # -- device-resident arm (`value`): inputs and the gradient output live in HBM
    hp, Ap, gp = ctx.dev_alloc(h.size), ctx.dev_alloc(A.size), ctx.dev_alloc(A.size)
    ctx.dev_upload(hp, h); ctx.dev_upload(Ap, A)
    for _ in range(args.warmup):
        ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    dist.barrier()
    ctx.reset_launch_count()
    ctx.timer_start()
    e_last = 0.0
    for _ in range(args.steps):
        e_last = ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
        if os.environ.get("BENCH_DEBUG_CALLS"):
            print("[bench debug] device arm call:", ctx.last_timing(), file=sys.stderr, flush=True)
    ms = ctx.timer_stop()
    launches = ctx.launch_count()
    dist.barrier()
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
        if os.environ.get("BENCH_DEBUG_CALLS"):
            print("[bench debug] e2e arm call:", ctx.last_timing(), file=sys.stderr, flush=True)
    ms_e2e = dist.max(ctx.timer_stop())
    sampler.stop_flag = True
    h2d = int(h.nbytes + A.nbytes)
    d2h = int(8 + A.nbytes)

    extra = {}
    if world > 1:
        # the north-star multi-GPU configurations (BASELINE configs[4]): the 64-instance TRG beta sweep (no collective)
        # and the chi-sharded ctmrgstep at d=5, chi=256 (NCCL all-gather of the enlarged corner)
        import bench_sweep
        import bench_sharded
        try:
            sw = bench_sweep.measure(dist, ctx, 64, 20, 20)
            extra["sweep_instances_per_s"] = sw["instances_per_s"]
            extra["sweep"] = sw
        except Exception as ex:   # noqa: BLE001
            extra["sweep_error"] = repr(ex)
        try:
            if 256 % world == 0:
                sh = bench_sharded.measure(dist, ctx, 5, 256, 2, 1, False)
                extra["sharded_s_per_step"] = sh["s_per_step"]
                extra["sharded_ms_contract"] = sh["ms_contract"]
                extra["sharded_ms_gather"] = sh["ms_gather"]
                extra["sharded_ms_svd"] = sh["ms_svd"]
                extra["sharded"] = sh
        except Exception as ex:   # noqa: BLE001
            extra["sharded_error"] = repr(ex)

    if rank == 0:
        # -- per-kernel-family device time (extra instrumented pass, outside the timed regions)
        ctx.set_kernel_timing(True)
        ctx.energy_device(hp, Ap, d, s, CHI, 0.0, args.maxit, gp)
        kt = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        dmma_peak = ctx.dmma_peak()
        pk, pk_kind = peaks()
        hbm = float(pk.get("hbm_gbs", 0.0)) or None
        n_mat = CHI * D_IPEPS ** 2
        n_svd = steps_done
        fam_ms = {k: v["ms"] for k, v in kt.items()}
        gemm_flops = float(kt["gemm"].get("flops", 0.0))
        dom = max(fam_ms, key=lambda k: fam_ms[k])
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        # algorithmic work of the eigensolver's kernels per decomposition
        b = 32
        fl_symm = 2.0 * b * sum((n_mat - j - b) ** 2 for j in range(0, n_mat - b - 1, b))          # Z0 = A22 Y per panel
        fl_r64 = 2.0 * 2 * b * sum((n_mat - j - b) ** 2 for j in range(0, n_mat - b - 1, b))       # A22 -= [Y W][W Y]'
        fl_q2 = 2.0 * n_mat ** 3                                                                  # n^2 / (2 b) reflectors x 4 b n
        by_chase = 8.0 * (33 * n_mat + n_mat * (n_mat - 2) / 2 + n_mat * (n_mat - 2) / (2 * b))    # band in, reflectors + taus out

        def tf(flops_per_svd, fam):
            t = kt[fam]["ms"] * 1e-3
            return flops_per_svd * n_svd / t / 1e12 if t > 0 else None
        kernels = {
            "k_chase (band -> tridiagonal, systolic bulge chasing; bound by the chain of 2n dependent hops)": {
                "ms": round(kt["chase"]["ms"], 3), "launches": kt["chase"]["launches"], "bound": "hbm",
                "achieved_gbs": by_chase * n_svd / (kt["chase"]["ms"] * 1e-3) / 1e9 if kt["chase"]["ms"] > 0 else None,
                "algorithmic_bytes_per_launch": by_chase},
            "k_q2_stage (back-transformation with Q2, FP64 vector FMA, register resident)": {
                "ms": round(kt["q2_stage"]["ms"], 3), "launches": kt["q2_stage"]["launches"], "bound": "fp64 vector",
                "achieved_tflops": tf(fl_q2, "q2_stage"), "frac_of_fp64_peak": (tf(fl_q2, "q2_stage") or 0) / dmma_peak if dmma_peak else None},
            "k_panel_gram / k_panel_qr (panel QR of the band reduction, one thread-block cluster)": {
                "ms": round(kt["panel_qr"]["ms"], 3), "launches": kt["panel_qr"]["launches"], "bound": "latency (serial column recurrence)"},
            "k_symm_y (+ k_reduce_g): Z0 = A22 Y, FP64 DMMA": {
                "ms": round(kt["symm_y"]["ms"], 3), "launches": kt["symm_y"]["launches"], "bound": "tensor",
                "achieved_tflops": tf(fl_symm, "symm_y"), "frac": (tf(fl_symm, "symm_y") or 0) / dmma_peak if dmma_peak else None},
            "k_rank64_update: A22 -= [Y W][W Y]', FP64 DMMA": {
                "ms": round(kt["rank64_update"]["ms"], 3), "launches": kt["rank64_update"]["launches"], "bound": "tensor",
                "achieved_tflops": tf(fl_r64, "rank64_update"), "frac": (tf(fl_r64, "rank64_update") or 0) / dmma_peak if dmma_peak else None},
            "gemm_tma_kernel / gemm_dmma_kernel (einsum contractions, svd_back, explicit Q1 / Q1 Q2 / Qfull Z, divide-and-conquer merges)": {
                "ms": round(kt["gemm"]["ms"], 3), "launches": kt["gemm"]["launches"], "bound": "tensor",
                "launches_tma": kt["gemm"].get("launches_tma"), "launches_cp_async": kt["gemm"].get("launches_cp_async"),
                "flops": gemm_flops, "flops_on_tma_kernel": kt["gemm"].get("flops_tma"),
                "achieved_tflops": gemm_flops / (kt["gemm"]["ms"] * 1e-3) / 1e12 if kt["gemm"]["ms"] > 0 else None,
                "frac": gemm_flops / (kt["gemm"]["ms"] * 1e-3) / 1e12 / dmma_peak if (kt["gemm"]["ms"] > 0 and dmma_peak) else None},
            "divide and conquer (non-GEMM kernels)": {"ms": round(kt["stedc"]["ms"], 3), "launches": kt["stedc"]["launches"]},
        }
        if dom == "gemm" and gemm_flops > 0:
            # the GEMM family (one kernel template, all einsum contractions + the dense products of the eigensolver) is the
            # largest share of the call: FP64 tensor roofline, algorithmic flops = sum of 2 M N K over its launches
            ach = gemm_flops / (kt["gemm"]["ms"] * 1e-3) / 1e12
            nl = max(1, kt["gemm"]["launches"])
            roof = {"bound": "tensor", "kernel": "gemm_tma_kernel (FP64 DMMA m8n8k4, TMA tensor-map loads, persistent two-group CTAs) + its cp.async "
                                                  "fallback gemm_dmma_kernel for operands the tensor maps cannot describe",
                    "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak if dmma_peak else None, "traffic": None,
                    "flops_per_launch": gemm_flops / nl, "avg_launch_us": 1e3 * kt["gemm"]["ms"] / nl,
                    "peak_source": "tnad_dmma_peak (register-resident DMMA issue loop measured in this run; MEASURED_PEAKS.json has no FP64 entry)",
                    "note": "dominant kernel family by device time of the call (events around every launch of one instrumented pass); "
                            "achieved = sum of 2 M N K over the family's launches / the family's device time, small and skinny products included. "
                            "k_chase is the largest single launch (latency bound, see `kernels`)"}
        elif dom == "chase" or kt["chase"]["ms"] >= max(kt["q2_stage"]["ms"], kt["panel_qr"]["ms"]):
            ach = by_chase * n_svd / (kt["chase"]["ms"] * 1e-3) / 1e9 if kt["chase"]["ms"] > 0 else 0.0
            roof = {"bound": "hbm", "kernel": "k_chase: band (half bandwidth 32) -> tridiagonal by bulge chasing, one launch per eigen-decomposition; "
                                              "systolic array of warps, band resident in shared memory of a thread-block cluster",
                    "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if hbm else None,
                    "traffic": traffic.get("k_chase_dram_bytes_per_launch"),
                    "bytes_per_launch": by_chase, "avg_launch_us": 1e3 * kt["chase"]["ms"] / max(1, kt["chase"]["launches"]),
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs [{pk_kind}]",
                    "note": "largest single kernel of the call; its HBM traffic (33 n doubles in, n^2/2 reflector doubles out) is negligible: the kernel is "
                            "bound by the chain of 2n dependent Householder hops (about 2.3 us per sweep), not by any throughput roof. The throughput-bound "
                            "kernels of the call are listed under `kernels` with their own fractions (FP64 tensor peak from tnad_dmma_peak)"}
        else:
            famname = {"gemm": "gemm_dmma_kernel", "q2_stage": "k_q2_stage", "panel_qr": "k_panel_gram"}.get(dom, dom)
            roof = {"bound": "tensor", "kernel": famname, "achieved": tf(fl_q2, "q2_stage") if dom == "q2_stage" else None, "peak": dmma_peak,
                    "unit": "TFLOP/s", "frac": None, "traffic": None}
            if roof["achieved"]:
                roof["frac"] = roof["achieved"] / dmma_peak
        roof["kernel_ms_instrumented_pass"] = {k: round(v["ms"], 3) for k, v in kt.items()}
        roof["kernel_launches"] = {k: v["launches"] for k, v in kt.items()}
        roof["largest_family_by_device_time"] = dom
        roof["kernels"] = kernels
        roof["fp64_tensor_peak_tflops"] = dmma_peak
        # -- the step's contractions on their real shapes (not a synthetic GEMM): on-device span timers of the last call
        f_step, f_expv = contraction_flops(CHI, D_IPEPS ** 2, S_PHYS)
        cms = timing["contractions"]
        ctf = (f_step * steps_done + f_expv) / (cms * 1e-3) / 1e12 if cms > 0 else None
        contractions = {"flops_per_step": f_step, "flops_expectationvalue": f_expv, "ms": cms, "steps": steps_done,
                        "tflops": ctf, "frac_of_dmma_peak": ctf / dmma_peak if (ctf and dmma_peak) else None,
                        "what": "forward contractions of all ctmrgsteps + expectationvalue of one energy call (optimal-order flop count, "
                                "SURVEY.md 8d), timed by CUDA events around the einsum chains"}
        # -- CPU baseline on a bounded sample (same box, same run) and parity of the GPU result at that sample
        cpu, parity = None, None
        if not (args.no_cpu_baseline or world > 1):      # the CPU baseline is measured on rank 0 at N = 1 only
            bm = args.cpu_maxit
            cdt, cns, ce, cg = cpu_reference_sample(bm)
            cpu = {"value": cdt / cns, "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": f"oracle port (OpenBLAS dgemm + LAPACK dgesdd) energy+gradient, same tensor and chi, maxit={bm} ({cns} steps, {cdt:.1f} s)"}
            ge, gg = ctx.energy(hh, Ah, CHI, 0.0, bm, grad=True)
            parity = {"maxit": bm, "energy_gpu": ge, "energy_cpu": ce, "energy_rel_err": abs(ge - ce) / abs(ce),
                      "grad_rel_err": float(np.linalg.norm(gg - cg) / np.linalg.norm(cg)),
                      "ok": bool(abs(ge - ce) <= 1e-10 * abs(ce) and np.linalg.norm(gg - cg) <= 1e-8 * np.linalg.norm(cg))}
        trg64 = None
        if world == 1 and not args.no_trg:
            try:
                trg64 = measure_trg_chi64(ctx, T)
            except Exception as ex:   # noqa: BLE001
                trg64 = {"error": repr(ex)}
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
            "contractions": contractions,
            "cpu_baseline": cpu,
            "parity_check": parity,
            "trg_chi64_iters_per_s": trg64["iters_per_s"] if trg64 and "iters_per_s" in trg64 else None,
            "trg_chi64": trg64,
            "breakdown_ms_last_call": {k: round(v, 3) for k, v in timing.items()},
            "energy": e_last,
        }
        line.update(extra)
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
    ap.add_argument("--ref-maxit", type=int, default=None, help="reference arm: maxit of the oracle run (default: --maxit, same config)")
    ap.add_argument("--cpu-maxit", type=int, default=3, help="our arm: maxit of the bounded CPU baseline sample / parity check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-trg", action="store_true", help="skip the TRG chi=64 part of the metric")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
