"""Measurement helper: wall/device time of the BASELINE configs other than the headline (record for profiles/)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T
import tnad_oracle as O
ctx = T.Context(0)
which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
NO_ORACLE = bool(os.environ.get("NO_ORACLE"))   # GPU timing only (the chi = 64 CTMRG oracle alone takes a minute per beta)
h = T.hamiltonian(T.Heisenberg())
def timed(f, reps=2):
    f(); best = 1e9
    for _ in range(reps):
        ctx.timer_start(); out = f(); best = min(best, ctx.timer_stop())
    return best, out
if "c1" in which:
    a = T.model_tensor(T.Ising(), 0.5)
    ms, (lnz, g) = timed(lambda: T.trg_value_and_grad(a, 20, 20, ctx=ctx))
    cpu = float('nan')
    if not NO_ORACLE:
        t0 = time.time(); ref = O.trg_dbeta(0.5, 20, 20); cpu = time.time() - t0
    print(f"C1 TRG Ising beta=0.5 chi=20 niter=20 value+grad: GPU {ms:.1f} ms, CPU oracle {cpu*1e3:.1f} ms; lnZ={lnz!r} dbeta={float(np.sum(g*T.dmodel_tensor(T.Ising(),0.5)))!r}", flush=True)
if "c2" in which:
    for beta in (0.3, 0.5):
        a, m = T.model_tensor(T.Ising(), beta), T.mag_tensor(T.Ising(), beta)
        c0, e0 = O.init_random(a, 64, np.random.default_rng(5))
        ms, (c, e, vals, steps) = timed(lambda: ctx.ctmrg(a, c0, e0, 1e-10, 3000), reps=1)
        mag = ctx.magnetisation_readout(a, m, c, e)
        if NO_ORACLE: cpu, no, mo = float('nan'), -1, float('nan')
        else:
            t0 = time.time(); co, eo, vo, no = O.ctmrg(a, c0, e0, 1e-10, 3000); cpu = time.time() - t0
            mo = O.magnetisation_readout(a, m, co, eo)
        print(f"C2 CTMRG Ising beta={beta} chi=64 tol=1e-10 (:random seed 5): GPU {ms:.1f} ms for {steps} steps ({ms/steps:.2f} ms/step), CPU oracle {cpu*1e3:.0f} ms for {no} steps; "
              f"mag GPU {mag:.12f} oracle {mo:.12f} onsager {T.magofbeta(T.Ising(), beta):.12f}", flush=True)
if "c3" in which:
    A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))
    ms, (e, g) = timed(lambda: ctx.energy(h, A, 20, 1e-6, 100, grad=True))
    t0 = time.time(); eo, go = O.energy_value_and_grad(h, A, 20, 1e-6, 100); cpu = time.time() - t0
    print(f"C3 Heisenberg d=2 chi=20 tol=1e-6 energy+grad: GPU {ms:.1f} ms ({ctx.last_steps} steps), CPU oracle {cpu*1e3:.1f} ms; E={e!r} rel {abs(e-eo)/abs(eo):.1e} grad rel {np.linalg.norm(g-go)/np.linalg.norm(go):.1e}", flush=True)
if "c5" in which:
    A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((5, 5, 5, 5, 2)))
    ap, a = O.double_layer(A)
    c0, e0 = ctx.ctmrg_init_raw(a, 256)
    t0 = time.time(); c, e, vals, steps = ctx.ctmrg(a, c0, e0, 0.0, 2); dt = time.time() - t0
    tm = ctx.last_timing()
    print(f"C5a CTMRG d=5 (D=25) chi=256 (n=6400) forward: {steps} steps in {dt:.2f} s wall; timing {tm}; vals[1]={vals[1]:.6f} vals[255]={vals[255]:.3e}", flush=True)
