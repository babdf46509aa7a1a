"""The reference's only published timing (README.md:117-153): optimiseipeps on the Heisenberg model, d = 2, chi = 20,
tol = 1e-6, maxit = 100, Optim.Options(f_tol = 1e-6): 16 L-BFGS iterations in 4.84 s (0.30 s per iteration, hardware
not stated), final energy -0.6602311.  Same configuration through tnad_b200 (SciPy L-BFGS-B over one fused tnad_energy).
usage: optimise_readme.py [seed] [reps]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tnad_b200 as T
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = T.Context(0)
h = T.hamiltonian(T.Heisenberg())
out = []
for r in range(reps):
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(seed + r).random((2, 2, 2, 2, 2))))
    e0 = T.energy(h, ipeps, chi=20, tol=1e-6, maxit=100, ctx=ctx)
    t0 = time.perf_counter()
    res = T.optimiseipeps(ipeps, h, 20, 1e-6, 100, optimargs={"f_tol": 1e-6}, ctx=ctx)
    dt = time.perf_counter() - t0
    out.append(dict(seed=seed + r, e_initial=e0, e_final=res.minimum, iterations=int(res.nit), energy_gradient_calls=int(res.nfev),
                    wall_s=dt, s_per_iteration=dt / max(1, res.nit), ms_per_energy_gradient_call=1e3 * dt / max(1, res.nfev)))
    print(json.dumps(out[-1]), flush=True)
best = min(out, key=lambda o: o["s_per_iteration"])
print(json.dumps(dict(config="Heisenberg d=2 chi=20 tol=1e-6 maxit=100 f_tol=1e-6 (README.md:117-153 of the reference)",
                      reference_published=dict(iterations=16, wall_s=4.84, s_per_iteration=0.30, e_final=-0.6602311),
                      ours_best=best)))
