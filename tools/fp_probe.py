import sys, numpy as np
sys.path.insert(0, '/root/repo')
import tnad_b200 as T
ctx = T.Context(0)
h = T.hamiltonian(T.Heisenberg())
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for seed, chi in [(0, 8), (3, 12)]:
    A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(seed).standard_normal((2, 2, 2, 2, 2)))).bulk
    e_f, g_f = T.energy_and_gradient_fixedpoint(h, A, chi, 0.0, 400, bwd_tol=1e-14, bwd_maxit=3000, ctx=ctx)
    print("fixedpoint: e", e_f, "bwd iters", ctx.last_bwd_iters)
    for maxit in (50, 150, 400, 1000):
        e_u, g_u = T.energy_and_gradient(h, A, chi, 0.0, maxit, ctx=ctx)
        print(seed, chi, maxit, "e diff", abs(e_u - e_f), "grad rel diff", rel(g_u, g_f))
    for bm in (10, 30, 100, 300):
        e2, g2 = T.energy_and_gradient_fixedpoint(h, A, chi, 0.0, 400, bwd_tol=1e-30, bwd_maxit=bm, ctx=ctx)
        print("  neumann terms", bm, "rel diff to full", rel(g2, g_f))
