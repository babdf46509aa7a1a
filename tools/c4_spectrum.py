import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T
import tnad_oracle as O
ctx = T.Context(0)
A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))
ap, a = O.double_layer(A)
c0, e0 = ctx.ctmrg_init_raw(a, 128)
for maxit in (0, 2, 9):
    c, e, vals, steps = ctx.ctmrg(a, c0, e0, 0.0, maxit)
    print("steps", steps, "vals at", [0, 1, 10, 50, 100, 127, 128, 200, 300, 500, 800, 1200, 1600, 2047], "=",
          np.array2string(vals[[0, 1, 10, 50, 100, 127, 128, 200, 300, 500, 800, 1200, 1600, 2047]], precision=2))
    print("  count > 1e-8:", int((vals > 1e-8).sum()), " > 1e-12:", int((vals > 1e-12).sum()), " > 1e-15:", int((vals > 1e-15).sum()))
ctx.set_kernel_timing(True)
c, e, vals, steps = ctx.ctmrg(a, c0, e0, 0.0, 3)
print(ctx.kernel_timing())
