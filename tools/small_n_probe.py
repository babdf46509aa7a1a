import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnad_b200 as T
ctx = T.Context(0)
rng = np.random.default_rng(0)
for n in (64, 96, 128, 192, 256):
    a = rng.standard_normal((n, n)) * (10.0 ** (-rng.uniform(0, 8, n)))[None, :]; a = a + a.T
    for mode in ("1", "2"):
        os.environ["TNAD_SYMEIG"] = mode
        for _ in range(3): ctx.svd_sym(a)
        ctx.timer_start()
        for _ in range(20): ctx.svd_sym(a)
        ms = ctx.timer_stop() / 20
        print(f"n={n} mode={mode}: {ms:.3f} ms per svd_sym (incl. small h2d/d2h)", flush=True)
