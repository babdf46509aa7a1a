import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnad_b200 as T
ctx = T.Context(0)
a = T.model_tensor(T.Ising(), 0.44)
T.trg(a, 64, 2, ctx=ctx)
ctx.timer_start(); l = T.trg(a, 64, 10, ctx=ctx); print("total ms", ctx.timer_stop(), l)
