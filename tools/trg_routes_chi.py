"""Compare the two TRG SVD routes at a given chi (development helper): lnZ and dlnZ/dbeta."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnad_b200 as T
chi, niter, beta = int(sys.argv[1]), int(sys.argv[2]), 0.44
ctx = T.Context(0)
a = T.model_tensor(T.Ising(), beta)
da = T.dmodel_tensor(T.Ising(), beta)
out = {}
for mode in ("jacobi", "dc"):
    os.environ["TNAD_TRG_SVD"] = mode
    lnz, g = T.trg_value_and_grad(a, chi, niter, ctx=ctx)
    out[mode] = (lnz, float(np.sum(g * da)))
    print(mode, repr(lnz), repr(out[mode][1]), flush=True)
(l1, d1), (l2, d2) = out["jacobi"], out["dc"]
print(f"chi={chi} niter={niter}: rel diff lnZ {abs(l1 - l2) / abs(l1):.2e}, dlnZ/dbeta {abs(d1 - d2) / abs(d1):.2e}")
