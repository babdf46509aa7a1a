"""Bring-up helper: run the oracle's TRG reverse sweep with the GPU SVD plugged in."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_b200 as T
import tnad_oracle as O
ctx = T.Context(0)
beta, chi, niter = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ref = O.trg_dbeta(beta, chi, niter)
orig = O.svd
log = []
def gsvd(A, driver="gesdd"):
    U, S, V = ctx.svd(np.asfortranarray(A))
    Ul, Sl, Vl = orig(A)
    log.append((A.shape, S.copy(), Sl.copy(), np.abs(U.T @ U - np.eye(len(S))).max(), np.abs(V.T @ V - np.eye(len(S))).max(),
                np.linalg.norm((U * S) @ V.T - A)))
    return U, S, V
O.svd = gsvd
got = O.trg_dbeta(beta, chi, niter)
O.svd = orig
print("oracle(LAPACK)", ref)
print("oracle(GPU svd)", got, "rel", abs(got[1] - ref[1]) / abs(ref[1]))
lnz, g = T.trg_value_and_grad(O.model_tensor_ising(beta), chi, niter, ctx=ctx)
print("library", lnz, float(np.sum(g * O.dmodel_tensor_ising(beta))))
for shp, S, Sl, ou, ov, rec in log[:12]:
    print(shp, "orthU %.1e orthV %.1e rec %.1e" % (ou, ov, rec), "S gpu", np.array2string(S[:8], precision=3), "lapack", np.array2string(Sl[:8], precision=3))
