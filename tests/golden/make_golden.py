"""Generates tests/golden/published.json and tests/golden/vectors.npz.

published.json: every exact value the reference publishes for the hot path, with its source line.
vectors.npz   : outputs of the CPU oracle (oracle/tnad_oracle.py) on seeded inputs.  The reference is pure
                Julia and cannot be executed in this image (no Julia), so these vectors come from the pinned
                restatement, not from the reference itself; they freeze the oracle so that the GPU parity tests
                do not depend on the LAPACK build of the machine they run on.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_oracle as O  # noqa: E402

published = {
    "trg_beta0.4_chi5_n5": {"value": 0.8919788686747141, "source": "test/trg.jl:18 (tensorgrad)"},
    "trg_beta0.5_chi20_n20": {"value": 1.0257933734351765, "source": "README.md:59-60"},
    "dtrg_beta0.5_chi20_n20": {"value": 1.7455677143228514, "source": "README.md:69-70"},
    "dtrg_beta0.5_chi5_n5": {"value": 1.7502426939979507, "source": "docs/src/userguide.md:25-29"},
    "heisenberg_h": {"value": O.hamiltonian_heisenberg().reshape(-1, order="F").tolist(),
                     "expected_print": {"[:,:,1,1]": [[-0.5, 0.0], [0.0, 0.5]], "[:,:,2,1]": [[0.0, 0.0], [-1.0, 0.0]],
                                        "[:,:,1,2]": [[0.0, -1.0], [0.0, 0.0]], "[:,:,2,2]": [[0.5, 0.0], [0.0, -0.5]]},
                     "source": "README.md:83-100"},
    "heisenberg_energy_d2": {"value": -0.66023, "atol": 1e-3, "source": "test/variationalipeps.jl:102, README.md:152-155"},
    "onsager": {"source": "test/ctmrg.jl:37-42, src/exampletensors.jl:77",
                "cases": [[1.0, 2, 1e-8], [0.6, 4, 1e-8], [0.8, 2, 1e-8]]},
}

vec = {}
rng = np.random.default_rng(2024)
h = O.hamiltonian_heisenberg()
# C3: d=2, chi=20, tol=1e-6, maxit=100
A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))
info = {}
e, g = O.energy_value_and_grad(h, A, 20, 1e-6, 100, info=info)
vec.update(c3_A=A, c3_e=e, c3_grad=g, c3_steps=info["nsteps"])
# fixed-maxit small cases
for name, d, chi, maxit, seed in [("e_d2_chi4", 2, 4, 10, 1), ("e_d3_chi12", 3, 12, 6, 2), ("e_d2_chi16", 2, 16, 8, 3)]:
    A = O.indexperm_symmetrize(np.random.default_rng(seed).standard_normal((d, d, d, d, 2)))
    info = {}
    e, g = O.energy_value_and_grad(h, A, chi, 0.0, maxit, info=info)
    vec.update({name + "_A": A, name + "_e": e, name + "_grad": g, name + "_steps": info["nsteps"],
                name + "_cfg": np.array([d, chi, maxit])})
# TRG values + gradients
for beta, chi, n in [(0.4, 5, 5), (0.5, 5, 5), (0.44, 8, 10), (0.5, 20, 20)]:
    lnz, ga = O.trg_value_and_grad(O.model_tensor_ising(beta), chi, n)
    key = f"trg_{beta}_{chi}_{n}"
    vec[key + "_lnz"] = lnz
    vec[key + "_grad"] = ga
    vec[key + "_dbeta"] = float(np.sum(ga * O.dmodel_tensor_ising(beta)))
# CTMRG Ising, :raw init (deterministic), singular-value spectrum + step counts
for beta, chi in [(0.3, 16), (0.5, 16)]:
    a = O.model_tensor_ising(beta)
    c0, e0 = O.init_raw(a, chi)
    c, ed, vals, ns = O.ctmrg(a, c0, e0, 1e-10, 500)
    vec[f"ctmrg_raw_{beta}_{chi}_vals"] = vals
    vec[f"ctmrg_raw_{beta}_{chi}_steps"] = ns
# svd_back sample (square, U/S/V cotangents)
M = rng.standard_normal((12, 12))
U, S, V = O.svd(M)
dU, dS, dV = rng.standard_normal((12, 12)), rng.standard_normal(12), rng.standard_normal((12, 12))
vec.update(sb_U=U, sb_S=S, sb_V=V, sb_dU=dU, sb_dS=dS, sb_dV=dV, sb_out=O.svd_back(U, S, V, dU, dS, dV))

with open(os.path.join(HERE, "published.json"), "w") as f:
    json.dump(published, f, indent=1)
np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)
print("wrote", len(vec), "arrays")
