"""Generates tests/golden/large.npz: oracle outputs at the sizes bench.py measures (BASELINE configs 2, 4, 5 and
the TRG chi=64 part of the metric).  Minutes of CPU time (LAPACK dgesdd of 2048^2, 4096^2 and 6400^2 matrices), so
the vectors are frozen here instead of being recomputed by the GPU tests.

The reference is pure Julia and cannot run in this image; like vectors.npz these come from the pinned restatement
oracle/tnad_oracle.py.  Inputs are regenerated from seeds by the tests (NumPy's default_rng streams are
platform-independent), only outputs are stored.

    python tests/golden/make_golden_large.py [c4] [c2] [trg64] [c5a]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tnad_oracle as O  # noqa: E402

OUT = os.path.join(HERE, "large.npz")


def load():
    return dict(np.load(OUT)) if os.path.exists(OUT) else {}


def save(vec):
    np.savez_compressed(OUT, **vec)


def c4_input():
    """BASELINE configs[3]: A = indexperm_symmetrize(standard_normal(4,4,4,4,2), seed 0), Heisenberg h."""
    return O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))


def c5a_input():
    """BASELINE configs[4]: d=5, chi=256; one ctmrgstep from a seeded :random environment."""
    A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((5, 5, 5, 5, 2)))
    _, a = O.double_layer(A)
    c0, e0 = O.init_random(a, 256, np.random.default_rng(1))
    return a, c0, e0


def c5a_probes(chi, D):
    rng = np.random.default_rng(7)
    return rng.standard_normal(chi), rng.standard_normal(D), rng.standard_normal(chi)


def main(which):
    vec = load()
    h = O.hamiltonian_heisenberg()
    if "c4" in which:
        t = time.time()
        A = c4_input()
        info = {}
        e, g = O.energy_value_and_grad(h, A, 128, 0.0, 3, info=info)
        vec.update(c4_e=e, c4_grad=g, c4_steps=info["nsteps"], c4_vals=info["vals"], c4_cfg=np.array([4, 128, 3]))
        print("c4", e, info["nsteps"], time.time() - t, flush=True)
        save(vec)
    if "c2" in which:
        for beta in (0.3, 0.5):
            t = time.time()
            a, m = O.model_tensor_ising(beta), O.mag_tensor_ising(beta)
            c0, e0 = O.init_raw(a, 64)
            c, ed, vals, ns = O.ctmrg(a, c0, e0, 1e-10, 5000)
            vec[f"c2_raw_{beta}_vals"] = vals
            vec[f"c2_raw_{beta}_steps"] = ns
            vec[f"c2_raw_{beta}_mag"] = O.magnetisation_readout(a, m, c, ed)
            # noise floor of the step count / spectrum: the same run with LAPACK's other SVD driver
            _, _, vals2, ns2 = O.ctmrg(a, c0, e0, 1e-10, 5000, driver="gesvd")
            vec[f"c2_raw_{beta}_steps_gesvd"] = ns2
            vec[f"c2_raw_{beta}_vals_spread"] = np.abs(vals - vals2).max()
            print("c2 raw", beta, ns, ns2, vec[f"c2_raw_{beta}_mag"], vec[f"c2_raw_{beta}_vals_spread"], time.time() - t, flush=True)
        for beta in (0.3, 0.5):      # :random environment from a host-seeded generator (ctmrg.jl:66-72)
            a, m = O.model_tensor_ising(beta), O.mag_tensor_ising(beta)
            c0, e0 = O.init_random(a, 64, np.random.default_rng(3))
            c, ed, vals, ns = O.ctmrg(a, c0, e0, 1e-10, 5000)
            vec[f"c2_random_{beta}_vals"] = vals
            vec[f"c2_random_{beta}_steps"] = ns
            vec[f"c2_random_{beta}_mag"] = O.magnetisation_readout(a, m, c, ed)
            print("c2 random", beta, ns, vec[f"c2_random_{beta}_mag"], O.magofbeta(beta), flush=True)
        save(vec)
    if "trg64" in which:
        t = time.time()
        beta = 0.44
        # the bond dimension saturates at 64 in iteration 5; iterations 6 and 7 decompose full 4096 x 4096 matrices
        lnz, ga = O.trg_value_and_grad(O.model_tensor_ising(beta), 64, 7)
        vec["trg64_lnz"] = lnz
        vec["trg64_dbeta"] = float(np.sum(ga * O.dmodel_tensor_ising(beta)))
        vec["trg64_cfg"] = np.array([beta, 64, 7])
        print("trg64", lnz, vec["trg64_dbeta"], time.time() - t, flush=True)
        save(vec)
    if "c5a" in which:
        t = time.time()
        a, c0, e0 = c5a_input()
        c, ed, vals = O.ctmrgstep(a, c0, e0, signfix=True)
        w1, w2, w3 = c5a_probes(256, 25)
        vec.update(c5a_vals=vals, c5a_corner_w=c @ w1, c5a_edge_w=np.einsum("ijk,i,k->j", ed, w1, w3),
                   c5a_edge_w2=np.einsum("ijk,j,k->i", ed, w2, w3), c5a_corner_diag=np.diag(c).copy())
        print("c5a", vals[:4], time.time() - t, flush=True)
        save(vec)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c4", "c2", "trg64", "c5a"])
