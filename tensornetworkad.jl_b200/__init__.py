"""tensornetworkad.jl_b200 -- host-side mirror of TensorNetworkAD.jl's public API over libtnad_b200.so.

Julia is not available in this image, so this Python layer plays the role of the Julia shim
(julia/TensorNetworkAD_b200.jl): same names, argument meaning and error
behaviour as the reference's exports (src/TensorNetworkAD.jl:6-10, ctmrg.jl:7,11, ipeps.jl:1) and
its documented internals (`energy`, `expectationvalue`, `magnetisation`, `trg_svd`, `svd_back`,
`fixedpoint`, `indexperm_symmetrize`).  All tensor arithmetic of the hot path runs in the CUDA
library through the C ABI of include/tnad.h; only the 2^4-element model-tensor builders and the
optimiser loop (the reference's Optim.jl caller, variationalipeps.jl:67-75) are host code, as they
are in the reference.  There is no CPU fallback for the hot path.

Import it as ``import tnad_b200`` (the alias module at the repository root; a directory name with
a dot cannot be imported directly).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from ._lib import (Context, Tape, TnadError, DimensionMismatch, contract_plan, load_library, farray, SIGNATURES,
                   trg_sweep)

__all__ = [
    "trg", "trg_value_and_grad", "num_grad", "ctmrg", "ctmrgstep", "optimiseipeps", "hamiltonian", "model_tensor",
    "mag_tensor", "Ising", "TFIsing", "Heisenberg", "AbstractLattice", "SquareLattice", "CTMRGRuntime",
    "SquareCTMRGRuntime", "IPEPS", "SquareIPEPS", "energy", "energy_and_gradient", "expectationvalue",
    "magnetisation", "magofbeta", "isingbetac", "trg_svd", "svd", "svd_back", "fixedpoint", "StopFunction",
    "indexperm_symmetrize", "diaglocalhamiltonian", "tensorfromclassical", "getchi", "getD", "getd", "gets",
    "Context", "default_context", "DimensionMismatch", "TnadError", "magnetisation_value_and_grad", "dmag_tensor",
    "dmodel_tensor", "trg_sweep", "energy_and_gradient_fixedpoint",
]

_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Lazily created context on the device given by LOCAL_RANK (one process per GPU) or device 0."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("TNAD_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _default_ctx


def _ctx(ctx):
    return ctx if ctx is not None else default_context()


# ------------------------------------------------------------------------------------------------------
# models and host-side input builders (hamiltonianmodels.jl, exampletensors.jl) -- 16-element host math
# ------------------------------------------------------------------------------------------------------
class HamiltonianModel:
    pass


class Ising(HamiltonianModel):
    """hamiltonianmodels.jl:11"""


@dataclass
class TFIsing(HamiltonianModel):
    """hamiltonianmodels.jl:18-20"""
    hx: float


@dataclass
class Heisenberg(HamiltonianModel):
    """hamiltonianmodels.jl:41-46"""
    Jz: float = 1.0
    Jx: float = 1.0
    Jy: float = 1.0


_SX = np.array([[0.0, 1.0], [1.0, 0.0]])
_SY = np.array([[0.0, -1j], [1j, 0.0]])
_SZ = np.array([[1.0, 0.0], [0.0, -1.0]])
_ID2 = np.eye(2)


def hamiltonian(model: HamiltonianModel) -> np.ndarray:
    """hamiltonianmodels.jl:28-33 (TFIsing), :53-59 (Heisenberg)."""
    if isinstance(model, TFIsing):
        h = (-2 * np.einsum("ij,kl->ijkl", _SZ, _SZ) - model.hx / 2 * np.einsum("ij,kl->ijkl", _SX, _ID2)
             - model.hx / 2 * np.einsum("ij,kl->ijkl", _ID2, _SX))
        return np.asfortranarray(h)
    if isinstance(model, Heisenberg):
        h = (model.Jz * np.einsum("ij,kl->ijkl", _SZ, _SZ) - model.Jx * np.einsum("ij,kl->ijkl", _SX, _SX)
             - model.Jy * np.einsum("ij,kl->ijkl", _SY, _SY))
        h = np.einsum("ijcd,kc,ld->ijkl", h, _SX, _SX.conj().T)
        return np.asfortranarray(np.real(h / 2))
    raise TypeError(f"hamiltonian is not defined for {type(model).__name__}")


isingbetac = math.log(1 + math.sqrt(2)) / 2


def _ising_q(beta):
    cb, sb = math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))
    return 1 / math.sqrt(2) * np.array([[cb + sb, cb - sb], [cb - sb, cb + sb]])


def _ising_dq(beta):
    cb, sb = math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))
    dcb, dsb = math.sinh(beta) / (2 * cb), math.cosh(beta) / (2 * sb)
    return 1 / math.sqrt(2) * np.array([[dcb + dsb, dcb - dsb], [dcb - dsb, dcb + dsb]])


def _ising_core(sign):
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 1.0
    a[1, 1, 1, 1] = sign
    return a


def model_tensor(model: HamiltonianModel, beta: float) -> np.ndarray:
    """exampletensors.jl:30-35"""
    if not isinstance(model, Ising):
        raise TypeError("model_tensor is defined for Ising()")
    q = _ising_q(beta)
    return np.asfortranarray(np.einsum("abcd,ai,bj,ck,dl->ijkl", _ising_core(1.0), q, q, q, q))


def dmodel_tensor(model: HamiltonianModel, beta: float) -> np.ndarray:
    """d model_tensor / d beta: the host end of the chain rule Zygote applies in README.md:67-70."""
    if not isinstance(model, Ising):
        raise TypeError("dmodel_tensor is defined for Ising()")
    q, dq = _ising_q(beta), _ising_dq(beta)
    out = np.zeros((2, 2, 2, 2))
    for slot in range(4):
        qs = [q, q, q, q]
        qs[slot] = dq
        out += np.einsum("abcd,ai,bj,ck,dl->ijkl", _ising_core(1.0), *qs)
    return np.asfortranarray(out)


def mag_tensor(model: HamiltonianModel, beta: float) -> np.ndarray:
    """exampletensors.jl:43-48"""
    if not isinstance(model, Ising):
        raise TypeError("mag_tensor is defined for Ising()")
    q = _ising_q(beta)
    return np.asfortranarray(np.einsum("abcd,ai,bj,ck,dl->ijkl", _ising_core(-1.0), q, q, q, q))


def dmag_tensor(model: HamiltonianModel, beta: float) -> np.ndarray:
    """d mag_tensor / d beta (exampletensors.jl:43-48 differentiated by the product rule; the chain-rule factor of
    `Zygote.gradient(beta -> magnetisation(Ising(), beta, chi), beta)`, test/ctmrg.jl:44-46)."""
    if not isinstance(model, Ising):
        raise TypeError("dmag_tensor is defined for Ising()")
    q, dq = _ising_q(beta), _ising_dq(beta)
    out = np.zeros((2, 2, 2, 2))
    for slot in range(4):
        qs = [q, q, q, q]
        qs[slot] = dq
        out += np.einsum("abcd,ai,bj,ck,dl->ijkl", _ising_core(-1.0), *qs)
    return np.asfortranarray(out)


def tensorfromclassical(ham) -> np.ndarray:
    """exampletensors.jl:17-21"""
    import scipy.linalg as sla
    q = np.real(sla.sqrtm(np.exp(np.asarray(ham, dtype=float))))
    return np.asfortranarray(np.einsum("ij,ik,il,im->jklm", q, q, q, q))


def magofbeta(model: HamiltonianModel, beta: float) -> float:
    """exampletensors.jl:77"""
    return (1 - math.sinh(2 * beta) ** -4) ** (1 / 8) if beta > isingbetac else 0.0


def diaglocalhamiltonian(diag) -> np.ndarray:
    """variationalipeps.jl:15-20"""
    diag = np.asarray(diag, dtype=float)
    n = len(diag)
    h, idm = np.diag(diag), np.eye(n)
    return np.asfortranarray(h.reshape(n, n, 1, 1) * idm.reshape(1, 1, n, n)
                             + h.reshape(1, 1, n, n) * idm.reshape(n, n, 1, 1))


def num_grad(f: Callable, x, delta: float = 1e-5):
    """autodiff.jl:44, 58-63"""
    if np.isscalar(x):
        return (f(x + delta / 2) - f(x - delta / 2)) / delta
    x = np.array(x, dtype=float)
    g = np.zeros_like(x)
    for i in np.ndindex(*x.shape):
        xp = x.copy(); xp[i] += delta / 2
        xm = x.copy(); xm[i] -= delta / 2
        g[i] = (f(xp) - f(xm)) / delta
    return g


# ------------------------------------------------------------------------------------------------------
# lattices, iPEPS, CTMRG runtime  (ctmrg.jl:7-37, ipeps.jl:8-24)
# ------------------------------------------------------------------------------------------------------
class AbstractLattice:
    pass


class SquareLattice(AbstractLattice):
    pass


class IPEPS:
    """ipeps.jl:8-12"""

    def __init__(self, bulk, lattice=SquareLattice):
        self.bulk = farray(bulk)
        self.lattice = lattice


class SquareIPEPS(IPEPS):
    """ipeps.jl:16-22: bulk is (d, d, d, d, s); anything else throws DimensionMismatch."""

    def __init__(self, bulk):
        bulk = np.asarray(bulk, dtype=float)
        if bulk.ndim != 5 or not (bulk.shape[0] == bulk.shape[1] == bulk.shape[2] == bulk.shape[3]):
            raise DimensionMismatch(f"size of tensor error, should be `(d, d, d, d, s)`, got {bulk.shape}.")
        super().__init__(bulk, SquareLattice)


def getd(ipeps: IPEPS) -> int:
    return ipeps.bulk.shape[0]


def gets(ipeps: IPEPS) -> int:
    return ipeps.bulk.shape[4]


_SYM_PERMS = [(0, 3, 2, 1, 4), (2, 1, 0, 3, 4), (1, 0, 3, 2, 4), (3, 2, 1, 0, 4)]


def indexperm_symmetrize(ipeps: IPEPS) -> SquareIPEPS:
    """ipeps.jl:32-39 as a stand-alone input builder (inside `energy` it runs on the device)."""
    x = ipeps.bulk
    for p in _SYM_PERMS:
        x = x + np.transpose(x, p)
    return SquareIPEPS(x / np.linalg.norm(x))


class CTMRGRuntime:
    """ctmrg.jl:23-32: holder of bulk (D^4), corner (chi x chi), edge (chi x D x chi)."""

    def __init__(self, bulk, corner, edge, lattice=SquareLattice):
        self.bulk, self.corner, self.edge = farray(bulk), farray(corner), farray(edge)
        self.lattice = lattice


def SquareCTMRGRuntime(bulk, env_or_corner, chi_or_edge, rng=None, ctx=None) -> CTMRGRuntime:
    """ctmrg.jl:34 (bulk, corner, edge) and ctmrg.jl:62-86 (bulk, env, chi) with env in {"raw", "random"}."""
    bulk = farray(bulk)
    if isinstance(env_or_corner, str):
        chi = int(chi_or_edge)
        D = bulk.shape[0]
        if env_or_corner == "raw":
            corner, edge = _ctx(ctx).ctmrg_init_raw(bulk, chi)         # ctmrg.jl:74-86 on the device
        elif env_or_corner == "random":                                # ctmrg.jl:66-72 (host RNG, as in Julia)
            rng = rng or np.random.default_rng()
            corner = rng.standard_normal((chi, chi))
            edge = rng.standard_normal((chi, D, chi))
            corner = corner + corner.T
            edge = edge + np.transpose(edge, (2, 1, 0))
        else:
            raise ValueError("env must be 'raw' or 'random'")
        return CTMRGRuntime(bulk, corner, edge)
    return CTMRGRuntime(bulk, env_or_corner, chi_or_edge)


def getchi(rt: CTMRGRuntime) -> int:
    return rt.corner.shape[0]


def getD(rt: CTMRGRuntime) -> int:
    return rt.bulk.shape[0]


# ------------------------------------------------------------------------------------------------------
# fixedpoint.jl (generic host loop, kept for API parity; `ctmrg` runs its loop inside the library)
# ------------------------------------------------------------------------------------------------------
def fixedpoint(f: Callable, guess, stopfun: Callable):
    """fixedpoint.jl:11-15: the seed itself is the first state tested."""
    state = guess
    while True:
        if stopfun(state):
            return state
        state = f(state)


class StopFunction:
    """fixedpoint.jl:17-41"""

    def __init__(self, oldvals, counter, tol, maxit):
        self.oldvals, self.counter, self.tol, self.maxit = oldvals, counter, tol, maxit

    def __call__(self, state):
        self.counter += 1
        if self.counter > self.maxit:
            return True
        vals = state[1]
        with np.errstate(invalid="ignore"):
            diff = np.linalg.norm(np.asarray(vals) - np.asarray(self.oldvals))
        if diff <= self.tol:
            return True
        self.oldvals = vals
        return False


# ------------------------------------------------------------------------------------------------------
# hot path: every function below is one C-ABI call
# ------------------------------------------------------------------------------------------------------
def svd(A, ctx=None):
    """LinearAlgebra.svd as used at trg.jl:36 / ctmrg.jl:136 -> (U, S, V), A = U diag(S) V'."""
    return _ctx(ctx).svd(A)


def trg_svd(t, dmax, tol, ctx=None):
    """trg.jl:33-44"""
    return _ctx(ctx).trg_svd(t, dmax, tol)


def svd_back(U, S, V, dU, dS, dV, eta=1e-40, ctx=None):
    """trg.jl:72-105 (real case)"""
    return _ctx(ctx).svd_back(U, S, V, dU, dS, dV, eta)


def trg(a, chi, niter, tol: float = 1e-16, ctx=None) -> float:
    """trg.jl:13-30"""
    return _ctx(ctx).trg_forward(a, chi, niter, tol)


def trg_value_and_grad(a, chi, niter, tol: float = 1e-16, ctx=None):
    """lnZ and d lnZ/d a: what `Zygote.gradient(a -> trg(a, chi, niter), a)` returns (README.md:67-70)."""
    c = _ctx(ctx)
    lnz, tape = c.trg_forward(a, chi, niter, tol, want_tape=True)
    try:
        g = c.trg_backward(tape, 1.0)
    finally:
        tape.free()
    return lnz, g


def ctmrgstep(rt: CTMRGRuntime, vals=None, ctx=None):
    """ctmrg.jl:126-153 -> (CTMRGRuntime, vals)"""
    co, ed, v = _ctx(ctx).ctmrgstep(rt.bulk, rt.corner, rt.edge)
    return CTMRGRuntime(rt.bulk, co, ed), v


def ctmrg(rt: CTMRGRuntime, tol: float, maxit: int, ctx=None) -> CTMRGRuntime:
    """ctmrg.jl:110-117"""
    co, ed, vals, steps = _ctx(ctx).ctmrg(rt.bulk, rt.corner, rt.edge, tol, maxit)
    out = CTMRGRuntime(rt.bulk, co, ed)
    out.vals, out.steps = vals, steps
    return out


def expectationvalue(h, ap, rt: CTMRGRuntime, ctx=None) -> float:
    """variationalipeps.jl:49-56"""
    return _ctx(ctx).expectationvalue(h, ap, rt.corner, rt.edge)


def _bulk_of(ipeps):
    return ipeps.bulk if isinstance(ipeps, IPEPS) else np.asarray(ipeps, dtype=float)


def energy(h, ipeps, chi: int, tol: float, maxit: int, ctx=None) -> float:
    """variationalipeps.jl:28-40"""
    return _ctx(ctx).energy(h, _bulk_of(ipeps), chi, tol, maxit, grad=False)


def energy_and_gradient(h, ipeps, chi: int, tol: float, maxit: int, ctx=None):
    """energy and `Zygote.gradient(x -> energy(h, x; ...), ipeps)[1].bulk` in one call (forward shared)."""
    return _ctx(ctx).energy(h, _bulk_of(ipeps), chi, tol, maxit, grad=True)


def energy_and_gradient_fixedpoint(h, ipeps, chi: int, tol: float, maxit: int, bwd_tol: float = 1e-12, bwd_maxit: int = 500,
                                   ctx=None):
    """Opt-in: energy with the implicit fixed-point gradient (tnad_energy_fixedpoint); equals energy_and_gradient for a
    converged CTMRG, needs one step record instead of a tape over all iterations."""
    return _ctx(ctx).energy_fixedpoint(h, _bulk_of(ipeps), chi, tol, maxit, bwd_tol, bwd_maxit)


def magnetisation(model: HamiltonianModel, beta: float, chi: int, rng=None, tol=1e-6, maxit=100, env="random",
                  ctx=None) -> float:
    """exampletensors.jl:57-69 (the reference hard-codes :random, tol=1e-6, maxit=100)."""
    c = _ctx(ctx)
    a, m = model_tensor(model, beta), mag_tensor(model, beta)
    rt = SquareCTMRGRuntime(a, env, chi, rng=rng, ctx=c)
    rt = ctmrg(rt, tol, maxit, ctx=c)
    return c.magnetisation_readout(a, m, rt.corner, rt.edge)


def magnetisation_value_and_grad(model: HamiltonianModel, beta: float, chi: int, rng=None, tol=1e-6, maxit=100,
                                 env="random", ctx=None):
    """magnetisation and `Zygote.gradient(beta -> magnetisation(model, beta, chi), beta)[1]` (test/ctmrg.jl:44-46).

    The environment initialisation is constant under AD (autodiff.jl:5); the gradient flows through the read-out
    (tnad_magnetisation_backward), through every executed ctmrgstep (tnad_ctmrg_backward, bulk = a) and, on the
    host, through model_tensor / mag_tensor."""
    c = _ctx(ctx)
    a, m = model_tensor(model, beta), mag_tensor(model, beta)
    rt = SquareCTMRGRuntime(a, env, chi, rng=rng, ctx=c)
    co, ed, _, _, tape = c.ctmrg(rt.bulk, rt.corner, rt.edge, tol, maxit, want_tape=True)
    try:
        y = c.magnetisation_readout(a, m, co, ed)
        da, dm, dc, de = c.magnetisation_backward(a, m, co, ed, 1.0)
        da = da + c.ctmrg_backward(tape, dc, de)
    finally:
        tape.free()
    return y, float(np.sum(da * dmodel_tensor(model, beta)) + np.sum(dm * dmag_tensor(model, beta)))


def optimiseipeps(ipeps: IPEPS, h, chi: int, tol: float, maxit: int, optimargs: Optional[dict] = None,
                  optimmethod: str = "L-BFGS-B", ctx=None):
    """variationalipeps.jl:67-75: L-BFGS (m = 20) over energy / its gradient.

    The optimiser is the host-side caller of the hot path (Optim.jl in the reference, SciPy here);
    value and gradient come from one fused `tnad_energy` call.  `optimargs` maps onto
    scipy.optimize.minimize options (`f_tol` -> ftol, `iterations` -> maxiter)."""
    from scipy.optimize import minimize
    c = _ctx(ctx)
    shape = ipeps.bulk.shape
    opts = {"maxcor": 20, "gtol": 1e-8, "ftol": 1e-12, "maxiter": 1000}
    for k, v in (optimargs or {}).items():
        opts[{"f_tol": "ftol", "g_tol": "gtol", "iterations": "maxiter"}.get(k, k)] = v

    def fg(x):
        e, g = c.energy(h, np.reshape(x, shape, order="F"), chi, tol, maxit, grad=True)
        return e, np.reshape(g, -1, order="F")

    res = minimize(fg, np.reshape(ipeps.bulk, -1, order="F"), jac=True, method=optimmethod, options=opts)
    res.minimizer = np.reshape(res.x, shape, order="F")
    res.minimum = float(res.fun)
    return res
