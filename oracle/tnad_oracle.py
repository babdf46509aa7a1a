"""CPU oracle for the TRG / CTMRG / iPEPS-energy hot path of TensorNetworkAD.jl.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or as the reported CPU baseline.

It is a NumPy/SciPy restatement of the reference's Julia sources (the reference
cannot be executed here: no Julia in the image).  All arithmetic of the reference
lives in un-vendored, un-pinned third-party packages (OMEinsum -> BLAS dgemm,
LinearAlgebra.svd -> LAPACK dgesdd, Zygote reverse mode); this file restates
their published semantics with the same LAPACK driver family (``gesdd``) and
standard einsum adjoints, and follows the reference's own call sites:

  src/trg.jl:13-44          trg, trg_svd
  src/trg.jl:72-105         svd_back (literal)
  src/ctmrg.jl:66-86        _initializect_square (:random / :raw)
  src/ctmrg.jl:110-153      ctmrg, ctmrgstep
  src/fixedpoint.jl:11-41   fixedpoint, StopFunction (counter starts at -1, ctmrg.jl:114)
  src/ipeps.jl:32-39        indexperm_symmetrize
  src/variationalipeps.jl:15-56  diaglocalhamiltonian, energy, expectationvalue
  src/autodiff.jl:23-29     norm pullback; :4-5 what is constant under AD; :44 num_grad
  src/exampletensors.jl:17-77    Ising tensors, magnetisation, Onsager
  src/hamiltonianmodels.jl:28-59 TFIsing / Heisenberg two-site operators

PARITY PINNED: tests/test_oracle_golden.py checks this file against every exact
golden value the reference publishes for the path (test/trg.jl:18, README.md:60,
README.md:70, docs/src/userguide.md:28), its analytic tests (Onsager
magnetisation test/ctmrg.jl:37-42; non-interacting energies
test/variationalipeps.jl:14-25), the README Hamiltonian print-out
(README.md:83-100) and finite differences of every adjoint (test/svd.jl,
test/trg.jl:19, test/variationalipeps.jl:121-134).  Nothing in the reference
pins results at the d=4, chi=128 size: there parity is against this oracle only.

Index convention: arrays carry the Julia index order; every reshape is
column-major (``order='F'``), as in Julia.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.linalg as sla

__all__ = [
    "model_tensor_ising", "mag_tensor_ising", "tensorfromclassical", "magofbeta", "ISING_BETA_C",
    "hamiltonian_heisenberg", "hamiltonian_tfising", "diaglocalhamiltonian",
    "svd", "trg_svd", "svd_back", "trg", "trg_value_and_grad", "trg_dbeta",
    "init_raw", "init_random", "ctmrgstep", "ctmrgstep_literal", "ctmrg", "ctmrg_backward",
    "indexperm_symmetrize", "double_layer", "expectationvalue", "energy", "energy_value_and_grad",
    "magnetisation_readout", "magnetisation", "num_grad", "canonical_gauge", "magnetisation_readout_back",
    "magnetisation_value_and_dbeta", "dmag_tensor_ising",
]

ISING_BETA_C = math.log(1 + math.sqrt(2)) / 2  # exampletensors.jl:2


def rsh(x, shape):
    """Column-major reshape (Julia `reshape`)."""
    return np.reshape(x, shape, order="F")


def es(spec, *ops):
    """einsum through BLAS (pairwise specs only are used on the fast path)."""
    return np.einsum(spec, *ops, optimize=True)


# ----------------------------------------------------------------------------
# host-side model tensors  (exampletensors.jl, hamiltonianmodels.jl)
# ----------------------------------------------------------------------------
def _ising_q(beta):
    cb, sb = math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))
    return 1 / math.sqrt(2) * np.array([[cb + sb, cb - sb], [cb - sb, cb + sb]])


def model_tensor_ising(beta):
    """exampletensors.jl:30-35"""
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 1.0
    a[1, 1, 1, 1] = 1.0
    q = _ising_q(beta)
    return np.einsum("abcd,ai,bj,ck,dl->ijkl", a, q, q, q, q)


def dmodel_tensor_ising(beta):
    """d model_tensor / d beta (closed form; used for the chain rule dlnZ/dbeta)."""
    cb, sb = math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))
    dcb = math.sinh(beta) / (2 * cb)
    dsb = math.cosh(beta) / (2 * sb)
    q = _ising_q(beta)
    dq = 1 / math.sqrt(2) * np.array([[dcb + dsb, dcb - dsb], [dcb - dsb, dcb + dsb]])
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 1.0
    a[1, 1, 1, 1] = 1.0
    out = np.zeros((2, 2, 2, 2))
    for slot in range(4):
        qs = [q, q, q, q]
        qs[slot] = dq
        out += np.einsum("abcd,ai,bj,ck,dl->ijkl", a, *qs)
    return out


def mag_tensor_ising(beta):
    """exampletensors.jl:43-48"""
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 1.0
    a[1, 1, 1, 1] = -1.0
    q = _ising_q(beta)
    return np.einsum("abcd,ai,bj,ck,dl->ijkl", a, q, q, q, q)


def tensorfromclassical(ham):
    """exampletensors.jl:17-21 (sqrt is the matrix square root)."""
    w = np.exp(np.asarray(ham, dtype=float))
    q = np.real(sla.sqrtm(w))
    return np.einsum("ij,ik,il,im->jklm", q, q, q, q)


def magofbeta(beta):
    """exampletensors.jl:77 (Onsager)."""
    return (1 - math.sinh(2 * beta) ** -4) ** (1 / 8) if beta > ISING_BETA_C else 0.0


_SX = np.array([[0.0, 1.0], [1.0, 0.0]])
_SY = np.array([[0.0, -1j], [1j, 0.0]])
_SZ = np.array([[1.0, 0.0], [0.0, -1.0]])
_ID = np.eye(2)


def hamiltonian_tfising(hx):
    """hamiltonianmodels.jl:28-33"""
    return (-2 * np.einsum("ij,kl->ijkl", _SZ, _SZ)
            - hx / 2 * np.einsum("ij,kl->ijkl", _SX, _ID)
            - hx / 2 * np.einsum("ij,kl->ijkl", _ID, _SX))


def hamiltonian_heisenberg(Jz=1.0, Jx=1.0, Jy=1.0):
    """hamiltonianmodels.jl:53-59 (sublattice rotation by sigma_x on site 2, /2, real part)."""
    h = (Jz * np.einsum("ij,kl->ijkl", _SZ, _SZ)
         - Jx * np.einsum("ij,kl->ijkl", _SX, _SX)
         - Jy * np.einsum("ij,kl->ijkl", _SY, _SY))
    h = np.einsum("ijcd,kc,ld->ijkl", h, _SX, _SX.T.conj())
    return np.ascontiguousarray(np.real(h / 2))


def diaglocalhamiltonian(diag):
    """variationalipeps.jl:15-20"""
    diag = np.asarray(diag, dtype=float)
    n = len(diag)
    h = np.diag(diag)
    idm = np.eye(n)
    return h.reshape(n, n, 1, 1) * idm.reshape(1, 1, n, n) + h.reshape(1, 1, n, n) * idm.reshape(n, n, 1, 1)


def num_grad(f, x, delta=1e-5):
    """autodiff.jl:44, 58-63"""
    if np.isscalar(x):
        return (f(x + delta / 2) - f(x - delta / 2)) / delta
    x = np.array(x, dtype=float)
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp = x.copy(); xp[i] += delta / 2
        xm = x.copy(); xm[i] -= delta / 2
        g[i] = (f(xp) - f(xm)) / delta
    return g


# ----------------------------------------------------------------------------
# SVD and its adjoint  (trg.jl:33-105)
# ----------------------------------------------------------------------------
def svd(A, driver="gesdd"):
    """LinearAlgebra.svd: thin SVD, returns U, S, V with A = U diag(S) V^H (LAPACK gesdd)."""
    U, S, Vh = sla.svd(A, full_matrices=False, lapack_driver=driver)
    return U, S, Vh.conj().T


def canonical_gauge(U, V=None):
    """Canonical column signs: the largest-magnitude entry of every column of U
    (first one on ties) is made positive; V, when given, follows column by column so U diag(S) V^H is unchanged.
    Not part of the reference (its gauge is whatever LAPACK returns); used by the tests so that gauge-dependent
    tensors (corner, edge, U itself) of two implementations can be compared entry by entry."""
    idx = np.argmax(np.abs(U), axis=0)
    sg = np.sign(U[idx, np.arange(U.shape[1])])
    sg = np.where(sg == 0, 1.0, sg)
    return (U * sg[None, :], None if V is None else V * sg[None, :])


def rank_rule(S, dmax, tol):
    """trg.jl:37: min(searchsortedfirst(s, tol, rev=true), dmax, length(s))."""
    idx = len(S) + 1
    for i, x in enumerate(S):
        if x <= tol:
            idx = i + 1
            break
    return min(idx, dmax, len(S))


def trg_svd(t, dmax, tol, driver="gesdd"):
    """trg.jl:33-44. returns u (d1,d2,k), v (k,d3,d4) and the tape (U,S,V,k)."""
    d1, d2, d3, d4 = t.shape
    tmat = rsh(t, (d1 * d2, d3 * d4))
    U, S, V = svd(tmat, driver)
    k = rank_rule(S, dmax, tol)
    sq = np.sqrt(S[:k])
    u = rsh(U[:, :k] * sq[None, :], (d1, d2, k))
    v = rsh(sq[:, None] * V[:, :k].conj().T, (k, d3, d4))
    return u, v, (U, S, V, k)


def svd_back(U, S, V, dU, dS, dV, eta=1e-40):
    """trg.jl:72-105, literal. Any of dU/dS/dV may be None. A = U diag(S) V^H."""
    if dU is None and dS is None and dV is None:
        return None
    S2 = S ** 2
    Sinv = S / (S2 + eta)
    F = S2[None, :] - S2[:, None]
    F = F / (F ** 2 + eta)
    res = 0
    if dU is not None:
        UdU = U.conj().T @ dU
        J = F * UdU
        res = res + (J + J.conj().T) * S[None, :] + np.diag(1j * np.imag(np.diag(UdU)) * Sinv)
    if dV is not None:
        VdV = V.conj().T @ dV
        K = F * VdV
        res = res + S[:, None] * (K + K.conj().T)
    if dS is not None:
        res = res + np.diag(dS)
    res = U @ res @ V.conj().T
    if dU is not None and U.shape[0] != U.shape[1]:
        res = res + ((dU - U @ (U.conj().T @ dU)) * Sinv[None, :]) @ V.conj().T
    if dV is not None and V.shape[0] != V.shape[1]:
        res = res + (U * Sinv[None, :]) @ (dV.conj().T - (dV.conj().T @ V) @ V.conj().T)
    if not np.iscomplexobj(U):
        res = np.real(res)
    return res


def _trg_svd_back(tape, du, dv, eta=1e-40):
    """Reverse of trg.jl:37-41 (slice, scale by sqrt(s)) followed by svd_back.

    du: cotangent of u (d1,d2,k); dv: cotangent of v (k,d3,d4). Returns d(tmat)."""
    U, S, V, k = tape
    m, n = U.shape[0], V.shape[0]
    sq = np.sqrt(S[:k])
    dum = rsh(du, (m, k))
    dvm = rsh(dv, (k, n))
    dU = np.zeros_like(U)
    dV = np.zeros_like(V)
    dU[:, :k] = dum * sq[None, :]
    dV[:, :k] = (sq[:, None] * dvm).T          # dVt[1:k,:] = sqrt(s) dv ; dV = dVt'
    dsq = np.einsum("ij,ij->j", U[:, :k], dum) + np.einsum("ji,ij->i", V[:, :k], dvm)
    dS = np.zeros_like(S)
    dS[:k] = dsq / (2 * sq)
    return svd_back(U, S, V, dU, dS, dV, eta)


# ----------------------------------------------------------------------------
# TRG  (trg.jl:13-30)
# ----------------------------------------------------------------------------
def trg(a, chi, niter, tol=1e-16, driver="gesdd", tape=None):
    """trg.jl:13-30 forward. If `tape` is a list, per-iteration records are appended."""
    a = np.asarray(a, dtype=float)
    lnZ = 0.0
    for n in range(1, niter + 1):
        maxval = np.max(np.abs(a))
        a_in = a
        a = a / maxval
        lnZ += 2.0 ** (1 - n) * math.log(maxval)
        dr_ul = np.transpose(a, (2, 1, 0, 3))     # ein"urdl -> drul"
        ld_ru = np.transpose(a, (3, 2, 1, 0))     # ein"urdl -> ldru"
        dr, ul, t1 = trg_svd(dr_ul, chi, tol, driver)
        ld, ru, t2 = trg_svd(ld_ru, chi, tol, driver)
        a_new = es("npu,por,dom,lmn->urdl", dr, ld, ul, ru)
        if tape is not None:
            tape.append(dict(a_in=a_in, maxval=maxval, a=a, dr=dr, ul=ul, ld=ld, ru=ru, t1=t1, t2=t2))
        a = a_new
    trace = np.einsum("ijij->", a)
    lnZ += math.log(trace) / 2.0 ** niter
    if tape is not None:
        tape.append(dict(a_final=a, trace=trace))
    return lnZ


def trg_value_and_grad(a, chi, niter, tol=1e-16, driver="gesdd", eta=1e-40):
    """lnZ and d lnZ / d a as Zygote + the reference's rules compute it."""
    tape = []
    lnZ = trg(a, chi, niter, tol, driver, tape)
    fin = tape.pop()
    af = fin["a_final"]
    abar = np.zeros_like(af)
    w = 1.0 / (2.0 ** niter * fin["trace"])
    for i in range(af.shape[0]):
        for j in range(af.shape[1]):
            abar[i, j, i, j] += w
    for n in range(niter, 0, -1):
        rec = tape[n - 1]
        dr, ld, ul, ru = rec["dr"], rec["ld"], rec["ul"], rec["ru"]
        # adjoints of ein"npu,por,dom,lmn->urdl"
        ddr = es("urdl,por,dom,lmn->npu", abar, ld, ul, ru)
        dld = es("urdl,npu,dom,lmn->por", abar, dr, ul, ru)
        dul = es("urdl,npu,por,lmn->dom", abar, dr, ld, ru)
        dru = es("urdl,npu,por,dom->lmn", abar, dr, ld, ul)
        dt1 = _trg_svd_back(rec["t1"], ddr, dul, eta)   # d(dr_ul matrix)
        dt2 = _trg_svd_back(rec["t2"], dld, dru, eta)
        a = rec["a"]
        du_, dr_, dd_, dl_ = a.shape
        d_drul = rsh(dt1, (dd_, dr_, du_, dl_))
        d_ldru = rsh(dt2, (dl_, dd_, dr_, du_))
        da = np.transpose(d_drul, (2, 1, 0, 3)) + np.transpose(d_ldru, (3, 2, 1, 0))
        # a = a_in / maxval ; lnZ += 2^(1-n) log(maxval) ; maxval = maximum(abs.(a_in))
        maxval, a_in = rec["maxval"], rec["a_in"]
        da_in = da / maxval
        dmax = -np.sum(da * a_in) / maxval ** 2 + 2.0 ** (1 - n) / maxval
        idx = np.unravel_index(np.argmax(np.abs(a_in)), a_in.shape)
        da_in[idx] += dmax * np.sign(a_in[idx])
        abar = da_in
    return lnZ, abar


def trg_dbeta(beta, chi, niter, tol=1e-16, driver="gesdd"):
    """README.md:67-70: gradient(beta -> trg(model_tensor(Ising(), beta), chi, niter))."""
    lnZ, abar = trg_value_and_grad(model_tensor_ising(beta), chi, niter, tol, driver)
    return lnZ, float(np.sum(abar * dmodel_tensor_ising(beta)))


# ----------------------------------------------------------------------------
# CTMRG  (ctmrg.jl, fixedpoint.jl)
# ----------------------------------------------------------------------------
def init_raw(bulk, chi):
    """ctmrg.jl:74-86"""
    D = bulk.shape[0]
    corner = np.zeros((chi, chi), dtype=bulk.dtype)
    edge = np.zeros((chi, D, chi), dtype=bulk.dtype)
    cinit = np.einsum("ijkl->ij", bulk)
    tinit = np.einsum("ijkl->ijk", bulk)
    m = min(D, chi)
    corner[:m, :m] = cinit[:m, :m]
    edge[:m, :, :m] = tinit[:m, :, :m]
    return corner, edge


def init_random(bulk, chi, rng):
    """ctmrg.jl:66-72 with a caller-supplied generator (the reference uses Julia's global RNG)."""
    D = bulk.shape[0]
    corner = rng.standard_normal((chi, chi))
    edge = rng.standard_normal((chi, D, chi))
    corner = corner + corner.T
    edge = edge + np.transpose(edge, (2, 1, 0))
    return corner, edge


def ctmrgstep_literal(bulk, corner, edge, driver="gesdd"):
    """ctmrg.jl:126-153 with the reference's own einsum strings (including `tp`)."""
    D, chi = bulk.shape[0], corner.shape[0]
    cp = np.einsum("ad,iba,dcl,jkcb->ijlk", corner, edge, edge, bulk, optimize=True)
    tp = np.einsum("iam,jkla->ijklm", edge, bulk, optimize=True)
    cpmat = rsh(cp, (chi * D, chi * D))
    cpmat = cpmat + cpmat.T
    u, s, v = svd(cpmat, driver)
    z = rsh(u[:, :chi], (chi, D, chi))
    corner = np.einsum("abcd,abi,cdj->ij", cp, z, z, optimize=True)
    edge = np.einsum("abjcd,abi,dck->ijk", tp, z, z, optimize=True)
    vals = s / s[0]
    corner = corner + corner.T
    edge = edge + np.transpose(edge, (2, 1, 0))
    corner = corner / np.linalg.norm(corner)
    edge = edge / np.linalg.norm(edge)
    return corner, edge, vals


def ctmrgstep(bulk, corner, edge, driver="gesdd", tape=None, signfix=False):
    """ctmrg.jl:126-153 in the optimal pairwise order, never materialising `tp`.
    signfix=True puts U into the canonical gauge (canonical_gauge) before the projection."""
    D, chi = bulk.shape[0], corner.shape[0]
    n = chi * D
    X1 = es("iba,ad->ibd", edge, corner)
    X2 = es("ibd,dcl->ibcl", X1, edge)
    cp = es("ibcl,jkcb->ijlk", X2, bulk)
    CP = rsh(cp, (n, n))
    M = CP + CP.T
    U, S, V = svd(M, driver)
    if signfix:
        U, V = canonical_gauge(U, V)
    Z = U[:, :chi]
    z = rsh(Z, (chi, D, chi))
    W = CP @ Z
    c1 = Z.T @ W
    Y1 = es("abi,aed->ibed", z, edge)
    Y2 = es("ibed,bjce->ijcd", Y1, bulk)
    e1 = es("ijcd,dck->ijk", Y2, z)
    vals = S / S[0]
    c2 = c1 + c1.T
    e2 = e1 + np.transpose(e1, (2, 1, 0))
    nc, ne = np.linalg.norm(c2), np.linalg.norm(e2)
    c3, e3 = c2 / nc, e2 / ne
    if tape is not None:
        tape.append(dict(corner=corner, edge=edge, X1=X1, X2=X2, CP=CP, U=U, S=S, V=V, z=z,
                         Y1=Y1, Y2=Y2, c2=c2, e2=e2, nc=nc, ne=ne))
    return c3, e3, vals


def _norm_back(xbar, x2, nrm):
    """x3 = x2/||x2|| with the reference's norm pullback (autodiff.jl:23-29)."""
    return xbar / nrm - np.sum(xbar * x2) * x2 / nrm ** 3


def ctmrgstep_backward(bulk, rec, cbar3, ebar3, eta=1e-40):
    """Reverse of ctmrgstep. Returns (dbulk, dcorner_in, dedge_in)."""
    corner, edge = rec["corner"], rec["edge"]
    chi, D = corner.shape[0], bulk.shape[0]
    n = chi * D
    U, S, V, z, CP = rec["U"], rec["S"], rec["V"], rec["z"], rec["CP"]
    Z = rsh(z, (n, chi))
    # normalise + symmetrise
    cbar2 = _norm_back(cbar3, rec["c2"], rec["nc"])
    ebar2 = _norm_back(ebar3, rec["e2"], rec["ne"])
    cbar1 = cbar2 + cbar2.T
    ebar1 = ebar2 + np.transpose(ebar2, (2, 1, 0))
    # corner projection c1 = Z' CP Z
    CPbar = Z @ cbar1 @ Z.T
    Zbar = CP @ (Z @ cbar1.T) + CP.T @ (Z @ cbar1)
    # edge projection: Y1 = z*edge, Y2 = Y1*bulk, e1 = Y2*z
    Y1, Y2 = rec["Y1"], rec["Y2"]
    Y2bar = es("ijk,dck->ijcd", ebar1, z)
    zbar = es("ijcd,ijk->dck", Y2, ebar1)
    Y1bar = es("ijcd,bjce->ibed", Y2bar, bulk)
    bulkbar = es("ibed,ijcd->bjce", Y1, Y2bar)
    zbar = zbar + es("ibed,aed->abi", Y1bar, edge)
    edgebar = es("abi,ibed->aed", z, Y1bar)
    Zbar = Zbar + rsh(zbar, (n, chi))
    # svd (square; only dU[:, :chi] non-zero; dS = dV = nothing)
    dU = np.zeros_like(U)
    dU[:, :chi] = Zbar
    Mbar = svd_back(U, S, V, dU, None, None, eta)
    CPbar = CPbar + Mbar + Mbar.T
    # grow: cp = X2 * bulk ; X2 = X1 * edge ; X1 = edge * corner
    cpbar = rsh(CPbar, (chi, D, chi, D))
    X1, X2 = rec["X1"], rec["X2"]
    X2bar = es("ijlk,jkcb->ibcl", cpbar, bulk)
    bulkbar = bulkbar + es("ibcl,ijlk->jkcb", X2, cpbar)
    X1bar = es("ibcl,dcl->ibd", X2bar, edge)
    edgebar = edgebar + es("ibd,ibcl->dcl", X1, X2bar)
    edgebar = edgebar + es("ibd,ad->iba", X1bar, corner)
    cornerbar = es("iba,ibd->ad", edge, X1bar)
    return bulkbar, cornerbar, edgebar


def ctmrg(bulk, corner, edge, tol, maxit, driver="gesdd", tape=None, step=None):
    """ctmrg.jl:110-117 + fixedpoint.jl:11-41. Returns corner, edge, vals, number of steps run."""
    step = step or ctmrgstep
    D, chi = bulk.shape[0], corner.shape[0]
    oldvals = np.full(chi * D, np.inf)
    vals = oldvals
    counter = -1                       # ctmrg.jl:114
    nsteps = 0
    while True:
        counter += 1                   # fixedpoint.jl:32
        if counter > maxit:
            break
        with np.errstate(invalid="ignore"):
            diff = np.linalg.norm(vals - oldvals)
        if diff <= tol:                # NaN <= tol is False
            break
        oldvals = vals
        if tape is not None:
            corner, edge, vals = step(bulk, corner, edge, driver, tape)
        else:
            corner, edge, vals = step(bulk, corner, edge, driver)
        nsteps += 1
    return corner, edge, vals, nsteps


def ctmrg_backward(bulk, tape, cbar, ebar, eta=1e-40):
    """Unrolled reverse sweep through every executed ctmrgstep (what Zygote does; init is constant)."""
    bulkbar = np.zeros_like(bulk)
    for rec in reversed(tape):
        db, cbar, ebar = ctmrgstep_backward(bulk, rec, cbar, ebar, eta)
        bulkbar += db
    return bulkbar, cbar, ebar


# ----------------------------------------------------------------------------
# iPEPS energy  (ipeps.jl:32-39, variationalipeps.jl:28-56)
# ----------------------------------------------------------------------------
_SYM_PERMS = [(0, 3, 2, 1, 4), (2, 1, 0, 3, 4), (1, 0, 3, 2, 4), (3, 2, 1, 0, 4)]  # 0-based ipeps.jl:34-37


def indexperm_symmetrize(x, tape=None):
    """ipeps.jl:32-39"""
    for p in _SYM_PERMS:
        x = x + np.transpose(x, p)
    nrm = np.linalg.norm(x)
    if tape is not None:
        tape["sym_x"] = x
        tape["sym_n"] = nrm
    return x / nrm


def indexperm_symmetrize_back(ybar, tape):
    xbar = _norm_back(ybar, tape["sym_x"], tape["sym_n"])
    for p in reversed(_SYM_PERMS):           # all four permutations are involutions
        xbar = xbar + np.transpose(xbar, p)
    return xbar


def double_layer(A):
    """variationalipeps.jl:30-34: ap (D,D,D,D,s,s) with merged index = ket + d*bra, a = tr_s ap."""
    d, s = A.shape[0], A.shape[4]
    D = d * d
    ap = np.einsum("abcdx,ijkly->aibjckdlxy", A, np.conj(A))
    ap = rsh(ap, (D, D, D, D, s, s))
    a = np.einsum("ijklaa->ijkl", ap)
    return ap, a


def double_layer_back(A, apbar, abar):
    d, s = A.shape[0], A.shape[4]
    apbar = apbar + np.einsum("ijkl,xy->ijklxy", abar, np.eye(s))
    ap10 = rsh(apbar, (d, d, d, d, d, d, d, d, s, s))       # a i b j c k d l x y
    Abar = np.einsum("aibjckdlxy,ijkly->abcdx", ap10, A) + np.einsum("aibjckdlxy,abcdx->ijkly", ap10, A)
    return Abar


def expectationvalue(h, ap, corner, edge, tape=None):
    """variationalipeps.jl:49-56 in pairwise order."""
    nap = np.linalg.norm(ap)
    apn = ap / nap
    CT1 = es("ica,ab->icb", edge, corner)          # C[a,b] T[i,c,a]
    CTr = es("eg,gfk->efk", corner, edge)          # C[e,g] T[g,f,k]
    X = es("icb,bde->icde", CT1, edge)
    Y = es("icde,cjfdlm->iejflm", X, apn)
    l = es("iejflm,efk->ijklm", Y, CTr)
    lh = es("abckl,ijkl->abcij", l, h)
    e = np.sum(l * lh)
    tl = np.einsum("ijkaa->ijk", l)
    nn = np.sum(tl * tl)
    if tape is not None:
        tape.update(ap=ap, nap=nap, apn=apn, CT1=CT1, CTr=CTr, X=X, Y=Y, l=l, e=e, nn=nn, tl=tl, h=h)
    return e / nn


def expectationvalue_back(corner, edge, tape, ybar=1.0):
    """SURVEY appendix B.2. Returns (dap, dcorner, dedge)."""
    h, l, e, nn, tl = tape["h"], tape["l"], tape["e"], tape["nn"], tape["tl"]
    s = h.shape[0]
    ebar = ybar / nn
    nbar = -ybar * e / nn ** 2
    hs = h + np.transpose(h, (2, 3, 0, 1))
    lbar = ebar * es("abckl,ijkl->abcij", l, hs) + 2 * nbar * np.einsum("abc,ij->abcij", tl, np.eye(s))
    Y, CTr, X, apn, CT1 = tape["Y"], tape["CTr"], tape["X"], tape["apn"], tape["CT1"]
    Ybar = es("ijklm,efk->iejflm", lbar, CTr)
    CTrbar = es("iejflm,ijklm->efk", Y, lbar)
    Xbar = es("iejflm,cjfdlm->icde", Ybar, apn)
    apnbar = es("icde,iejflm->cjfdlm", X, Ybar)
    CT1bar = es("icde,bde->icb", Xbar, edge)
    edgebar = es("icb,icde->bde", CT1, Xbar)
    cornerbar = es("efk,gfk->eg", CTrbar, edge)
    edgebar = edgebar + es("eg,efk->gfk", corner, CTrbar)
    edgebar = edgebar + es("icb,ab->ica", CT1bar, corner)
    cornerbar = cornerbar + es("ica,icb->ab", edge, CT1bar)
    apbar = _norm_back(apnbar, tape["ap"], tape["nap"])
    return apbar, cornerbar, edgebar


def energy(h, A, chi, tol, maxit, driver="gesdd"):
    """variationalipeps.jl:28-40 forward."""
    As = indexperm_symmetrize(np.asarray(A, dtype=float))
    ap, a = double_layer(As)
    corner, edge = init_raw(a, chi)
    corner, edge, _, _ = ctmrg(a, corner, edge, tol, maxit, driver)
    return expectationvalue(h, ap, corner, edge)


def energy_value_and_grad(h, A, chi, tol, maxit, driver="gesdd", eta=1e-40, info=None):
    """energy and dE/dA exactly as Zygote + src/autodiff.jl compute them (init constant, BPTT)."""
    A = np.asarray(A, dtype=float)
    t0 = {}
    As = indexperm_symmetrize(A, t0)
    ap, a = double_layer(As)
    c0, e0 = init_raw(a, chi)
    steps = []
    corner, edge, vals, nsteps = ctmrg(a, c0, e0, tol, maxit, driver, steps)
    t1 = {}
    y = expectationvalue(h, ap, corner, edge, t1)
    apbar, cbar, ebar = expectationvalue_back(corner, edge, t1, 1.0)
    abar, _, _ = ctmrg_backward(a, steps, cbar, ebar, eta)
    Asbar = double_layer_back(As, apbar, abar)
    Abar = indexperm_symmetrize_back(Asbar, t0)
    if info is not None:
        info["nsteps"] = nsteps
        info["vals"] = vals
    return y, Abar


# ----------------------------------------------------------------------------
# Ising read-out  (exampletensors.jl:57-69)
# ----------------------------------------------------------------------------
def magnetisation_readout(a, m, corner, edge):
    """exampletensors.jl:63-68"""
    ctc = es("ia,ajb,bk->ijk", corner, edge, corner)
    env = es("alc,ckd,bjd,bia->ijkl", ctc, edge, ctc, edge)
    mag = np.sum(env * m)
    nrm = np.sum(env * a)
    return abs(mag / nrm)


def magnetisation_readout_back(a, m, corner, edge, ybar=1.0):
    """Reverse of magnetisation_readout (exampletensors.jl:63-68 under Zygote): cotangents of
    (a, m, corner, edge) for y = |mag/norm|."""
    ct = es("ia,ajb->ijb", corner, edge)
    ctc = es("ijb,bk->ijk", ct, corner)
    e1 = es("alc,ckd->alkd", ctc, edge)
    e2 = es("bjd,bia->jdia", ctc, edge)
    env = es("alkd,jdia->ijkl", e1, e2)
    mag, nrm = np.sum(env * m), np.sum(env * a)
    r = mag / nrm
    sg = 1.0 if r >= 0 else -1.0
    magbar, nrmbar = ybar * sg / nrm, -ybar * sg * mag / nrm ** 2
    envbar = magbar * m + nrmbar * a
    abar, mbar = nrmbar * env, magbar * env
    e1bar = es("ijkl,jdia->alkd", envbar, e2)
    e2bar = es("alkd,ijkl->jdia", e1, envbar)
    ctcbar = es("alkd,ckd->alc", e1bar, edge) + es("jdia,bia->bjd", e2bar, edge)
    edgebar = es("alc,alkd->ckd", ctc, e1bar) + es("bjd,jdia->bia", ctc, e2bar)
    ctbar = es("ijk,bk->ijb", ctcbar, corner)
    cornerbar = es("ijb,ijk->bk", ct, ctcbar) + es("ijb,ajb->ia", ctbar, edge)
    edgebar = edgebar + es("ia,ijb->ajb", corner, ctbar)
    return abar, mbar, cornerbar, edgebar


def dmag_tensor_ising(beta):
    """d mag_tensor / d beta (exampletensors.jl:43-48 differentiated: product rule over the four q factors)."""
    cb, sb = math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))
    q = _ising_q(beta)
    dcb, dsb = math.sinh(beta) / (2 * cb), math.cosh(beta) / (2 * sb)
    dq = 1 / math.sqrt(2) * np.array([[dcb + dsb, dcb - dsb], [dcb - dsb, dcb + dsb]])
    a = np.zeros((2, 2, 2, 2))
    a[0, 0, 0, 0] = 1.0
    a[1, 1, 1, 1] = -1.0
    out = np.zeros((2, 2, 2, 2))
    for slot in range(4):
        qs = [q, q, q, q]
        qs[slot] = dq
        out += np.einsum("abcd,ai,bj,ck,dl->ijkl", a, *qs)
    return out


def magnetisation_value_and_dbeta(beta, chi, corner0, edge0, tol=1e-6, maxit=100):
    """magnetisation and d/d beta as Zygote computes it at test/ctmrg.jl:44-46: the environment initialisation is
    constant (autodiff.jl:5), the gradient flows through every ctmrgstep (bulk = a) and through the read-out."""
    a, m = model_tensor_ising(beta), mag_tensor_ising(beta)
    tape = []
    corner, edge, _, _ = ctmrg(a, corner0, edge0, tol, maxit, tape=tape)
    y = magnetisation_readout(a, m, corner, edge)
    abar, mbar, cbar, ebar = magnetisation_readout_back(a, m, corner, edge)
    bbar, _, _ = ctmrg_backward(a, tape, cbar, ebar)
    abar = abar + bbar
    return y, float(np.sum(abar * dmodel_tensor_ising(beta)) + np.sum(mbar * dmag_tensor_ising(beta)))


def magnetisation(beta, chi, rng=None, tol=1e-6, maxit=100, env="random"):
    """exampletensors.jl:57-69 (reference hard-codes :random, tol=1e-6, maxit=100)."""
    a = model_tensor_ising(beta)
    m = mag_tensor_ising(beta)
    if env == "random":
        corner, edge = init_random(a, chi, rng or np.random.default_rng(0))
    else:
        corner, edge = init_raw(a, chi)
    corner, edge, _, _ = ctmrg(a, corner, edge, tol, maxit)
    return magnetisation_readout(a, m, corner, edge)
