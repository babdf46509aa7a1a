"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle, the committed golden
vectors and the reference's own known-answer tests.  Tolerances are BASELINE.json's:
1e-10 relative on lnZ / energies, 1e-8 relative on gradients; contractions and SVD pieces are held to
1e-12 or better."""
import numpy as np
import pytest

import tnad_b200 as T
import tnad_oracle as O

pytestmark = pytest.mark.gpu

TOL_E, TOL_G = 1e-10, 1e-8


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


# ---- contractions (OMEinsum call sites) -------------------------------------------------------------------------
CONTRACT_CASES = [
    ("ab,bc->ac", (5, 7), (7, 3)), ("ab,bc->ac", (257, 130), (130, 191)), ("ab,cb->ac", (200, 96), (150, 96)),
    ("ba,bc->ac", (96, 200), (96, 150)), ("ba,cb->ac", (96, 200), (150, 96)), ("ab,bc->ac", (512, 384), (384, 640)),
    ("iba,ad->ibd", (6, 4, 6), (6, 6)), ("ibcl,jkcb->ijlk", (20, 4, 4, 20), (4, 4, 4, 4)),
    ("ibcl,jkcb->ijlk", (7, 3, 3, 7), (3, 3, 3, 3)), ("abi,aed->ibed", (20, 4, 20), (20, 4, 20)),
    ("ibed,bjce->ijcd", (20, 4, 4, 20), (4, 4, 4, 4)), ("ijcd,dck->ijk", (20, 4, 4, 20), (20, 4, 20)),
    ("icde,cjfdlm->iejflm", (10, 4, 4, 10), (4, 4, 4, 4, 2, 2)), ("iejflm,efk->ijklm", (10, 10, 4, 4, 2, 2), (10, 4, 10)),
    ("abcij,ij->abc", (10, 4, 10, 2, 2), (2, 2)), ("abc,ij->abcij", (10, 4, 10), (2, 2)),
    ("npu,por->nour", (9, 7, 5), (7, 9, 6)), ("nour,dlno->urdl", (9, 9, 5, 6), (5, 6, 9, 9)),
    ("mk,m->k", (37, 11), (37,)), ("mk,k->m", (37, 11), (11,)), ("ibcl,jkcb->ijlk", (64, 9, 9, 64), (9, 9, 9, 9)),
    ("a,a->", (1000,), (1000,)),
]


@pytest.mark.parametrize("spec,sa,sb", CONTRACT_CASES)
def test_contract(ctx, spec, sa, sb):
    rng = np.random.default_rng(abs(hash((spec, sa))) % 2 ** 32)
    A, B = rng.standard_normal(sa), rng.standard_normal(sb)
    ref = np.einsum(spec, A, B, optimize=True)
    assert rel(ctx.contract(spec, A, B), ref) < 1e-13
    C0 = rng.standard_normal(ref.shape)
    assert rel(ctx.contract(spec, A, B, alpha=-0.5, beta=2.0, Cin=C0), -0.5 * ref + 2.0 * C0) < 1e-13


def test_contract_linearity_at_full_size(ctx):
    # size-independent property at the headline shape: contraction is bilinear
    rng = np.random.default_rng(0)
    X2a, X2b = rng.standard_normal((128, 16, 16, 128)), rng.standard_normal((128, 16, 16, 128))
    bulk = rng.standard_normal((16, 16, 16, 16))
    spec = "ibcl,jkcb->ijlk"
    lhs = ctx.contract(spec, X2a + 2.0 * X2b, bulk)
    rhs = ctx.contract(spec, X2a, bulk) + 2.0 * ctx.contract(spec, X2b, bulk)
    assert rel(lhs, rhs) < 1e-13
    assert rel(lhs[:3], np.einsum(spec, (X2a + 2.0 * X2b)[:3], bulk, optimize=True)) < 1e-13


# Shapes the tensor-map kernel (gemm_tma.cu) accepts: k-fast / row-fast operands, two-level rows and K, batch, ragged
# edges (zero-filled boxes), split-K with the last-arriver reduction.  Each one must agree with numpy AND with the
# cp.async kernel (TNAD_GEMM_TMA=0) on the same operands.
TMA_CASES = [
    ("ab,bc->ac", (256, 128), (128, 192)), ("ab,cb->ac", (200, 96), (150, 96)), ("ba,bc->ac", (96, 200), (96, 150)),
    ("ba,cb->ac", (96, 200), (150, 96)), ("ab,bc->ac", (130, 2048), (2048, 70)), ("ab,bc->ac", (64, 4096), (4096, 64)),
    ("ibd,dcl->ibcl", (32, 16, 32), (32, 16, 32)), ("ibcl,jkcb->ijlk", (32, 16, 16, 32), (16, 16, 16, 16)),
    ("abi,aed->ibed", (32, 16, 32), (32, 16, 32)), ("ibed,bjce->ijcd", (32, 16, 16, 32), (16, 16, 16, 16)),
    ("ijcd,dck->ijk", (32, 16, 16, 32), (32, 16, 32)), ("pi,pj->ij", (2048, 48), (2048, 48)),
    ("rpi,rpj->ijp", (300, 3, 64), (300, 3, 64)), ("ab,bc->ac", (77, 33), (33, 18)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("spec,sa,sb", TMA_CASES)
def test_contract_tma_kernel_vs_cp_async_kernel(ctx, opt, spec, sa, sb):
    rng = np.random.default_rng(abs(hash((spec, sa, sb))) % 2 ** 32)
    A, B = rng.standard_normal(sa), rng.standard_normal(sb)
    ref = np.einsum(spec, A, B, optimize=True)
    c_tma = ctx.contract(spec, A, B)
    C0 = rng.standard_normal(ref.shape)
    c_tma_ab = ctx.contract(spec, A, B, alpha=-0.5, beta=2.0, Cin=C0)
    c_tma_again = ctx.contract(spec, A, B)
    opt("TNAD_GEMM_TMA", "0")
    c_old = ctx.contract(spec, A, B)
    assert rel(c_tma, ref) < 1e-13 and rel(c_old, ref) < 1e-13
    assert rel(c_tma_ab, -0.5 * ref + 2.0 * C0) < 1e-13
    assert np.array_equal(c_tma, c_tma_again)           # split-K sums in a fixed order: bit-for-bit repeatable


@pytest.mark.gpu
def test_contract_tma_forced_split_k(ctx, opt):
    rng = np.random.default_rng(3)
    A, B = rng.standard_normal((128, 1024)), rng.standard_normal((1024, 192))
    ref = A @ B
    for s in ("1", "2", "5", "16"):
        opt("TNAD_GEMM_SPLITK", s)
        assert rel(ctx.contract("ab,bc->ac", A, B), ref) < 1e-13, s


@pytest.mark.gpu
@pytest.mark.parametrize("n", [520, 1000, 1536])
def test_svd_sym_explicit_q_same_as_reflector_passes(ctx, opt, n):
    """Explicit-Q mode (Q1, Q2, Q1 Q2 on the side stream, one product at the end) against the reflector passes."""
    opt("TNAD_SYMEIG", "2")
    opt("TNAD_EIG_2STAGE", "1")
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)); a = a + a.T
    res = {}
    for mode, ident in (("0", "1"), ("2", "0"), ("2", "1")):
        opt("TNAD_EXPLICIT_Q", mode)
        opt("TNAD_Q2_IDENT", ident)
        res[(mode, ident)] = ctx.svd_sym(a)
    u0, s0, v0 = res[("0", "1")]
    for key in (("2", "0"), ("2", "1")):
        u, s, v = res[key]
        assert np.abs(s - s0).max() <= 1e-13 * s0[0]
        assert np.abs((u * s) @ v.T - a).max() <= 1e-12 * s0[0], key
        assert np.abs(u.T @ u - np.eye(n)).max() <= 1e-12, key
        # canonical gauge: well separated vectors agree entry by entry
        gap_ok = np.abs(np.diff(s0)) > 1e-6 * s0[0]
        ok = np.concatenate([[True], gap_ok]) & np.concatenate([gap_ok, [True]])
        assert np.abs(u[:, ok] - u0[:, ok]).max() <= 1e-8, key


# ---- SVD (LinearAlgebra.svd call sites) -----------------------------------------------------------------------
def _check_svd(ctx, A, tol_rec=5e-14):
    U, S, V = ctx.svd(A)
    k = min(A.shape)
    Sref = np.linalg.svd(A, compute_uv=False)
    assert rel((U * S) @ V.T, A) < tol_rec
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-12 and np.abs(V.T @ V - np.eye(k)).max() < 1e-12
    assert np.abs(S - Sref).max() / Sref[0] < 1e-13 and np.all(np.diff(S) <= 0)


@pytest.mark.parametrize("shape", [(1, 1), (4, 4), (9, 9), (30, 30), (64, 64), (65, 65), (80, 80), (100, 60), (60, 100),
                                   (128, 128), (200, 200), (400, 400), (7, 300), (300, 7)])
def test_svd_random(ctx, shape):
    _check_svd(ctx, np.random.default_rng(sum(shape)).standard_normal(shape))


def test_svd_rank_deficient_and_structured(ctx):
    rng = np.random.default_rng(3)
    _check_svd(ctx, rng.standard_normal((50, 7)) @ rng.standard_normal((7, 50)))
    a = O.model_tensor_ising(0.5)
    _check_svd(ctx, np.reshape(np.transpose(a, (2, 1, 0, 3)), (4, 4), order="F"))
    Q, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    lam = np.concatenate([np.linspace(1, 2, 40), -np.linspace(1, 2, 40), np.zeros(16)])
    _check_svd(ctx, (Q * lam) @ Q.T)
    _check_svd(ctx, np.diag(np.arange(1.0, 11.0)))
    _check_svd(ctx, np.ones((20, 20)))


def test_svd_large_symmetric(ctx):
    A = np.random.default_rng(0).standard_normal((1024, 1024))
    _check_svd(ctx, A + A.T, tol_rec=2e-13)


def _check_svd_sym(ctx, A, tol_rec=5e-13):
    U, S, V = ctx.svd_sym(A)
    n = A.shape[0]
    Sref = np.linalg.svd(A, compute_uv=False)
    scale = max(Sref[0], 1e-300)
    assert np.linalg.norm((U * S) @ V.T - A) <= tol_rec * max(np.linalg.norm(A), 1e-300)
    assert np.abs(U.T @ U - np.eye(n)).max() < 1e-12 and np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    assert np.abs(S - Sref).max() / scale < 1e-12 and np.all(np.diff(S) <= 0)
    # V = U sign(lambda): columns of U are eigenvectors
    lam = np.einsum("ij,ij->j", U, A @ U)
    assert np.abs(np.abs(lam) - S).max() / scale < 1e-12


@pytest.mark.parametrize("n", [1, 2, 7, 31, 64, 65, 100, 200, 512])
def test_svd_sym_random(ctx, n):
    A = np.random.default_rng(n).standard_normal((n, n))
    _check_svd_sym(ctx, A + A.T)


def test_svd_sym_structured(ctx):
    rng = np.random.default_rng(8)
    Q, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    lam = np.concatenate([np.linspace(1, 2, 40), -np.linspace(1.01, 2.01, 40), np.zeros(16)])   # indefinite + null space
    _check_svd_sym(ctx, (Q * lam) @ Q.T)
    lam = np.concatenate([np.logspace(0, -14, 80), np.zeros(16)]) * rng.choice([-1.0, 1.0], 96)  # graded spectrum
    _check_svd_sym(ctx, (Q * lam) @ Q.T)
    _check_svd_sym(ctx, np.diag(np.arange(1.0, 41.0)))
    _check_svd_sym(ctx, np.ones((40, 40)))
    B = rng.standard_normal((300, 12))
    _check_svd_sym(ctx, B @ B.T)                                                                  # rank 12 of 300
    U, S, V = ctx.svd_sym(np.zeros((10, 10)))
    assert np.all(S == 0) and np.abs(U.T @ U - np.eye(10)).max() < 1e-14


def test_svd_sym_matches_general_svd_on_ctmrg_matrix(ctx):
    # the matrix ctmrgstep decomposes: cp + cp' for a random environment (ctmrg.jl:134-136)
    rng = np.random.default_rng(12)
    D, chi = 3, 14
    bulk = rng.standard_normal((D, D, D, D)); bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c, e = O.init_random(bulk, chi, rng)
    X2 = np.einsum("ibd,dcl->ibcl", np.einsum("iba,ad->ibd", e, c), e)
    cp = np.einsum("ibcl,jkcb->ijlk", X2, bulk).reshape(chi * D, chi * D, order="F")
    M = cp + cp.T
    U, S, V = ctx.svd_sym(M)
    U2, S2, V2 = ctx.svd(M)
    assert np.abs(S - S2).max() / S[0] < 1e-12
    # leading subspace (what the projector z uses) agrees between the two solvers
    P1, P2 = U[:, :chi] @ U[:, :chi].T, U2[:, :chi] @ U2[:, :chi].T
    if S[chi - 1] - S[chi] > 1e-8 * S[0]:
        assert np.abs(P1 - P2).max() < 1e-9


def test_trg_svd_unit(ctx):
    # test/trg.jl:6-10
    rng = np.random.default_rng(0)
    t = rng.standard_normal((10, 10, 10, 10))
    u, v = ctx.trg_svd(t, 100, 0.0)
    assert u.shape == (10, 10, 100) and v.shape == (100, 10, 10)
    assert rel(np.einsum("ija,akl->ijkl", u, v), t) < 1e-12
    t = rng.standard_normal((6, 5, 4, 7))
    u2, v2 = ctx.trg_svd(t, 8, 1e-16)
    uo, vo, _ = O.trg_svd(t, 8, 1e-16)
    assert u2.shape == uo.shape and rel(np.einsum("ija,akl->ijkl", u2, v2), np.einsum("ija,akl->ijkl", uo, vo)) < 1e-12


@pytest.mark.parametrize("m,n", [(6, 3), (3, 6), (3, 3), (40, 40)])
def test_svd_back_matches_reference_formula(ctx, m, n):
    # test/svd.jl shapes; every combination of present / `nothing` cotangents (trg.jl:72-105)
    rng = np.random.default_rng(m * 100 + n)
    A = rng.standard_normal((m, n))
    U, S, V = O.svd(A)
    k = min(m, n)
    dU, dS, dV = rng.standard_normal((m, k)), rng.standard_normal(k), rng.standard_normal((n, k))
    for mask in [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1)]:
        a = [x if f else None for x, f in zip((dU, dS, dV), mask)]
        assert rel(ctx.svd_back(U, S, V, *a), O.svd_back(U, S, V, *a)) < 1e-12
    assert ctx.svd_back(U, S, V, None, None, None) is None


def test_svd_back_golden(ctx, golden):
    _, vec = golden
    got = ctx.svd_back(vec["sb_U"], vec["sb_S"], vec["sb_V"], vec["sb_dU"], vec["sb_dS"], vec["sb_dV"])
    assert rel(got, vec["sb_out"]) < 1e-12


def test_svd_gradient_check_through_gpu_svd(ctx):
    # test/svd.jl:6-13 gradient_check, real case: loss of U[:,0], V[:,0] and S through the GPU svd + svd_back
    rng = np.random.default_rng(5)
    for (m, n) in [(6, 3), (3, 6), (3, 3)]:
        A = rng.standard_normal((m, n))
        H1 = rng.standard_normal((m, m)); H1 = H1 + H1.T
        H2 = rng.standard_normal((n, n)); H2 = H2 + H2.T

        def loss(A):
            U, S, V = ctx.svd(A)
            return U[:, 0] @ H1 @ U[:, 0] + V[:, 0] @ H2 @ V[:, 0] + S.sum()
        U, S, V = ctx.svd(A)
        dU = np.zeros_like(U); dU[:, 0] = 2 * H1 @ U[:, 0]
        dV = np.zeros_like(V); dV[:, 0] = 2 * H2 @ V[:, 0]
        g = ctx.svd_back(U, S, V, dU, np.ones_like(S), dV)
        eta = 1e-5
        dy = loss(A) - loss(A - eta * g)
        assert dy == pytest.approx(eta * np.sum(g * g), rel=1e-2, abs=1e-8)


# ---- TRG ---------------------------------------------------------------------------------------------------------
def test_trg_published_goldens(ctx, golden):
    pub, _ = golden
    a = T.model_tensor(T.Ising(), 0.4)
    assert T.trg(a, 5, 5, ctx=ctx) == pytest.approx(pub["trg_beta0.4_chi5_n5"]["value"], rel=TOL_E)
    lnz, g = T.trg_value_and_grad(T.model_tensor(T.Ising(), 0.5), 20, 20, ctx=ctx)
    assert lnz == pytest.approx(pub["trg_beta0.5_chi20_n20"]["value"], rel=TOL_E)
    assert float(np.sum(g * T.dmodel_tensor(T.Ising(), 0.5))) == pytest.approx(pub["dtrg_beta0.5_chi20_n20"]["value"], rel=TOL_G)
    _, g = T.trg_value_and_grad(T.model_tensor(T.Ising(), 0.5), 5, 5, ctx=ctx)
    assert float(np.sum(g * T.dmodel_tensor(T.Ising(), 0.5))) == pytest.approx(pub["dtrg_beta0.5_chi5_n5"]["value"], rel=TOL_G)


@pytest.mark.parametrize("beta,chi,n", [(0.4, 5, 5), (0.5, 5, 5), (0.44, 8, 10), (0.5, 20, 20)])
def test_trg_vs_golden_vectors(ctx, golden, beta, chi, n):
    _, vec = golden
    key = f"trg_{beta}_{chi}_{n}"
    lnz, g = T.trg_value_and_grad(T.model_tensor(T.Ising(), beta), chi, n, ctx=ctx)
    assert lnz == pytest.approx(float(vec[key + "_lnz"]), rel=TOL_E)
    assert float(np.sum(g * T.dmodel_tensor(T.Ising(), beta))) == pytest.approx(float(vec[key + "_dbeta"]), rel=TOL_G)
    # The Ising tensor is exactly rank-deficient in the first iterations: sqrt(s) is not differentiable at s = 0,
    # so only derivatives along structure-preserving directions (d/d beta) are defined; the components of the
    # tensor gradient that lift the rank depend on LAPACK's arbitrary null vectors in the reference itself.
    # Full tensor gradients are compared on a generic full-rank tensor in test_trg_generic_tensor_against_oracle.


def test_trg_gradient_vs_numgrad(ctx):
    # test/trg.jl:19
    f = lambda b: T.trg(T.model_tensor(T.Ising(), b), 5, 5, ctx=ctx)  # noqa: E731
    _, g = T.trg_value_and_grad(T.model_tensor(T.Ising(), 0.4), 5, 5, ctx=ctx)
    assert T.num_grad(f, 0.4, 1e-6) == pytest.approx(float(np.sum(g * T.dmodel_tensor(T.Ising(), 0.4))), rel=2e-8)


def test_trg_generic_tensor_against_oracle(ctx):
    rng = np.random.default_rng(11)
    a = np.abs(rng.standard_normal((3, 2, 3, 2))) + 0.1
    ref, gref = O.trg_value_and_grad(a, 6, 4)
    lnz, g = T.trg_value_and_grad(a, 6, 4, ctx=ctx)
    assert lnz == pytest.approx(ref, rel=TOL_E) and rel(g, gref) < TOL_G


def test_trg_errors(ctx):
    with pytest.raises(T.DimensionMismatch):
        T.trg(np.zeros((2, 3, 3, 2)), 4, 2, ctx=ctx)
    with pytest.raises(T.TnadError):
        T.trg(np.zeros((2, 2, 2, 2)), 4, 2, ctx=ctx)     # vanished tensor: log(0) in the reference


def test_trg_zero_iterations(ctx):
    a = T.model_tensor(T.Ising(), 0.3)
    assert T.trg(a, 4, 0, ctx=ctx) == pytest.approx(O.trg(a, 4, 0), rel=1e-14)


# ---- CTMRG -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,chi", [(2, 5), (3, 10), (4, 20), (2, 1), (5, 3)])
def test_ctmrgstep_vs_oracle(ctx, D, chi):
    rng = np.random.default_rng(D * 10 + chi)
    bulk = rng.standard_normal((D, D, D, D))
    bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c, e = O.init_random(bulk, chi, rng)
    cr, er, vr = O.ctmrgstep(bulk, c, e, signfix=True)
    cg, eg, vg = ctx.ctmrgstep(bulk, c, e)
    assert np.abs(vg - vr).max() < 1e-12
    # corner / edge are defined up to column signs of U: both sides use the canonical gauge
    # (largest-magnitude entry of every U column positive), so the tensors themselves are compared
    assert np.abs(cg - cr).max() < 1e-10 and np.abs(eg - er).max() < 1e-10
    assert np.linalg.norm(cg) == pytest.approx(1.0, rel=1e-13) and np.linalg.norm(eg) == pytest.approx(1.0, rel=1e-13)
    assert np.allclose(cg, cg.T, atol=1e-15) and np.allclose(eg, np.transpose(eg, (2, 1, 0)), atol=1e-15)


def test_init_raw(ctx):
    for D, chi in [(2, 4), (4, 3), (3, 3)]:
        bulk = np.random.default_rng(D).standard_normal((D, D, D, D))
        c, e = ctx.ctmrg_init_raw(bulk, chi)
        co, eo = O.init_raw(bulk, chi)
        assert np.abs(c - co).max() < 1e-14 and np.abs(e - eo).max() < 1e-14


def test_ctmrg_stop_rule(ctx):
    a = T.model_tensor(T.Ising(), 0.3)
    rt = T.SquareCTMRGRuntime(a, "raw", 4, ctx=ctx)
    for maxit in (0, 1, 5):
        assert T.ctmrg(rt, 0.0, maxit, ctx=ctx).steps == maxit + 1        # counter starts at -1 (ctmrg.jl:114)
    out = T.ctmrg(rt, 1e-10, 500, ctx=ctx)
    c0, e0 = O.init_raw(a, 4)
    assert out.steps == O.ctmrg(a, c0, e0, 1e-10, 500)[3]


@pytest.mark.parametrize("beta,chi", [(0.3, 16), (0.5, 16)])
def test_ctmrg_raw_spectrum_golden(ctx, golden, beta, chi):
    _, vec = golden
    a = T.model_tensor(T.Ising(), beta)
    c0, e0 = ctx.ctmrg_init_raw(a, chi)
    _, _, vals, steps = ctx.ctmrg(a, c0, e0, 1e-10, 500)
    assert steps == int(vec[f"ctmrg_raw_{beta}_{chi}_steps"])
    assert np.abs(vals - vec[f"ctmrg_raw_{beta}_{chi}_vals"]).max() < 1e-9


@pytest.mark.parametrize("beta,chi,atol", [(1.0, 2, 1e-8), (0.6, 4, 1e-8), (0.8, 2, 1e-8), (0.2, 10, 1e-4), (0.4, 10, 2e-3)])
def test_onsager_magnetisation(ctx, beta, chi, atol):
    # test/ctmrg.jl:37-42 (:random environment, tol 1e-6, maxit 100)
    m = T.magnetisation(T.Ising(), beta, chi, rng=np.random.default_rng(5), ctx=ctx)
    assert abs(m - T.magofbeta(T.Ising(), beta)) < atol


def test_magnetisation_gradient(ctx):
    # test/ctmrg.jl:44-46: Zygote.gradient(beta -> magnetisation(Ising(), beta, 2), 0.5) against num_grad (atol 1e-2)
    # and, tighter, against the oracle's reverse sweep on the same seeded :random environment
    f = lambda b: T.magnetisation(T.Ising(), b, 2, rng=np.random.default_rng(9), ctx=ctx)  # noqa: E731
    y, g = T.magnetisation_value_and_grad(T.Ising(), 0.5, 2, rng=np.random.default_rng(9), ctx=ctx)
    assert y == pytest.approx(f(0.5), rel=1e-12)
    assert abs(g - T.num_grad(f, 0.5, 1e-3)) < 1e-2
    for beta, chi in [(0.5, 2), (0.6, 4), (0.3, 5)]:
        a = O.model_tensor_ising(beta)
        c0, e0 = O.init_random(a, chi, np.random.default_rng(9))
        yo, go = O.magnetisation_value_and_dbeta(beta, chi, c0, e0, tol=1e-6, maxit=100)
        y, g = T.magnetisation_value_and_grad(T.Ising(), beta, chi, rng=np.random.default_rng(9), ctx=ctx)
        assert y == pytest.approx(yo, rel=1e-10, abs=1e-12) and g == pytest.approx(go, rel=1e-7, abs=1e-10)


def test_init_random_on_device(ctx):
    # ctmrg.jl:66-72: randn + symmetrisation; reproducible from the seed, different seeds differ, moments of N(0, 2) / N(0, 4)
    c, e = ctx.ctmrg_init_random(4, 96, 7)
    c2, e2 = ctx.ctmrg_init_random(4, 96, 7)
    c3, _ = ctx.ctmrg_init_random(4, 96, 8)
    assert np.array_equal(c, c2) and np.array_equal(e, e2) and not np.array_equal(c, c3)
    assert np.array_equal(c, c.T) and np.array_equal(e, np.transpose(e, (2, 1, 0)))
    off = e[np.triu_indices(96, 1)[0], :, np.triu_indices(96, 1)[1]]        # x_ij + x_ji, i != j: variance 2
    assert abs(off.mean()) < 0.02 and abs(off.var() - 2.0) < 0.05
    assert abs(np.diag(c).var() - 4.0) < 1.5 and np.isfinite(e).all()
    # and it is a usable start: CTMRG Ising converges to the Onsager magnetisation from it
    a, m = T.model_tensor(T.Ising(), 0.5), T.mag_tensor(T.Ising(), 0.5)
    c0, e0 = ctx.ctmrg_init_random(2, 16, 3)
    cc, ee, vals, steps = ctx.ctmrg(a, c0, e0, 1e-10, 500)
    assert abs(abs(ctx.magnetisation_readout(a, m, cc, ee)) - T.magofbeta(T.Ising(), 0.5)) < 1e-6


def test_magnetisation_readout_backward_vs_oracle(ctx):
    rng = np.random.default_rng(4)
    a, m = O.model_tensor_ising(0.45), O.mag_tensor_ising(0.45)
    c, e = O.init_random(a, 6, rng)
    got = ctx.magnetisation_backward(a, m, c, e, 0.7)
    ref = O.magnetisation_readout_back(a, m, c, e, 0.7)
    for g, r in zip(got, ref):
        assert rel(g, r) < 1e-12


def test_ctmrgstep_backward_vs_oracle(ctx):
    # pullback of a single step in the canonical gauge
    rng = np.random.default_rng(31)
    D, chi = 3, 6
    bulk = rng.standard_normal((D, D, D, D)); bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c, e = O.init_random(bulk, chi, rng)
    tape = []
    O.ctmrgstep(bulk, c, e, tape=tape, signfix=True)
    cb, eb = rng.standard_normal((chi, chi)), rng.standard_normal((chi, D, chi))
    bo, co, eo = O.ctmrgstep_backward(bulk, tape[0], cb, eb)
    bg, cg, eg = ctx.ctmrgstep_backward(bulk, c, e, cb, eb)
    assert rel(bg, bo) < TOL_G and rel(cg, co) < TOL_G and rel(eg, eo) < TOL_G


def test_expectationvalue_backward_vs_oracle(ctx):
    rng = np.random.default_rng(6)
    h = T.hamiltonian(T.Heisenberg())
    A = O.indexperm_symmetrize(rng.standard_normal((2, 2, 2, 2, 2)))
    ap, a = O.double_layer(A)
    c, e = O.init_random(a, 5, rng)
    t = {}
    O.expectationvalue(h, ap, c, e, t)
    ref = O.expectationvalue_back(c, e, t, 1.3)
    got = ctx.expectationvalue_backward(h, ap, c, e, 1.3)
    for g, r in zip(got, ref):
        assert rel(g, r) < 1e-11


def test_ctmrg_backward_vs_oracle(ctx):
    rng = np.random.default_rng(21)
    D, chi = 3, 7
    bulk = rng.standard_normal((D, D, D, D)); bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c0, e0 = O.init_random(bulk, chi, rng)
    tape = []
    co, eo, _, ns = O.ctmrg(bulk, c0, e0, 0.0, 3, tape=tape)
    cg, eg, _, ng, gtape = ctx.ctmrg(bulk, c0, e0, 0.0, 3, want_tape=True)
    assert ns == ng == 4
    # gauge-invariant loss: L = sum(corner^2 * Wc) is NOT invariant; use |.|-free invariants instead
    Wc = rng.standard_normal((chi, chi)); We = rng.standard_normal((chi, D, chi))
    # cotangents must be expressed in each implementation's own gauge: use L = <c, c>_Wsym with sign-insensitive weights
    cbar_o, ebar_o = 2 * co * np.abs(Wc), 2 * eo * np.abs(We)
    cbar_g, ebar_g = 2 * cg * np.abs(Wc), 2 * eg * np.abs(We)
    bo, c0o, e0o = O.ctmrg_backward(bulk, tape, cbar_o, ebar_o)
    bg, c0g, e0g = ctx.ctmrg_backward(gtape, cbar_g, ebar_g, want_init=True)
    gtape.free()
    assert rel(bg, bo) < TOL_G and rel(c0g, c0o) < TOL_G and rel(e0g, e0o) < TOL_G


# ---- energy ----------------------------------------------------------------------------------------------------------
def test_expectationvalue_vs_oracle(ctx):
    rng = np.random.default_rng(2)
    h = T.hamiltonian(T.Heisenberg())
    A = O.indexperm_symmetrize(rng.standard_normal((2, 2, 2, 2, 2)))
    ap, a = O.double_layer(A)
    c, e = O.init_random(a, 6, rng)
    assert ctx.expectationvalue(h, ap, c, e) == pytest.approx(O.expectationvalue(h, ap, c, e), rel=1e-12)


@pytest.mark.parametrize("name", ["e_d2_chi4", "e_d3_chi12", "e_d2_chi16"])
def test_energy_and_gradient_vs_golden(ctx, golden, name):
    _, vec = golden
    d, chi, maxit = [int(x) for x in vec[name + "_cfg"]]
    h = T.hamiltonian(T.Heisenberg())
    e, g = T.energy_and_gradient(h, vec[name + "_A"], chi, 0.0, maxit, ctx=ctx)
    assert ctx.last_steps == int(vec[name + "_steps"]) == maxit + 1
    assert e == pytest.approx(float(vec[name + "_e"]), rel=TOL_E)
    assert rel(g, vec[name + "_grad"]) < TOL_G
    assert T.energy(h, vec[name + "_A"], chi, 0.0, maxit, ctx=ctx) == pytest.approx(e, rel=1e-13)


def test_energy_c3_readme_config(ctx, golden):
    # BASELINE configs[2]: Heisenberg d=2, chi=20, tol=1e-6, maxit=100
    _, vec = golden
    h = T.hamiltonian(T.Heisenberg())
    e, g = T.energy_and_gradient(h, T.SquareIPEPS(vec["c3_A"]), 20, 1e-6, 100, ctx=ctx)
    assert ctx.last_steps == int(vec["c3_steps"])
    assert e == pytest.approx(float(vec["c3_e"]), rel=TOL_E) and rel(g, vec["c3_grad"]) < TOL_G


def test_energy_gradient_vs_numgrad(ctx):
    # test/variationalipeps.jl:121-134
    h = T.hamiltonian(T.Heisenberg())
    A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))).bulk
    _, g = T.energy_and_gradient(h, A, 4, 0.0, 100, ctx=ctx)
    gn = T.num_grad(lambda x: T.energy(h, x, 4, 0.0, 100, ctx=ctx), A, 1e-3)
    assert np.allclose(g, gn, atol=1e-3) and np.abs(g - gn).max() < 1e-5


def test_noninteracting_energies(ctx):
    # test/variationalipeps.jl:9-25
    rng = np.random.default_rng(0)
    h = T.diaglocalhamiltonian([1, -1.0])
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 1] = rng.standard_normal()
    assert T.energy(h, T.SquareIPEPS(a), 4, 1e-12, 100, ctx=ctx) / 2 == pytest.approx(-1.0, abs=1e-9)
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 0] = rng.standard_normal()
    assert T.energy(h, T.SquareIPEPS(a), 10, 0, 300, ctx=ctx) / 2 == pytest.approx(1.0, abs=1e-9)
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 1] = a[0, 0, 0, 0, 0] = rng.standard_normal()
    assert abs(T.energy(h, T.SquareIPEPS(a), 10, 0, 300, ctx=ctx)) < 1e-9
    for _ in range(5):
        assert -1 < T.energy(h, T.SquareIPEPS(rng.random((3, 3, 3, 3, 2))), 5, 0, 10, ctx=ctx) / 2 < 1


def test_energy_three_level_physical_dim(ctx):
    # s = 3 (test/variationalipeps.jl:33-39 uses a 3-level diagonal Hamiltonian)
    h = T.diaglocalhamiltonian([0.3, 0.1, -0.43])
    A = np.random.default_rng(4).standard_normal((2, 2, 2, 2, 3))
    eo, go = O.energy_value_and_grad(h, A, 4, 0.0, 8)
    e, g = T.energy_and_gradient(h, A, 4, 0.0, 8, ctx=ctx)
    assert e == pytest.approx(eo, rel=TOL_E) and rel(g, go) < TOL_G


def test_energy_errors(ctx):
    h = T.hamiltonian(T.Heisenberg())
    with pytest.raises(T.DimensionMismatch):
        T.energy(h, np.zeros((3, 3, 4, 3, 2)), 4, 0.0, 1, ctx=ctx)
    with pytest.raises(T.DimensionMismatch):
        T.energy(np.zeros((3, 3, 3, 3)), np.ones((2, 2, 2, 2, 2)), 4, 0.0, 1, ctx=ctx)


def test_optimiseipeps_heisenberg(ctx, golden):
    # test/variationalipeps.jl:93-102: Heisenberg d=2, chi=4 -> -0.66023 (atol 1e-3)
    pub, _ = golden
    h = T.hamiltonian(T.Heisenberg())
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(2).standard_normal((2, 2, 2, 2, 2))))
    res = T.optimiseipeps(ipeps, h, chi=4, tol=0.0, maxit=50, optimargs={"f_tol": 1e-8, "iterations": 120}, ctx=ctx)
    # seed 2 reaches the global minimum (README.md:152: -0.6602311); seed 1 stalls in a local one with the oracle too
    assert abs(res.minimum - pub["heisenberg_energy_d2"]["value"]) < 1e-3


# ---- parity at the sizes bench.py measures (BASELINE configs 2, 4, 5 and the TRG chi=64 part of the metric) ----------
def _c4_A():
    return O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))


def test_c4_headline_energy_and_gradient_vs_oracle_fixture(ctx, large):
    """BASELINE configs[3], d=4 chi=128 (n = chi D = 2048: the eigensolver's large path, 128x128 GEMM tiles, split-K
    projector products): energy + gradient after maxit=3 (4 ctmrgsteps) against the frozen oracle run."""
    h = T.hamiltonian(T.Heisenberg())
    d, chi, maxit = [int(x) for x in large["c4_cfg"]]
    e, g = T.energy_and_gradient(h, _c4_A(), chi, 0.0, maxit, ctx=ctx)
    assert ctx.last_steps == int(large["c4_steps"]) == maxit + 1
    assert e == pytest.approx(float(large["c4_e"]), rel=TOL_E)
    assert rel(g, large["c4_grad"]) < TOL_G


def test_c4_headline_vs_oracle_live(ctx):
    """The same configuration at maxit=1 against the oracle run on this machine's LAPACK (about 4 s of CPU)."""
    h = T.hamiltonian(T.Heisenberg())
    A = _c4_A()
    eo, go = O.energy_value_and_grad(h, A, 128, 0.0, 1)
    e, g = T.energy_and_gradient(h, A, 128, 0.0, 1, ctx=ctx)
    assert e == pytest.approx(eo, rel=TOL_E) and rel(g, go) < TOL_G


@pytest.mark.parametrize("beta", [0.3, 0.5])
def test_c2_ising_chi64_raw(ctx, large, beta):
    """BASELINE configs[1]: CTMRG Ising chi=64, tol=1e-10, :raw environment: step count and converged spectrum."""
    a = T.model_tensor(T.Ising(), beta)
    c0, e0 = ctx.ctmrg_init_raw(a, 64)
    co, ed, vals, steps = ctx.ctmrg(a, c0, e0, 1e-10, 5000)
    assert steps == int(large[f"c2_raw_{beta}_steps"])
    assert np.abs(vals - large[f"c2_raw_{beta}_vals"]).max() < 1e-9
    # the :raw environment keeps the Z2 symmetry: no spontaneous magnetisation on either side
    assert ctx.magnetisation_readout(a, T.mag_tensor(T.Ising(), beta), co, ed) < 1e-6


@pytest.mark.parametrize("beta", [0.3, 0.5])
def test_c2_ising_chi64_random_env(ctx, large, beta):
    """configs[1] with the :random environment of a host-seeded generator (ctmrg.jl:66-72): step count, spectrum and
    magnetisation (Onsager above beta_c)."""
    a, m = T.model_tensor(T.Ising(), beta), T.mag_tensor(T.Ising(), beta)
    c0, e0 = O.init_random(a, 64, np.random.default_rng(3))
    co, ed, vals, steps = ctx.ctmrg(a, c0, e0, 1e-10, 5000)
    assert abs(steps - int(large[f"c2_random_{beta}_steps"])) <= 2      # tol sits on the rounding floor of ||dvals||
    assert np.abs(vals - large[f"c2_random_{beta}_vals"]).max() < 1e-8
    mag = ctx.magnetisation_readout(a, m, co, ed)
    assert mag == pytest.approx(float(large[f"c2_random_{beta}_mag"]), abs=1e-8)
    assert mag == pytest.approx(T.magofbeta(T.Ising(), beta), abs=1e-6)


def test_trg_chi64_vs_oracle_fixture(ctx, large):
    """TRG chi=64 (third part of BASELINE's metric): 7 iterations, the last two split full 4096 x 4096 matrices."""
    beta, chi, niter = float(large["trg64_cfg"][0]), int(large["trg64_cfg"][1]), int(large["trg64_cfg"][2])
    lnz, g = T.trg_value_and_grad(T.model_tensor(T.Ising(), beta), chi, niter, ctx=ctx)
    assert lnz == pytest.approx(float(large["trg64_lnz"]), rel=TOL_E)
    assert float(np.sum(g * T.dmodel_tensor(T.Ising(), beta))) == pytest.approx(float(large["trg64_dbeta"]), rel=TOL_G)


def test_c5a_step_vs_oracle_fixture(ctx, large):
    """BASELINE configs[4]: one ctmrgstep at d=5, chi=256 (n = 6400) from a seeded :random environment: spectrum and
    probes of the sign-fixed corner / edge."""
    A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((5, 5, 5, 5, 2)))
    _, a = O.double_layer(A)
    c0, e0 = O.init_random(a, 256, np.random.default_rng(1))
    co, ed, vals = ctx.ctmrgstep(a, c0, e0)
    rng = np.random.default_rng(7)
    w1, w2, w3 = rng.standard_normal(256), rng.standard_normal(25), rng.standard_normal(256)
    assert np.abs(vals - large["c5a_vals"]).max() < 1e-10
    assert np.abs(np.diag(co) - large["c5a_corner_diag"]).max() < 1e-9
    assert np.abs(co @ w1 - large["c5a_corner_w"]).max() < 1e-8
    assert np.abs(np.einsum("ijk,i,k->j", ed, w1, w3) - large["c5a_edge_w"]).max() < 1e-8
    assert np.abs(np.einsum("ijk,j,k->i", ed, w2, w3) - large["c5a_edge_w2"]).max() < 1e-8


def test_trg_sweep_matches_single_instances(ctx):
    """tnad_trg_sweep (in-library fan-out, one context per device) against instance-by-instance calls."""
    betas = [0.3, 0.4, 0.5]
    ts = [T.model_tensor(T.Ising(), b) for b in betas]
    lnz, grads = T.trg_sweep(ts, 8, 6, ngpu=1, grad=True)
    for i, b in enumerate(betas):
        l1, g1 = T.trg_value_and_grad(ts[i], 8, 6, ctx=ctx)
        assert lnz[i] == pytest.approx(l1, rel=1e-13)
        assert float(np.sum(grads[i] * T.dmodel_tensor(T.Ising(), b))) == pytest.approx(
            float(np.sum(g1 * T.dmodel_tensor(T.Ising(), b))), rel=1e-10)


def test_tape_outliving_close_is_safe():
    """a tape must not outlive its context's stream: Context.close() frees outstanding tapes first, and
    tnad_destroy refuses while tapes are alive."""
    c2 = T.Context(0)
    lnz, tape = c2.trg_forward(T.model_tensor(T.Ising(), 0.4), 4, 3, want_tape=True)
    assert c2.lib.tnad_destroy(c2.h) == 1          # TNAD_ERR_ARG: a live tape
    c2.close()                                     # frees the tape, then destroys
    assert tape.h is None


def test_headline_shape_runs_and_is_consistent(ctx):
    # d=4, chi=128 (BASELINE configs[3]) at maxit=1: size-independent checks -- the gradient of the
    # (scale-invariant) energy is orthogonal to A, energy reproducible call to call, launch counter moves.
    h = T.hamiltonian(T.Heisenberg())
    A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((4, 4, 4, 4, 2)))).bulk
    ctx.reset_launch_count()
    e1, g1 = T.energy_and_gradient(h, A, 128, 0.0, 1, ctx=ctx)
    assert ctx.launch_count() > 100 and ctx.last_steps == 2
    e2 = T.energy(h, A, 128, 0.0, 1, ctx=ctx)
    assert e2 == pytest.approx(e1, rel=1e-11)
    assert abs(np.sum(g1 * A)) < 1e-9 * np.linalg.norm(g1)
    for p in [(0, 3, 2, 1, 4), (2, 1, 0, 3, 4), (1, 0, 3, 2, 4), (3, 2, 1, 0, 4)]:
        assert rel(np.transpose(g1, p), g1) < 1e-9        # gradient inherits the index-permutation symmetry


# ---- chi-sharded step (tensornetworkad.jl_b200/sharded.py): world 1 here, world 2.. via bench_sharded.py -------------------
@pytest.mark.gpu
@pytest.mark.parametrize("D,chi", [(4, 8), (9, 12), (4, 32)])
def test_sharded_step_single_rank_vs_oracle(ctx, D, chi):
    from tnad_b200.sharded import ShardedCTMRG
    rng = np.random.default_rng(D * 100 + chi)
    bulk = rng.standard_normal((D, D, D, D))
    bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c, e = O.init_random(bulk, chi, rng)
    cr, er, vr = O.ctmrgstep(bulk, c, e, signfix=True)
    sh = ShardedCTMRG(ctx, chi, D)
    sh.load(bulk, c, e)
    sh.step()
    cg, eg, vg = sh.result()
    assert np.abs(vg - vr).max() < 1e-12
    assert np.abs(cg - cr).max() < 1e-10 and np.abs(eg - er).max() < 1e-10
    # and against the unsharded C-ABI step on the same inputs
    c1, e1, v1 = ctx.ctmrgstep(bulk, c, e)
    assert np.abs(vg - v1).max() < 1e-12 and np.abs(cg - c1).max() < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("D,chi", [(4, 8), (9, 12), (4, 40), (4, 160)])
def test_ctmrgstep_sharded_library_entry_single_rank(D, chi):
    """tnad_ctmrgstep_sharded (the chi-sharded step inside the library; NCCL only for world > 1, exercised by
    bench_sharded.py --check on 2+ GPUs) against the oracle and the unsharded entry point."""
    c2 = T.Context(0)
    try:
        c2.comm_init(None, 0, 1)
        rng = np.random.default_rng(D * 1000 + chi)
        bulk = rng.standard_normal((D, D, D, D))
        bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
        c, e = O.init_random(bulk, chi, rng)
        cg, eg, vg, ms = c2.ctmrgstep_sharded(bulk, c, e, timing=True)
        c1, e1, v1 = c2.ctmrgstep(bulk, c, e)
        assert np.abs(vg - v1).max() < 1e-12 and np.abs(cg - c1).max() < 1e-11 and np.abs(eg - e1).max() < 1e-11
        if chi * D <= 160:
            cr, er, vr = O.ctmrgstep(bulk, c, e, signfix=True)
            assert np.abs(vg - vr).max() < 1e-12
            assert np.abs(cg - cr).max() < 1e-10 and np.abs(eg - er).max() < 1e-10
        assert len(ms) == 3 and ms[2] > 0.0
        # the fixed-point loop over the same step: same number of steps and spectrum as tnad_ctmrg
        cl, el, vl, nl = c2.ctmrg_sharded(bulk, c, e, 1e-8, 30)
        cu, eu, vu, nu = c2.ctmrg(bulk, c, e, 1e-8, 30)
        assert nl == nu and np.abs(vl - vu).max() < 1e-10 and np.abs(cl - cu).max() < 1e-8
    finally:
        c2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chi,maxit", [(8, 5), (40, 3), (160, 2)])
def test_energy_and_gradient_through_the_sharded_loop(ctx, chi, maxit):
    """A context that joined a communicator runs its CTMRG steps through ctmrg_step_sharded and records them for the
    unrolled reverse sweep: energy and gradient must equal those of a plain context (world 1 here; 2+ GPUs:
    bench_sharded.py --check-energy)."""
    h = T.hamiltonian(T.Heisenberg())
    A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(chi).standard_normal((2, 2, 2, 2, 2)))).bulk
    e0, g0 = ctx.energy(h, A, chi, 0.0, maxit, grad=True)
    c2 = T.Context(0)
    try:
        c2.comm_init(None, 0, 1)
        c2.set_option("TNAD_SHARDED_LOOP", "1")      # opt-in: the call is collective over the communicator
        e1, g1 = c2.energy(h, A, chi, 0.0, maxit, grad=True)
    finally:
        c2.close()
    assert abs(e1 - e0) <= 1e-12 * abs(e0) and np.abs(g1 - g0).max() <= 1e-10 * np.abs(g0).max()


@pytest.mark.gpu
def test_permute_and_svd_symmetrized_device_pointers(ctx):
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 3, 4, 6))
    dev = torch.device("cuda", ctx.device)
    src = torch.from_numpy(np.ascontiguousarray(x.ravel(order="F"))).to(dev)
    dst = torch.empty_like(src)
    ctx.set_pointer_mode(1)
    try:
        ctx.dev_permute(src.data_ptr(), x.shape, (2, 0, 3, 1), dst.data_ptr())
        n = 40
        a = rng.standard_normal((n, n))
        A = torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).to(dev)
        U, S, V = torch.empty(n * n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n * n, dtype=torch.float64, device=dev)
        ctx.dev_svd_symmetrized(A.data_ptr(), n, U.data_ptr(), S.data_ptr(), V.data_ptr())
    finally:
        ctx.set_pointer_mode(0)
    got = dst.cpu().numpy().reshape((4, 5, 6, 3), order="F")
    assert np.array_equal(got, np.transpose(x, (2, 0, 3, 1)))
    u = U.cpu().numpy().reshape((n, n), order="F"); v = V.cpu().numpy().reshape((n, n), order="F"); s = S.cpu().numpy()
    assert np.abs(u * s @ v.T - (a + a.T)).max() < 1e-12
    assert np.abs(s - np.linalg.svd(a + a.T, compute_uv=False)).max() < 1e-12


# ---- direct symmetric eigensolver: tridiagonalisation (tridiag.cu) + divide and conquer (stedc.cu) -----------------
def _tridiag(d, e):
    return np.diag(d) + np.diag(e, 1) + np.diag(e, -1)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 65, 130, 257, 700])
def test_stedc_random_vs_lapack(ctx, n):
    rng = np.random.default_rng(n)
    d, e = rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))
    lam, Z = ctx.stedc(d, e)
    Tm = _tridiag(d, e) if n > 1 else np.array([[d[0]]])
    ref = np.linalg.eigvalsh(Tm)
    nrm = max(np.abs(ref).max(), 1e-300)
    assert np.all(np.diff(lam) >= 0)
    assert np.abs(lam - ref).max() <= 5e-14 * nrm
    assert np.abs(Tm @ Z - Z * lam).max() <= 5e-14 * nrm
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= 5e-14


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["graded", "clustered", "wilkinson", "zero_offdiag", "constant"])
def test_stedc_hard_spectra(ctx, kind):
    rng = np.random.default_rng(11)
    n = 300
    if kind == "graded":
        d, e = 10.0 ** (-rng.uniform(0, 18, n)), 10.0 ** (-rng.uniform(0, 18, n - 1))
    elif kind == "clustered":     # tridiagonal of a matrix with three huge eigenvalue clusters: deflation by Givens chains
        import scipy.linalg as sl
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        w = np.concatenate([np.ones(100), np.zeros(100), -np.ones(100)]) + 1e-14 * rng.standard_normal(n)
        h = sl.hessenberg((q * w) @ q.T)
        d, e = np.diag(h).copy(), np.diag(h, 1).copy()
    elif kind == "wilkinson":
        d, e = np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
    elif kind == "zero_offdiag":
        d, e = rng.standard_normal(n), np.zeros(n - 1)
    else:
        d, e = np.full(n, 2.0), np.full(n - 1, -1.0)
    lam, Z = ctx.stedc(d, e)
    Tm = _tridiag(d, e)
    ref = np.linalg.eigvalsh(Tm)
    nrm = np.abs(ref).max()
    assert np.abs(lam - ref).max() <= 1e-13 * nrm
    assert np.abs(Tm @ Z - Z * lam).max() <= 1e-13 * nrm
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= 1e-13


def _sym_cases(n, rng):
    a = rng.standard_normal((n, n))
    yield "random", a + a.T
    b = rng.standard_normal((n, max(2, n // 10)))
    yield "low_rank", b @ b.T
    q, _ = np.linalg.qr(a)
    m = (q * (10.0 ** (-np.arange(n) / 6.0) * rng.choice([-1.0, 1.0], n))) @ q.T
    yield "decaying", m + m.T


@pytest.mark.gpu
@pytest.mark.parametrize("one_barrier", ["1", "0"])
@pytest.mark.parametrize("n", [3, 4, 10, 33, 97, 130, 259, 600])
def test_sytrd_backward_error(ctx, n, one_barrier, opt):
    """A = Q T Q' to machine precision for both panel kernels, including rank-deficient input (the cancellation
    guard of the one-barrier kernel) and sizes that are odd / not multiples of the ownership quad."""
    opt("TNAD_SYTRD_1B", one_barrier)
    rng = np.random.default_rng(100 + n)
    for name, a in _sym_cases(n, rng):
        d, e, q = ctx.sytrd(a)
        Tm = _tridiag(d, e)
        nrm = np.linalg.norm(a, 2)
        assert np.abs(q.T @ q - np.eye(n)).max() <= 2e-14, (name, n)
        assert np.abs(q @ Tm @ q.T - a).max() <= 2e-14 * nrm, (name, n)
        assert np.abs(np.linalg.eigvalsh(Tm) - np.linalg.eigvalsh(a)).max() <= 2e-14 * nrm, (name, n)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("n", [64, 300, 520])
def test_svd_sym_both_solvers(ctx, n, mode, opt):
    """tnad_svd_sym through the block-Jacobi solver (1) and the tridiagonal divide-and-conquer solver (2)."""
    opt("TNAD_SYMEIG", mode)
    rng = np.random.default_rng(n)
    for name, a in _sym_cases(n, rng):
        u, s, v = ctx.svd_sym(a)
        ref = np.linalg.svd(a, compute_uv=False)
        assert np.all(np.diff(s) <= 0)
        assert np.abs(s - ref).max() <= 1e-12 * ref[0], (name, mode)
        assert np.abs((u * s) @ v.T - a).max() <= 1e-12 * ref[0], (name, mode)
        assert np.abs(u.T @ u - np.eye(n)).max() <= 1e-12, (name, mode)
        assert np.abs(np.abs(np.sum(u * v, axis=0)) - 1.0).max() <= 1e-12      # v = +-u column by column


@pytest.mark.gpu
def test_energy_gradient_same_with_both_solvers(ctx, opt):
    """The headline path (energy + gradient) must not depend on which eigensolver ran (chi*D = 256 >= TNAD_DC_MIN)."""
    rng = np.random.default_rng(5)
    h = T.hamiltonian(T.Heisenberg())
    A = T.indexperm_symmetrize(T.SquareIPEPS(rng.standard_normal((2, 2, 2, 2, 2)))).bulk
    out = {}
    for mode in ("1", "2"):
        opt("TNAD_SYMEIG", mode)
        out[mode] = ctx.energy(h, A, 64, 0.0, 4, grad=True)
    e1, g1 = out["1"]
    e2, g2 = out["2"]
    assert abs(e1 - e2) <= 1e-10 * abs(e1)
    assert np.abs(g1 - g2).max() <= 1e-8 * np.abs(g1).max()
    er, gr = O.energy_value_and_grad(h, A, 64, 0.0, 4)
    assert abs(e2 - er) <= 1e-10 * abs(er) and np.abs(g2 - gr).max() <= 1e-8 * np.abs(gr).max()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["zero", "identity", "diag_repeated", "tiny_scale", "denormal_squares", "huge_scale", "rank_one", "block_diag"])
def test_svd_sym_direct_solver_degenerate_inputs(ctx, kind, opt):
    """Inputs on which every reflector / merge degenerates (tau = 0, rho = 0, full deflation)."""
    opt("TNAD_SYMEIG", "2")
    rng = np.random.default_rng(9)
    n = 150
    if kind == "zero":
        a = np.zeros((n, n))
    elif kind == "identity":
        a = np.eye(n)
    elif kind == "diag_repeated":
        a = np.diag(np.repeat([3.0, -1.0, 0.5], n // 3))
    elif kind == "tiny_scale":
        b = rng.standard_normal((n, n))
        a = 1e-120 * (b + b.T)
    elif kind == "denormal_squares":      # squares of the entries underflow
        b = rng.standard_normal((n, n))
        a = 1e-200 * (b + b.T)
    elif kind == "huge_scale":            # squares of the entries overflow
        b = rng.standard_normal((n, n))
        a = 1e200 * (b + b.T)
    elif kind == "rank_one":
        v = rng.standard_normal(n)
        a = np.outer(v, v)
    else:
        a = np.zeros((n, n))
        b = rng.standard_normal((50, 50))
        a[:50, :50] = b + b.T
        a[100:, 100:] = np.eye(50) * 2.0
    u, s, v = ctx.svd_sym(a)
    ref = np.linalg.svd(a, compute_uv=False)
    scale = max(ref[0], 1e-300)
    assert np.all(np.isfinite(u)) and np.all(np.isfinite(s)) and np.all(np.isfinite(v))
    assert np.abs(s - ref).max() <= 1e-12 * scale
    assert np.abs((u * s) @ v.T - a).max() <= 1e-12 * scale
    assert np.abs(u.T @ u - np.eye(n)).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("beta", [0.3, 0.44, 0.6])
def test_trg_same_with_both_svd_routes(ctx, beta, opt):
    """TRG value + gradient through the one-sided Jacobi SVD and through the Jordan-Wielandt / direct-eigensolver route
    (rank-deficient splits: the null triplets enter svd_back only through the projectors)."""
    a = T.model_tensor(T.Ising(), beta)
    out = {}
    for mode in ("jacobi", "dc"):
        opt("TNAD_TRG_SVD", mode)
        out[mode] = T.trg_value_and_grad(a, 12, 9, ctx=ctx)
    (l1, g1), (l2, g2) = out["jacobi"], out["dc"]
    assert abs(l1 - l2) <= 1e-12 * abs(l1)
    # full gradient tensor: components along rank-lifting / degenerate-pair directions amplify 1e-15 differences of the
    # singular vectors by 1/(s_i^2 - s_j^2) (the same directions in which oracle and Zygote golden differ by percents);
    # the physical derivative must agree tightly
    assert np.abs(g1 - g2).max() <= 1e-6 * np.abs(g1).max()
    da_ = O.dmodel_tensor_ising(beta)
    assert abs(np.sum(g1 * da_) - np.sum(g2 * da_)) <= 1e-10 * abs(np.sum(g1 * da_))
    # against the oracle: the value and the physical derivative d lnZ / d beta (directions that lift the rank of a split
    # are ill-defined in the reference itself, see test_trg_gradient_published)
    lo, go = O.trg_value_and_grad(a, 12, 9)
    da = O.dmodel_tensor_ising(beta)
    assert abs(l2 - lo) <= 1e-10 * abs(lo)
    assert abs(np.sum(g2 * da) - np.sum(go * da)) <= 1e-8 * abs(np.sum(go * da))


# ---- two-stage tridiagonalisation (band.cu): dense -> band (cluster panel QR + DMMA updates) -> tridiagonal (systolic chase)
def _band_dense(band, n, b=32):
    B = np.zeros((n, n))
    for j in range(n):
        for o in range(min(b, n - 1 - j) + 1):
            B[j + o, j] = band[o, j]
            B[j, j + o] = band[o, j]
    return B


@pytest.mark.gpu
@pytest.mark.parametrize("n", [3, 5, 33, 34, 35, 40, 65, 66, 97, 130, 259, 600, 1100])
def test_sytrd2_backward_error(ctx, n):
    """A = Q T Q' to machine precision through the two-stage route, incl. rank-deficient and graded input, sizes around
    the panel / window boundaries (33, 34, 65, 66) and several chase CTAs (n = 1100: 35 positions)."""
    rng = np.random.default_rng(200 + n)
    for name, a in _sym_cases(n, rng):
        d, e, q, band = ctx.sytrd2(a)
        Tm = _tridiag(d, e)
        nrm = np.linalg.norm(a, 2)
        assert np.abs(q.T @ q - np.eye(n)).max() <= 3e-14, (name, n)
        assert np.abs(q @ Tm @ q.T - a).max() <= 3e-14 * nrm, (name, n)
        ev = np.linalg.eigvalsh(a)
        assert np.abs(np.linalg.eigvalsh(Tm) - ev).max() <= 3e-14 * nrm, (name, n)
        assert np.abs(np.linalg.eigvalsh(_band_dense(band, n)) - ev).max() <= 3e-14 * nrm, (name, n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [100, 300, 520, 1000])
def test_svd_sym_two_stage_route(ctx, n, opt):
    opt("TNAD_SYMEIG", "2")
    opt("TNAD_EIG_2STAGE", "1")
    rng = np.random.default_rng(n)
    for name, a in _sym_cases(n, rng):
        u, s, v = ctx.svd_sym(a)
        ref = np.linalg.svd(a, compute_uv=False)
        assert np.all(np.diff(s) <= 0)
        assert np.abs(s - ref).max() <= 1e-12 * ref[0], name
        assert np.abs((u * s) @ v.T - a).max() <= 1e-12 * ref[0], name
        assert np.abs(u.T @ u - np.eye(n)).max() <= 1e-12, name


@pytest.mark.gpu
def test_energy_gradient_same_with_one_and_two_stage(ctx, opt):
    """Energy + gradient must not depend on the tridiagonalisation route (chi D = 256)."""
    rng = np.random.default_rng(6)
    h = T.hamiltonian(T.Heisenberg())
    A = T.indexperm_symmetrize(T.SquareIPEPS(rng.standard_normal((2, 2, 2, 2, 2)))).bulk
    out = {}
    for mode in ("0", "1"):
        opt("TNAD_EIG_2STAGE", mode)
        out[mode] = ctx.energy(h, A, 64, 0.0, 4, grad=True)
    (e1, g1), (e2, g2) = out["0"], out["1"]
    assert abs(e1 - e2) <= 1e-11 * abs(e1) and np.abs(g1 - g2).max() <= 1e-9 * np.abs(g1).max()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [150, 700])
def test_symeig_three_phase_matches_single_call(ctx, n):
    """tnad_symeig_reduce / _backtransform (two column blocks, as two ranks would do) / _finish against tnad_svd_symmetrized."""
    import torch
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    dev = torch.device("cuda", ctx.device)
    A = torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).to(dev)
    U = torch.empty(n * n, dtype=torch.float64, device=dev); V = torch.empty_like(U)
    S = torch.empty(n, dtype=torch.float64, device=dev)
    ctx.set_pointer_mode(1)
    try:
        h, N = ctx.dev_symeig_reduce(A.data_ptr(), n, True)
        Z = torch.empty(N * N, dtype=torch.float64, device=dev)
        half = (N // 2 // 2) * 2
        ctx.dev_symeig_backtransform(h, 0, half, Z.data_ptr())
        ctx.dev_symeig_backtransform(h, half, N - half, Z.data_ptr() + 8 * N * half)
        ctx.dev_symeig_finish(h, Z.data_ptr(), U.data_ptr(), S.data_ptr(), V.data_ptr())
        ctx.symeig_free(h)
    finally:
        ctx.set_pointer_mode(0)
    u = U.cpu().numpy().reshape((n, n), order="F"); v = V.cpu().numpy().reshape((n, n), order="F"); s = S.cpu().numpy()
    m = a + a.T
    ref = np.linalg.svd(m, compute_uv=False)
    assert np.abs(s - ref).max() <= 1e-12 * ref[0]
    assert np.abs((u * s) @ v.T - m).max() <= 1e-12 * ref[0]
    u1, s1, v1 = ctx.svd_sym(m)
    assert np.abs(s - s1).max() <= 1e-13 * ref[0] and np.abs(u - u1).max() < 1e-9     # same canonical gauge


@pytest.mark.gpu
def test_fixedpoint_gradient_matches_unrolled_gradient_at_convergence(ctx):
    """Opt-in implicit gradient (tnad_energy_fixedpoint): for a converged CTMRG it must agree with the unrolled reverse
    sweep, i.e. with what the reference computes (test/variationalipeps.jl:121-134 runs tol=0, maxit=100)."""
    h = T.hamiltonian(T.Heisenberg())
    agreed = 0
    for seed, chi in [(0, 8), (1, 8), (2, 10), (3, 12)]:
        A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(seed).standard_normal((2, 2, 2, 2, 2)))).bulk
        e_u, g_u = T.energy_and_gradient(h, A, chi, 0.0, 150, ctx=ctx)
        try:
            e_f, g_f = T.energy_and_gradient_fixedpoint(h, A, chi, 0.0, 150, bwd_tol=1e-13, bwd_maxit=400, ctx=ctx)
        except T.TnadError as ex:
            # degenerate multiplet inside the kept spectrum: the environment converges only up to rotations and the
            # library refuses the implicit formula (error 4) instead of returning a wrong gradient
            assert ex.code == 4
            continue
        assert ctx.last_bwd_iters > 3
        assert e_f == pytest.approx(e_u, rel=1e-10)
        assert rel(g_f, g_u) < 1e-7
        agreed += 1
    assert agreed >= 1
    A = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))).bulk
    eo, go = O.energy_value_and_grad(h, A, 8, 0.0, 150)
    e_f, g_f = T.energy_and_gradient_fixedpoint(h, A, 8, 0.0, 150, bwd_tol=1e-13, bwd_maxit=400, ctx=ctx)
    assert rel(g_f, go) < 1e-7 and e_f == pytest.approx(eo, rel=1e-10)


# ---- optimiseipeps: the remaining known answers of the reference's test-suite (test/variationalipeps.jl:42-118) --------
def _classical_ising_h():
    h = np.zeros((2, 2, 2, 2))
    h[0, 0, 1, 1] = h[1, 1, 0, 0] = 1.0
    h[1, 1, 1, 1] = h[0, 0, 0, 0] = -1.0
    return h


OPT_CASES = [
    # name, hamiltonian builder, chi, tol, maxit, f_tol, expected minimum, atol   (source line in test/variationalipeps.jl)
    ("ising", lambda: _classical_ising_h(), 4, 0.0, 100, 1e-6, -1.0, 1e-3),                                  # :44-52
    ("tfising_1.0", lambda: T.hamiltonian(T.TFIsing(1.0)), 5, 0.0, 100, 1e-6, -2.12566, 1e-3),               # :67-73
    ("tfising_0.5", lambda: T.hamiltonian(T.TFIsing(0.5)), 5, 0.0, 100, 1e-6, -2.0312, 1e-2),                # :75-81
    ("tfising_2.0", lambda: T.hamiltonian(T.TFIsing(2.0)), 6, 1e-9, 100, 1e-8, -2.5113, 1e-3),               # :83-90
    ("heisenberg_2_2_1", lambda: T.hamiltonian(T.Heisenberg(2.0, 2.0, 1.0)), 6, 0.0, 100, 1e-6, -1.190, 1e-2),   # :104-110
    ("heisenberg_.5_.5_2", lambda: T.hamiltonian(T.Heisenberg(0.5, 0.5, 2.0)), 5, 0.0, 100, 1e-6, -1.0208, 1e-3),  # :112-118
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,hb,chi,tol,maxit,ftol,expected,atol", OPT_CASES, ids=[c[0] for c in OPT_CASES])
def test_optimiseipeps_known_answers(ctx, name, hb, chi, tol, maxit, ftol, expected, atol):
    """L-BFGS over the fused energy + gradient call reaches the energies the reference's tests expect.  The reference pins
    Julia's global RNG (Random.seed!) to land in the right basin; here the best of three NumPy seeds must reach it (the
    variational landscape has local minima: with the oracle, seed 0 reaches all six values)."""
    best = np.inf
    for seed in (0, 1, 2):
        ipeps = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(seed).standard_normal((2, 2, 2, 2, 2))))
        res = T.optimiseipeps(ipeps, hb(), chi=chi, tol=tol, maxit=maxit, optimargs={"f_tol": min(ftol, 1e-9), "iterations": 200}, ctx=ctx)
        best = min(best, res.minimum)
        if abs(best - expected) < atol:
            break
    assert abs(best - expected) < atol, (name, best)
    assert best > expected - 10 * atol           # variational: never below the converged literature value


@pytest.mark.gpu
def test_optimiseipeps_rotated_ising(ctx):
    # test/variationalipeps.jl:54-64: the classical Hamiltonian conjugated with a random orthogonal matrix on every leg
    rng = np.random.default_rng(4)
    u, _, _ = np.linalg.svd(rng.standard_normal((2, 2)))
    h = np.einsum("abcd,ai,bj,ck,dl->ijkl", _classical_ising_h(), u, u.T, u, u.T)
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2))))
    res = T.optimiseipeps(ipeps, h, chi=6, tol=0.0, maxit=200, optimargs={"f_tol": 1e-9, "iterations": 200}, ctx=ctx)
    assert abs(res.minimum + 1.0) < 1e-3


@pytest.mark.gpu
def test_readme_optimisation_config_time_per_iteration(ctx):
    """The only timing the reference publishes (README.md:117-153): optimiseipeps, Heisenberg d=2, chi=20, tol=1e-6,
    maxit=100, f_tol=1e-6: 16 L-BFGS iterations in 4.84 s (0.30 s per iteration, unspecified 2019 CPU), final energy
    -0.6602311.  Here: same configuration through the C ABI; the energy must agree, the time per iteration is recorded."""
    import time
    h = T.hamiltonian(T.Heisenberg())
    ipeps = T.indexperm_symmetrize(T.SquareIPEPS(np.random.default_rng(2).random((2, 2, 2, 2, 2))))
    t0 = time.perf_counter()
    res = T.optimiseipeps(ipeps, h, chi=20, tol=1e-6, maxit=100, optimargs={"f_tol": 1e-6, "iterations": 100}, ctx=ctx)
    dt = time.perf_counter() - t0
    assert abs(res.minimum + 0.6602311) < 1e-3
    per_it = dt / max(res.nit, 1)
    print(f"README config: {res.nit} L-BFGS iterations, {res.nfev} energy+gradient calls, {dt:.2f} s -> {per_it * 1e3:.0f} ms / iteration "
          f"(reference README: 300 ms / iteration on a 2019 CPU)")
    assert per_it < 0.30
