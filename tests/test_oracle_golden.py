"""CPU: pins the oracle (oracle/tnad_oracle.py) against every golden value, analytic test and adjoint check
the reference holds for the hot path.  No GPU, no product code."""
import numpy as np
import pytest

import tnad_oracle as O


# ---- exact goldens ---------------------------------------------------------------------------------
def test_trg_golden_tensorgrad(golden):
    pub, _ = golden
    assert O.trg(O.model_tensor_ising(0.4), 5, 5) == pytest.approx(pub["trg_beta0.4_chi5_n5"]["value"], rel=1e-12)


def test_trg_readme_value_and_gradient(golden):
    pub, _ = golden
    lnz, db = O.trg_dbeta(0.5, 20, 20)
    assert lnz == pytest.approx(pub["trg_beta0.5_chi20_n20"]["value"], rel=1e-12)
    assert db == pytest.approx(pub["dtrg_beta0.5_chi20_n20"]["value"], rel=1e-10)


def test_trg_userguide_gradient(golden):
    pub, _ = golden
    assert O.trg_dbeta(0.5, 5, 5)[1] == pytest.approx(pub["dtrg_beta0.5_chi5_n5"]["value"], rel=1e-10)


def test_trg_gradient_vs_numgrad():
    # test/trg.jl:19
    f = lambda b: O.trg(O.model_tensor_ising(b), 5, 5)  # noqa: E731
    assert O.num_grad(f, 0.4, 1e-6) == pytest.approx(O.trg_dbeta(0.4, 5, 5)[1], rel=2e-8)


def test_trg_svd_unit():
    # test/trg.jl:6-10
    t = np.random.default_rng(0).standard_normal((10, 10, 10, 10))
    u, v, _ = O.trg_svd(t, 100, 0)
    assert np.allclose(np.einsum("ija,akl->ijkl", u, v), t, atol=1e-11)


def test_gesdd_vs_gesvd_noise_floor():
    a = O.trg_dbeta(0.5, 8, 8, driver="gesdd")
    b = O.trg_dbeta(0.5, 8, 8, driver="gesvd")
    assert abs(a[0] - b[0]) < 1e-13 and abs(a[1] - b[1]) / abs(a[1]) < 1e-10


def test_heisenberg_hamiltonian_readme(golden):
    pub, _ = golden
    h = O.hamiltonian_heisenberg()
    exp = pub["heisenberg_h"]["expected_print"]
    assert np.allclose(h[:, :, 0, 0], exp["[:,:,1,1]"]) and np.allclose(h[:, :, 1, 0], exp["[:,:,2,1]"])
    assert np.allclose(h[:, :, 0, 1], exp["[:,:,1,2]"]) and np.allclose(h[:, :, 1, 1], exp["[:,:,2,2]"])


def test_model_tensor_vs_classical():
    # test/exampletensors.jl:6-9
    b = 0.37
    assert np.allclose(O.model_tensor_ising(b), O.tensorfromclassical([[b, -b], [-b, b]]), atol=1e-14)


def test_dmodel_tensor():
    f = lambda b: O.model_tensor_ising(b)  # noqa: E731
    fd = (f(0.5 + 1e-6) - f(0.5 - 1e-6)) / 2e-6
    assert np.allclose(O.dmodel_tensor_ising(0.5), fd, atol=1e-8)


# ---- CTMRG -------------------------------------------------------------------------------------------
def test_ctmrgstep_literal_equals_pairwise():
    rng = np.random.default_rng(1)
    bulk = rng.standard_normal((3, 3, 3, 3))
    c, e = O.init_random(bulk, 5, rng)
    c1, e1, v1 = O.ctmrgstep_literal(bulk, c, e)
    c2, e2, v2 = O.ctmrgstep(bulk, c, e)
    assert np.allclose(np.abs(c1), np.abs(c2), atol=1e-13) and np.allclose(np.abs(e1), np.abs(e2), atol=1e-13)
    assert np.allclose(v1, v2, atol=1e-14)


def test_stop_rule_counter_starts_at_minus_one():
    # ctmrg.jl:114 + fixedpoint.jl:31-41: tol = 0 => exactly maxit + 1 steps
    a = O.model_tensor_ising(0.3)
    c0, e0 = O.init_raw(a, 4)
    for maxit in (0, 1, 5):
        assert O.ctmrg(a, c0, e0, 0.0, maxit)[3] == maxit + 1


def test_init_raw_docstring():
    # ctmrg.jl:53-59 doctest
    bulk = np.random.default_rng(3).standard_normal((2, 2, 2, 2))
    c, e = O.init_raw(bulk, 4)
    assert np.allclose(c[:2, :2], bulk.sum(axis=(2, 3))) and np.allclose(e[:2, :, :2], bulk.sum(axis=3))
    c, e = O.init_raw(np.random.default_rng(3).standard_normal((5, 5, 5, 5)), 3)   # D > chi truncates
    assert c.shape == (3, 3) and e.shape == (3, 5, 3)


@pytest.mark.parametrize("beta,chi,atol", [(1.0, 2, 1e-8), (0.6, 4, 1e-8), (0.8, 2, 1e-8), (0.2, 10, 1e-4)])
def test_onsager_magnetisation(beta, chi, atol):
    # test/ctmrg.jl:37-42
    m = O.magnetisation(beta, chi, np.random.default_rng(5))
    assert abs(m - O.magofbeta(beta)) < atol


# ---- energy ------------------------------------------------------------------------------------------
def test_noninteracting_energies():
    # test/variationalipeps.jl:9-25
    rng = np.random.default_rng(0)
    h = O.diaglocalhamiltonian([1, -1.0])
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 1] = rng.standard_normal()
    assert O.energy(h, a, 4, 1e-12, 100) / 2 == pytest.approx(-1.0, abs=1e-9)
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 0] = rng.standard_normal()
    assert O.energy(h, a, 10, 0, 300) / 2 == pytest.approx(1.0, abs=1e-9)
    a = 1e-12 * rng.standard_normal((2, 2, 2, 2, 2)); a[0, 0, 0, 0, 1] = a[0, 0, 0, 0, 0] = rng.standard_normal()
    assert abs(O.energy(h, a, 10, 0, 300)) < 1e-9
    for _ in range(5):
        assert -1 < O.energy(h, rng.random((3, 3, 3, 3, 2)), 5, 0, 10) / 2 < 1


def test_energy_gradient_vs_numgrad():
    # test/variationalipeps.jl:121-134
    h = O.hamiltonian_heisenberg()
    A = O.indexperm_symmetrize(np.random.default_rng(0).standard_normal((2, 2, 2, 2, 2)))
    _, g = O.energy_value_and_grad(h, A, 4, 0, 100)
    gn = O.num_grad(lambda x: O.energy(h, x, 4, 0, 100), A, 1e-3)
    assert np.allclose(g, gn, atol=1e-3)
    assert np.abs(g - gn).max() < 1e-5


def test_energy_fixture_is_reproduced(golden):
    _, vec = golden
    h = O.hamiltonian_heisenberg()
    e, g = O.energy_value_and_grad(h, vec["c3_A"], 20, 1e-6, 100)
    assert e == pytest.approx(float(vec["c3_e"]), rel=1e-12)
    assert np.linalg.norm(g - vec["c3_grad"]) / np.linalg.norm(g) < 1e-9


# ---- adjoints ------------------------------------------------------------------------------------------
def _fd_directional(f, A, dA, eps=1e-6):
    return (f(A + eps * dA) - f(A - eps * dA)) / (2 * eps)


@pytest.mark.parametrize("shape", [(6, 3), (3, 6), (3, 3)])
def test_svd_back_real(shape):
    # test/svd.jl:15-101 (real restriction): losses of U, V and S
    rng = np.random.default_rng(4)
    m, n = shape
    A = rng.standard_normal(shape)
    H1 = rng.standard_normal((m, m)); H1 = H1 + H1.T
    H2 = rng.standard_normal((n, n)); H2 = H2 + H2.T
    w = rng.standard_normal(min(m, n))

    def loss(A):
        U, S, V = O.svd(A)
        return U[:, 0] @ H1 @ U[:, 0] + V[:, 0] @ H2 @ V[:, 0] + w @ S

    U, S, V = O.svd(A)
    dU = np.zeros_like(U); dU[:, 0] = 2 * H1 @ U[:, 0]
    dV = np.zeros_like(V); dV[:, 0] = 2 * H2 @ V[:, 0]
    g = O.svd_back(U, S, V, dU, w, dV)
    dA = rng.standard_normal(shape)
    assert np.sum(g * dA) == pytest.approx(_fd_directional(loss, A, dA), rel=1e-6)


def test_svd_back_complex_imag_diag():
    # test/svd.jl:69-86
    def loss(A):
        U, S, V = O.svd(A)
        return np.real(np.conj(U[0, 0]) * V[0, 0])
    A = np.array([[-1 + 1j, 2 + 1j], [1 - 2j, 3 + 0.8j]])
    U, S, V = O.svd(A)
    dU = np.zeros_like(U); dU[0, 0] = V[0, 0]
    dV = np.zeros_like(V); dV[0, 0] = U[0, 0]
    g = O.svd_back(U, S, V, dU, None, dV)
    da = np.array([[0, 0], [1, 0]], dtype=complex)
    nd = (loss(A + 1e-4 * da) - loss(A - 1e-4 * da)) / 2e-4 + 1j * (loss(A + 1e-4j * da) - loss(A - 1e-4j * da)) / 2e-4
    assert abs(g[1, 0] - nd) < 1e-3


def test_norm_pullback():
    # test/autodiff.jl:6-8 with the rule of autodiff.jl:23-29
    a = np.random.default_rng(0).standard_normal((10, 10))
    n = np.linalg.norm(a)
    assert np.allclose(a / n, O.num_grad(lambda x: np.linalg.norm(x), a), atol=1e-8)
    x2 = a * 3.0
    ybar = np.random.default_rng(1).standard_normal(a.shape)
    got = O._norm_back(ybar, x2, np.linalg.norm(x2))
    f = lambda x: np.sum(ybar * x / np.linalg.norm(x))  # noqa: E731
    assert np.allclose(got, O.num_grad(f, x2), atol=1e-7)


# ---- round 2 additions: canonical gauge, magnetisation pullback, large-size fixtures -------------------------------
def test_canonical_gauge_keeps_the_decomposition():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((9, 9))
    U, S, V = O.svd(A + A.T)
    U2, V2 = O.canonical_gauge(U, V)
    assert np.abs((U2 * S) @ V2.T - (A + A.T)).max() < 1e-13
    top = U2[np.argmax(np.abs(U2), axis=0), np.arange(9)]
    assert np.all(top > 0)
    # the step is covariant under the gauge: invariants agree, corner differs at most by signs
    bulk = rng.standard_normal((2, 2, 2, 2)); bulk = bulk + np.transpose(bulk, (2, 3, 0, 1))
    c, e = O.init_random(bulk, 4, rng)
    c1, e1, v1 = O.ctmrgstep(bulk, c, e)
    c2, e2, v2 = O.ctmrgstep(bulk, c, e, signfix=True)
    assert np.abs(v1 - v2).max() < 1e-14 and np.abs(np.abs(c1) - np.abs(c2)).max() < 1e-13


def test_magnetisation_pullback_against_finite_differences():
    # test/ctmrg.jl:44-46 differentiates magnetisation; here every piece of the oracle's reverse sweep is checked
    rng = np.random.default_rng(9)
    a, m = O.model_tensor_ising(0.5), O.mag_tensor_ising(0.5)
    c, e = O.init_random(a, 4, rng)
    ab, mb, cb, eb = O.magnetisation_readout_back(a, m, c, e)
    dc, de = rng.standard_normal(c.shape), rng.standard_normal(e.shape)
    da, dm = rng.standard_normal(a.shape), rng.standard_normal(a.shape)
    eps = 1e-6
    fd = (O.magnetisation_readout(a + eps * da, m + eps * dm, c + eps * dc, e + eps * de)
          - O.magnetisation_readout(a - eps * da, m - eps * dm, c - eps * dc, e - eps * de)) / (2 * eps)
    assert fd == pytest.approx(np.sum(ab * da) + np.sum(mb * dm) + np.sum(cb * dc) + np.sum(eb * de), rel=1e-7)
    h = 1e-6
    assert np.abs((O.mag_tensor_ising(0.5 + h) - O.mag_tensor_ising(0.5 - h)) / (2 * h) - O.dmag_tensor_ising(0.5)).max() < 1e-8
    c0, e0 = O.init_random(a, 2, np.random.default_rng(9))
    y, g = O.magnetisation_value_and_dbeta(0.5, 2, c0, e0, tol=1e-10, maxit=400)

    def f(b):
        ab_, mb_ = O.model_tensor_ising(b), O.mag_tensor_ising(b)
        cc, ee, _, _ = O.ctmrg(ab_, c0, e0, 1e-10, 400)
        return O.magnetisation_readout(ab_, mb_, cc, ee)
    assert abs(g - (f(0.5 + 5e-4) - f(0.5 - 5e-4)) / 1e-3) < 1e-2       # the reference's own tolerance


def test_large_fixture_is_consistent_with_known_answers():
    import os
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large.npz"))
    # CTMRG chi=64 with a :random environment reproduces Onsager's magnetisation above beta_c (test/ctmrg.jl:37-42)
    assert float(vec["c2_random_0.5_mag"]) == pytest.approx(O.magofbeta(0.5), abs=1e-10)
    assert float(vec["c2_random_0.3_mag"]) < 1e-10
    # LAPACK's two SVD drivers agree on step count and spectrum at chi=64 (the fixture's noise floor)
    for beta in (0.3, 0.5):
        assert int(vec[f"c2_raw_{beta}_steps"]) == int(vec[f"c2_raw_{beta}_steps_gesvd"])
        assert float(vec[f"c2_raw_{beta}_vals_spread"]) < 1e-13
    # TRG chi=64, 7 iterations at beta = 0.44 sits between the chi=20 published values' neighbours (sanity) and
    # d lnZ / d beta is positive
    assert 0.9 < float(vec["trg64_lnz"]) < 1.0 and float(vec["trg64_dbeta"]) > 1.0
    assert int(vec["c4_steps"]) == 4 and vec["c4_grad"].shape == (4, 4, 4, 4, 2)
