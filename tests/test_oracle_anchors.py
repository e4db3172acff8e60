"""CPU: the C oracle against the committed known-answer anchors (tests/golden/anchors.json, mpmath 50 digits)
and against LAPACK.  This is what pins the oracle ("parity unpinned" by the reference's own tests)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [c["name"] for c in json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))["cases"]]


def _gp(c):
    k = O.KernelDesc.make(c["ops"], c["params"])
    return O.OracleGaussianProcess(O.ZeroPrior(), k, c["noise"], None, np.array(c["X"]), np.array(c["y"]))


def test_doctest_anchor_values_match_survey():
    """The §8c literals themselves (independent of the JSON generator)."""
    k = O.KernelDesc.make([O.K_SQUARED_EXP], [1.0, 1.0])
    y = np.array([3.0, 4.0, -2.0, -2.0])
    noise = 0.1 * math.sqrt(O.variance(y))
    assert noise == 0.27726341266023546
    gp = O.OracleGaussianProcess(O.ZeroPrior(), k, noise, None, [[0.8], [1.2], [3.8], [4.2]], y)
    Lref = np.array([[1.037725879025863, 0, 0, 0], [0.8895570256503444, 0.5343812291951435, 0, 0],
                     [0.01070513587718423, 0.04589350180103869, 1.0366552882025388, 0],
                     [0.00297642708027694, 0.01583381761067837, 0.8897439915563061, 0.5338268076737618]])
    assert np.allclose(np.tril(gp.L), Lref, rtol=1e-13, atol=1e-16)
    q = [[1.0], [2.0], [3.0]]
    assert np.allclose(gp.predict(q), [3.43460089928164334, 2.61347089019187292, -0.483309215643327428], rtol=1e-13)
    assert np.allclose(gp.predict_variance(q), [0.0391981303494769655, 0.405545879974281484, 0.405545879974281298], rtol=1e-12)
    m, v = gp.predict_mean_variance(q)
    assert np.allclose(v, [0.0391981303494769655, 0.405545879974281484, 0.405545879974281298], rtol=1e-12)
    assert abs(gp.likelihood() - (-13.8047231440077)) < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_anchor(anchors, name):
    c = anchors[name]
    gp = _gp(c)
    n = len(c["X"])
    L = np.array(c["L"])
    assert np.all(np.isnan(gp.L[np.triu_indices(n, 1)]))  # upper triangle stays NaN (algebra/mod.rs:67)
    assert np.allclose(np.tril(gp.L), L, rtol=1e-11, atol=1e-14)
    Xq = np.array(c["Xq"])
    assert np.allclose(gp.predict(Xq), c["mean"], rtol=1e-9, atol=1e-12)
    assert np.allclose(gp.predict_variance(Xq), c["var"], rtol=1e-8, atol=1e-12)
    m, v = gp.predict_mean_variance(Xq)
    assert np.allclose(m, c["mean"], rtol=1e-9, atol=1e-12)
    assert np.allclose(v, c["var"], rtol=1e-8, atol=1e-12)
    assert np.allclose(gp.predict_covariance(Xq), c["cov"], rtol=1e-8, atol=1e-12)
    mean2, cov2 = gp.sample_at_params(Xq)
    assert np.allclose(cov2, c["cov"], rtol=1e-8, atol=1e-12)
    assert np.allclose(mean2, c["mean"], rtol=1e-9, atol=1e-12)
    assert abs(gp.likelihood() - c["likelihood"]) < 1e-9 * max(1.0, abs(c["likelihood"]))
    g = gp.gradient_marginal_likelihood(scaled=False)
    assert np.allclose(g, c["grad_unscaled"], rtol=1e-7, atol=1e-9)
    s, gs = gp.gradient_marginal_likelihood(scaled=True)
    assert abs(s - c["scale"]) < 1e-9 * abs(c["scale"])
    assert np.allclose(gs, c["grad_scaled"], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_oracle_add_samples_matches_full_refactor(anchors, name):
    c = anchors[name]
    gp = _gp(c)
    gp.add_samples(np.array(c["Xadd"]), np.array(c["yadd"]))
    assert np.allclose(np.tril(gp.L), np.array(c["L_after_add"]), rtol=1e-10, atol=1e-13)


def test_cholesky_against_lapack_and_substitute():
    import scipy.linalg as sl
    rng = np.random.default_rng(7)
    n = 200
    A = rng.standard_normal((n, n))
    K = np.asfortranarray(A @ A.T + n * np.eye(n))
    Lo = K.copy(order="F")
    assert O.cholesky_inplace(Lo) == 0
    Lr = sl.cholesky(K, lower=True)
    assert np.linalg.norm(np.tril(Lo) - Lr) / np.linalg.norm(Lr) < 1e-13
    # failure report and substitute semantics (nalgebra new_with_substitute; algebra/mod.rs:81-91)
    B = np.asfortranarray(np.array([[4.0, 0, 0], [2.0, 1.0, 0], [2.0, 1.0, 3.0]]))  # pivot 1 becomes exactly 0
    assert O.cholesky_inplace(B.copy(order="F")) == 2
    Bs = B.copy(order="F")
    assert O.cholesky_inplace(Bs, substitute=1e-4) == 0
    assert Bs[1, 1] == math.sqrt(1e-4)
    assert O.cholesky_inplace(B.copy(order="F"), substitute=0.0) == 2


def test_solves_against_scipy():
    import scipy.linalg as sl
    rng = np.random.default_rng(3)
    n, q = 150, 7
    A = rng.standard_normal((n, n))
    K = np.asfortranarray(A @ A.T + n * np.eye(n))
    L = K.copy(order="F")
    O.cholesky_inplace(L)
    B = np.asfortranarray(rng.standard_normal((n, q)))
    Y, ok = O.solve_lower(L, B)
    assert ok
    assert np.allclose(Y, sl.solve_triangular(np.tril(L), B, lower=True), rtol=1e-11, atol=1e-13)
    W = O.chol_solve(L, B)
    assert np.allclose(K @ W, B, rtol=1e-9, atol=1e-10)
    assert np.allclose(O.chol_inverse(L) @ K, np.eye(n), atol=1e-10)


def test_scaled_optimizer_runs_and_rescales():
    """optimizer.rs:211-283: after each step params are re-read post-rescale and noise *= scale."""
    from friedrich_b200.synthetic import make_dataset
    X, y = make_dataset(11, 96, 2)
    k = O.KernelDesc.make([O.K_SQUARED_EXP], [0.7, 1.0])
    gp = O.OracleGaussianProcess(O.ConstantPrior(0.0), k, 0.1, None, X, y)
    noise0 = gp.noise
    gp.fit_parameters(False, True, max_iter=3, convergence_fraction=0.0)
    assert len(gp.trace) == 3
    assert math.isclose(gp.trace[0]["noise"], noise0 * gp.trace[0]["scale"], rel_tol=1e-15)
    assert gp.kernel.params() == gp.trace[-1]["params"]


def test_unscaled_optimizer_log_noise():
    """optimizer.rs:69-149 for a non-scalable kernel (RationalQuadratic)."""
    from friedrich_b200.synthetic import make_dataset
    X, y = make_dataset(12, 64, 2)
    k = O.KernelDesc.make([O.K_RATIONAL_QUADRATIC], [1.0, 0.8])
    gp = O.OracleGaussianProcess(O.ConstantPrior(0.0), k, 0.2, None, X, y)
    gp.fit_parameters(False, True, max_iter=2, convergence_fraction=0.0)
    assert len(gp.trace) == 2 and gp.trace[0]["scale"] == 1.0
    assert gp.noise > 0 and len(gp.trace[0]["grads"]) == 3
