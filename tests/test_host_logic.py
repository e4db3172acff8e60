"""CPU: the host-side mirror of the reference API that needs no device — kernel algebra, parameter plumbing, rescale
rules, builder defaults (src/parameters/kernel.rs, src/gaussian_process/builder.rs)."""
import math

import numpy as np
import pytest

import friedrich_b200 as F
from friedrich_b200 import kernels as K
from oracle import oracle as O


def test_device_desc_is_postfix_program():
    k = (F.Linear(0.5) + F.SquaredExp(0.7, 1.3)) * F.Matern2(2.0, 3.0)
    d = k.device_desc()
    assert [d.op[i] for i in range(d.n_ops)] == [K.K_LINEAR, K.K_SQUARED_EXP, K.K_SUM, K.K_MATERN2, K.K_PROD]
    assert [d.param[i] for i in range(5)] == [0.5, 0.7, 1.3, 2.0, 3.0]
    assert k.nb_parameters() == 5 and k.get_parameters() == [0.5, 0.7, 1.3, 2.0, 3.0]


def test_desc_evaluates_like_oracle_kernel():
    """The descriptor the host emits is the byte-identical input of the oracle's kernel evaluation."""
    k = F.Polynomial(0.5, 1.0, 2.0) * F.Matern2(1.0, 1.5) + F.Exponential(0.8, 0.9)
    d = k.device_desc()
    od = O.KernelDesc.make([d.op[i] for i in range(d.n_ops)], [d.param[i] for i in range(k.nb_parameters())])
    x1, x2 = np.array([0.1, 0.2, 0.3]), np.array([0.5, -0.2, 0.9])
    r2 = float(((x1 - x2) ** 2).sum())
    r = math.sqrt(r2)
    dot = float(x1 @ x2)
    x = math.sqrt(5.0) * r / 1.0
    expect = (0.5 * dot + 1.0) ** 2.0 * (1.5 * (1 + x + 5 * r2 / 3.0) * math.exp(-x)) + 0.9 * math.exp(-r / (2 * 0.8 ** 2))
    assert O.kernel_value(od, x1, x2) == pytest.approx(expect, rel=1e-14)


def test_set_parameters_and_rescale_rules():
    k = F.SquaredExp(1.0, 2.0) + F.Matern1(3.0, 4.0)
    assert k.is_scalable()
    k.set_parameters([1.5, 2.5, 3.5, 4.5])
    assert k.get_parameters() == [1.5, 2.5, 3.5, 4.5]
    k.rescale(2.0)  # Sum rescales both (kernel.rs:174-178)
    assert k.get_parameters() == [1.5, 5.0, 3.5, 9.0]
    p = F.Linear(1.0) * F.SquaredExp(1.0, 2.0)
    assert p.is_scalable()  # Prod: OR (kernel.rs:239-242)
    p.rescale(3.0)          # Prod rescales the first scalable factor only (kernel.rs:264-274)
    assert p.get_parameters() == [1.0, 1.0, 6.0]
    assert not (F.Linear(1.0) + F.SquaredExp(1.0, 1.0)).is_scalable()  # Sum: AND (kernel.rs:150-153)
    for d in (F.Linear(), F.Polynomial(), F.HyperTan(), F.Multiquadric(), F.RationalQuadratic()):
        assert not d.is_scalable()


def test_defaults_match_reference():
    assert F.SquaredExp().get_parameters() == [1.0, 1.0]          # kernel.rs:531-534
    assert F.Polynomial().get_parameters() == [1.0, 0.0, 1.0]     # kernel.rs:438-441
    assert F.Linear().get_parameters() == [0.0]
    assert F.Gaussian is F.SquaredExp                             # kernel.rs:496
    b = F.GaussianProcessBuilder([[0.8], [1.2], [3.8], [4.2]], [3.0, 4.0, -2.0, -2.0])
    assert b.noise == 0.27726341266023546                         # builder.rs:73, SURVEY §8c anchor
    assert isinstance(b.kernel, F.SquaredExp) and isinstance(b.prior, F.ConstantPrior)
    assert (b.max_iter, b.convergence_fraction, b.cholesky_epsilon) == (100, 0.05, None)
    with pytest.raises(AssertionError):
        b.set_noise(-1.0)                                         # builder.rs:123


def test_priors():
    X = np.array([[0.0, 1.0], [1.0, 0.0], [2.0, 2.0], [3.0, 1.0]])
    y = 2.0 * X[:, 0] - X[:, 1] + 0.5
    c = F.ConstantPrior.default(2)
    c.fit(X, y)
    assert c.c == pytest.approx(y.mean()) and np.all(c.prior(X) == c.c)
    lp = F.LinearPrior.default(2)
    lp.fit(X, y)
    assert np.allclose(lp.prior(X), y) and lp.intercept == pytest.approx(0.5)
    assert np.all(F.ZeroPrior.default(2).prior(X) == 0.0)


def test_heuristic_fit_plumbing():
    k = F.Linear(1.0) + F.SquaredExp(1.0, 1.0)
    calls = []
    k.heuristic_fit(lambda: calls.append("bw") or 0.3, lambda: calls.append("amp") or 4.0)
    assert k.get_parameters() == [1.0, 0.3, 4.0] and calls == ["bw", "amp"]
