"""GPU tests of the sharded (multi-GPU) fit entry points on ONE GPU: a communicator of size 1 runs the very same
schedule (owned-panel Gram assembly, pack, look-ahead update from the packed panel buffer, unpack) without NCCL traffic.
The 2/4/8-GPU runs are checked by tools/sharded_check.py (under torchrun; results in profiles/)."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mods():
    from friedrich_b200 import _native as N
    from friedrich_b200 import sharded
    from friedrich_b200.kernels import Matern2, SquaredExp
    from friedrich_b200.synthetic import make_dataset
    from oracle import oracle as O
    return N, sharded, SquaredExp, Matern2, make_dataset, O


def _factor(N, h, n):
    L = np.zeros((n, n), order="F")
    h.check(N.lib().fgp_download_factor(h.ptr, N.dptr(L), n))
    return L


@pytest.mark.parametrize("n,d", [(100, 2), (640, 3), (1500, 8), (2049, 5)])
def test_single_rank_sharded_fit_matches_oracle_and_plain_fit(n, d):
    N, sharded, SquaredExp, Matern2, make_dataset, O = _mods()
    X, y = make_dataset(4000 + n, n, d)
    ls = math.sqrt(d / 6.0)
    kd = SquaredExp(ls, 1.0).device_desc()
    hs, hp = N.Handle(0), N.Handle(0)
    sharded.comm_init(hs, 0, 1)
    sharded.fit_sharded(hs, X, y, kd, 0.1)
    hp.check(N.lib().fgp_fit(hp.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
    Ls, Lp = _factor(N, hs, n), _factor(N, hp, n)
    # every tile sees the same operands in the same order whichever schedule applies the update => bitwise equal
    assert np.array_equal(np.tril(Ls), np.tril(Lp))
    ref = O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make([O.K_SQUARED_EXP], [ls, 1.0]), 0.1, None, X, y)
    err = np.linalg.norm(np.tril(Ls) - np.tril(ref.L)) / np.linalg.norm(np.tril(ref.L))
    assert err < 1e-10
    a_s, a_p = np.zeros(n), np.zeros(n)
    hs.check(N.lib().fgp_download_alpha(hs.ptr, N.dptr(a_s)))
    hp.check(N.lib().fgp_download_alpha(hp.ptr, N.dptr(a_p)))
    assert np.array_equal(a_s, a_p)
    # refit with another kernel through the collective entry point
    kd2 = Matern2(ls, 1.3).device_desc()
    sharded.refit_sharded(hs, kd2, 0.2)
    hp.check(N.lib().fgp_refit(hp.ptr, C.byref(kd2), 0.2, 0, 0.0))
    assert np.array_equal(np.tril(_factor(N, hs, n)), np.tril(_factor(N, hp, n)))


@pytest.mark.parametrize("n,pipe", [(4608, 1), (10600, 1), (4608, 0)])
def test_single_rank_sharded_fit_on_the_tcgen05_path_is_bitwise_equal_and_predicts_alike(n, pipe):
    """n = 4608: the first seven panels have >= 1024 rows below them, so the sharded schedule's look-ahead update of the next
    panel's column block and its grouped trailing update run on tcgen05 (csrc/ozaki.cu) — same per-tile arithmetic as the
    single-GPU schedule, hence the same bits; the sharded fit also keeps the digit slices and every W_p, so predict takes the
    tcgen05 panel solve afterwards and must reproduce the plain model's predictions bit for bit.  n = 10600 (np = 10624: a
    narrower last panel) makes the early panels travel in FOUR row pieces (512, 512, 2 x <= 8192: csrc/sharded.cu
    factor_sharded_pipe) and ends with one-piece panels: every transition of the row-piece schedule on one rank.  pipe = 0:
    the one-piece schedule (factor_sharded_head, FGP_OPT_SHARD_PIPE = 0) must give the same bits."""
    N, sharded, SquaredExp, Matern2, make_dataset, O = _mods()
    from friedrich_b200.synthetic import make_inputs
    d, q = 6, 300
    X, y = make_dataset(4000 + n, n, d)
    Xq = make_inputs(4001 + n, q, d)
    kd = SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
    hs, hp = N.Handle(0), N.Handle(0)
    sharded.comm_init(hs, 0, 1)
    assert N.lib().fgp_set_option(hs.ptr, N.FGP_OPT_SHARD_PIPE, pipe) == 0
    sharded.fit_sharded(hs, X, y, kd, 0.1)
    hp.check(N.lib().fgp_fit(hp.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
    assert np.array_equal(np.tril(_factor(N, hs, n)), np.tril(_factor(N, hp, n)))
    out = {}
    for name, h in (("sharded", hs), ("plain", hp)):
        mean, var = np.zeros(q), np.zeros(q)
        h.check(N.lib().fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(N.fcol(Xq)), q, q, N.dptr(mean), N.dptr(var)))
        out[name] = (mean, var)
    assert np.array_equal(out["sharded"][0], out["plain"][0]) and np.array_equal(out["sharded"][1], out["plain"][1])
    hs.close()
    hp.close()


def test_sharded_fit_reports_failing_column():
    N, sharded, SquaredExp, Matern2, make_dataset, O = _mods()
    n, d = 700, 2
    X, y = make_dataset(77, n, d)
    X[650] = X[3]  # duplicate point + zero noise => singular covariance
    kd = SquaredExp(0.5, 1.0).device_desc()
    h = N.Handle(0)
    sharded.comm_init(h, 0, 1)
    with pytest.raises(N.NotPositiveDefinite):
        sharded.fit_sharded(h, X, y, kd, 0.0)
    assert 0 <= N.lib().fgp_failed_column(h.ptr) <= 650


def test_sharded_fit_without_communicator_is_an_error():
    N, sharded, SquaredExp, Matern2, make_dataset, O = _mods()
    X, y = make_dataset(5, 64, 2)
    h = N.Handle(0)
    with pytest.raises(N.FgpError) as e:
        sharded.fit_sharded(h, X, y, SquaredExp(1.0, 1.0).device_desc(), 0.1)
    assert e.value.code == N.FGP_ERR_COMM


@pytest.mark.parametrize("n,d,kname", [(300, 3, "sqexp"), (1100, 4, "matern2"), (2049, 5, "sqexp")])
def test_single_rank_sharded_lml_gradient_matches_plain_and_oracle(n, d, kname):
    """fgp_lml_gradient_sharded on a communicator of one rank runs the sharded schedule itself (cyclic rows of L^-T, gather
    layout, K^-1 on owned panels with k_tile0, per-panel reductions) — against the single-GPU entry point and the oracle."""
    N, sharded, SquaredExp, Matern2, make_dataset, O = _mods()
    X, y = make_dataset(6000 + n, n, d)
    ls = math.sqrt(d / 6.0)
    kern = SquaredExp(ls, 1.0) if kname == "sqexp" else Matern2(ls, 1.0)
    okd = O.KernelDesc.make([O.K_SQUARED_EXP if kname == "sqexp" else O.K_MATERN2], [ls, 1.0])
    kd = kern.device_desc()
    hs, hp = N.Handle(0), N.Handle(0)
    sharded.comm_init(hs, 0, 1)
    sharded.fit_sharded(hs, X, y, kd, 0.1)
    hp.check(N.lib().fgp_fit(hp.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
    scale_s, g_s = sharded.lml_gradient_sharded(hs, kd, 0.1, 2, scaled=True)
    g_p, scale_p = np.zeros(3), C.c_double(1.0)
    hp.check(N.lib().fgp_lml_gradient(hp.ptr, C.byref(kd), 0.1, 1, C.cast(C.byref(scale_p), N._dp), N.dptr(g_p)))
    assert abs(scale_s - scale_p.value) <= 1e-12 * abs(scale_p.value)
    assert np.allclose(g_s, g_p[:2], rtol=1e-10, atol=1e-12)
    _, gu_s = sharded.lml_gradient_sharded(hs, kd, 0.1, 2, scaled=False)
    gu_p = np.zeros(3)
    hp.check(N.lib().fgp_lml_gradient(hp.ptr, C.byref(kd), 0.1, 0, None, N.dptr(gu_p)))
    assert np.allclose(gu_s, gu_p, rtol=1e-10, atol=1e-12)
    if n <= 1200:
        ref = O.OracleGaussianProcess(O.ZeroPrior(), okd, 0.1, None, X, y)
        sr, gr = ref.gradient_marginal_likelihood(scaled=True)
        assert abs(scale_s - sr) < 1e-9 * abs(sr) and np.allclose(g_s, gr, rtol=1e-8, atol=1e-9)
