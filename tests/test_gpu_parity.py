"""GPU parity tests: every entry point of the C-ABI (through the Python host mirror, which calls nothing else) against the
CPU oracle on the same seeded inputs, at sizes the oracle finishes in seconds; the golden anchors; and size-independent
properties at larger sizes.

Tolerances (BASELINE.json north star / SURVEY §8d): Cholesky factor rtol 1e-10 normwise (Frobenius), predicted
mean / variance rtol 1e-8 element-wise with an absolute floor of 1e-12; everything else that is a sum of such terms 1e-8.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

L_RTOL = 1e-10
P_RTOL = 1e-8
P_ATOL = 1e-12


def _mods():
    import friedrich_b200 as F
    from friedrich_b200 import _native as N
    from friedrich_b200.synthetic import make_dataset, make_inputs
    from oracle import oracle as O
    return F, N, O, make_dataset, make_inputs


def frob_rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def close(a, b, rtol=P_RTOL, atol=P_ATOL):
    return np.allclose(a, b, rtol=rtol, atol=atol)


def test_native_library_is_loaded_and_reports_sm100():
    F, N, O, *_ = _mods()
    assert b"sm_100a" in N.lib().fgp_version()
    h = N.Handle(0)
    h.close()


# ---------------------------------------------------------------------------------------------------------------------
# the last two shapes are large enough for the 64-row CTA shape, the others run the 32-row shape (csrc/gemm_nt.cu)
@pytest.mark.parametrize("shape", [(128, 128, 16, 0), (256, 384, 144, 0), (512, 512, 256, 1), (640, 128, 1024, 0),
                                   (2048, 2048, 80, 0), (3072, 3072, 48, 1)])
def test_gemm_nt_kernel_against_numpy(shape):
    F, N, O, *_ = _mods()
    M, Nn, K, lower = shape
    rng = np.random.default_rng(M + K)
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((Nn, K)))
    Cm = np.asfortranarray(rng.standard_normal((M, Nn)))
    ref = Cm - 0.5 * A @ B.T
    out = Cm.copy(order="F")
    assert N.lib().fgp_dbg_gemm_nt(0, N.dptr(out), M, N.dptr(A), M, N.dptr(B), Nn, M, Nn, K, -0.5, 1, lower) == 0
    if lower:
        mask = np.tril(np.ones((M, Nn), dtype=bool))
        assert np.allclose(out[mask], ref[mask], rtol=1e-12, atol=1e-12)
        assert np.array_equal(out[~mask], Cm[~mask])
    else:
        assert np.allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_gemm_nt_kernel_runs_two_ctas_per_sm():
    """Design point of gemm_nt.cu: two 64x128 CTAs resident per SM (one's epilogue overlaps the other's main loop)."""
    F, N, O, *_ = _mods()
    assert N.lib().fgp_dbg_gemm_occupancy(0) == 2
    assert N.lib().fgp_dbg_gemm_occupancy32(0) == 3   # the 32-row shape of sub-wave launches


# ---------------------------------------------------------------------------------------------------------------------
# tcgen05 trailing update (csrc/ozaki.cu): every step of it is exact except the two additions into C, so the numpy restatement
# oracle/ozaki_model.py must be reproduced BIT FOR BIT; and the result is an f64-quality product
@pytest.mark.parametrize("case", [(256, 128, 0, 0, False), (512, 512, 1, 0, True), (1024, 384, 1, 4, False), (1536, 512, 1, 2, True)])
def test_tcgen05_update_is_bit_equal_to_the_integer_model(case):
    F, N, O, *_ = _mods()
    from oracle import ozaki_model as OM
    M, K, lower, row_skip, scaled = case
    rng = np.random.default_rng(M + K)
    P = rng.standard_normal((M, K))
    if scaled:
        P *= np.exp2(rng.integers(-40, 40, size=(M, 1)).astype(np.float64))
    P[M // 2 + 3] = 0.0   # an all-zero row
    Cm = rng.standard_normal((M, M)) * 10.0
    out = np.asfortranarray(Cm)
    Pf = np.asfortranarray(P)
    assert N.lib().fgp_dbg_ozaki_syrk(0, N.dptr(out), M, N.dptr(Pf), M, M, K, lower, row_skip, 0, 0, 0) == 0
    model = OM.update(Cm, P)
    mask = OM.launch_mask(M, lower, row_skip)
    assert np.array_equal(out[mask], model[mask])
    assert np.array_equal(out[~mask], Cm[~mask])
    exact = Cm.astype(np.longdouble) - P.astype(np.longdouble) @ P.T.astype(np.longdouble)
    rowmax = np.abs(P).max(axis=1)
    budget = np.outer(rowmax, rowmax) * K * 2.0 ** -52 + 2.0 * np.abs(np.asarray(exact, dtype=np.float64)) * 2.0 ** -52
    assert np.all(np.abs(np.asarray(out - exact, dtype=np.float64))[mask] <= budget[mask])


def test_tcgen05_cta_pair_kernel_is_bit_equal_too():
    """csrc/ozaki.cu ozaki_pair_kernel (cta_group::2; an experiment kept behind fgp_dbg_ozaki_experiment(32)): same integer
    arithmetic, so the same bits — lower mode with odd column segments (lone tiles shadowed by the second CTA), row_skip, and a
    full rectangle."""
    F, N, O, *_ = _mods()
    from oracle import ozaki_model as OM
    N.lib().fgp_dbg_ozaki_experiment(32)
    try:
        for (M, K, lower, row_skip) in [(640, 256, 1, 0), (1152, 512, 1, 3), (384, 128, 0, 0)]:
            rng = np.random.default_rng(M * 3 + K)
            P = rng.standard_normal((M, K)) * np.exp2(rng.integers(-20, 20, size=(M, 1)).astype(np.float64))
            Cm = rng.standard_normal((M, M))
            out, Pf = np.asfortranarray(Cm), np.asfortranarray(P)
            assert N.lib().fgp_dbg_ozaki_syrk(0, N.dptr(out), M, N.dptr(Pf), M, M, K, lower, row_skip, 0, 0, 0) == 0
            mask = OM.launch_mask(M, lower, row_skip)
            assert np.array_equal(out[mask], OM.update(Cm, P)[mask])
            assert np.array_equal(out[~mask], Cm[~mask])
    finally:
        N.lib().fgp_dbg_ozaki_experiment(0)


def test_tcgen05_update_propagates_non_finite_rows():
    F, N, O, *_ = _mods()
    M, K = 256, 128
    rng = np.random.default_rng(5)
    P = np.asfortranarray(rng.standard_normal((M, K)))
    P[7, 3] = np.nan
    out = np.asfortranarray(rng.standard_normal((M, M)))
    assert N.lib().fgp_dbg_ozaki_syrk(0, N.dptr(out), M, N.dptr(P), M, M, K, 0, 0, 0, 0, 0) == 0
    assert np.all(np.isnan(out[7, :])) and np.all(np.isnan(out[:, 7]))
    assert np.all(np.isfinite(np.delete(np.delete(out, 7, 0), 7, 1)))


def test_fit_with_and_without_tcgen05_agree_and_meet_the_factor_tolerance():
    """n = 4096: the first six panels' trailing updates run on tcgen05 (>= 1024 rows left), A/B against the f64 DMMA kernel
    everywhere (FGP_OPT_TCGEN05 = 0); both within the 1e-10 factor tolerance of each other by a wide margin, same predictions."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 4096, 8
    X, y = make_dataset(0x5EED0040, n, d)
    Xq = make_inputs(0x5EED0041, 300, d)
    kd = F.SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
    res = {}
    for on in (1, 0):
        h = N.Handle(0)
        assert N.lib().fgp_set_option(h.ptr, N.FGP_OPT_TCGEN05, on) == 0
        h.check(N.lib().fgp_fit(h.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
        L = np.zeros((n, n), order="F")
        h.check(N.lib().fgp_download_factor(h.ptr, N.dptr(L), n))
        mean, var = np.zeros(300), np.zeros(300)
        h.check(N.lib().fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(N.fcol(Xq)), 300, 300, N.dptr(mean), N.dptr(var)))
        res[on] = (np.tril(L), mean, var)
        h.close()
    assert frob_rel(res[1][0], res[0][0]) < 1e-11   # measured 2.7e-13: two f64-accurate summation orders through a Cholesky
    assert close(res[1][1], res[0][1]) and close(res[1][2], res[0][2])


# ---------------------------------------------------------------------------------------------------------------------
def test_golden_anchors_on_gpu():
    """tests/golden/anchors.json (50-digit mpmath) through the device path."""
    F, N, O, *_ = _mods()
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))["cases"]
    tagmap = {O.K_SQUARED_EXP: F.SquaredExp, O.K_MATERN2: F.Matern2, O.K_EXPONENTIAL: F.Exponential,
              O.K_MATERN1: F.Matern1, O.K_RATIONAL_QUADRATIC: F.RationalQuadratic, O.K_LINEAR: F.Linear,
              O.K_POLYNOMIAL: F.Polynomial, O.K_HYPERTAN: F.HyperTan, O.K_MULTIQUADRIC: F.Multiquadric}

    def build_kernel(ops, params):
        st, po = [], 0
        for op in ops:
            if op == O.K_SUM:
                b, a = st.pop(), st.pop()
                st.append(a + b)
            elif op == O.K_PROD:
                b, a = st.pop(), st.pop()
                st.append(a * b)
            else:
                cls = tagmap[op]
                npar = len(cls._names)
                st.append(cls(*params[po:po + npar]))
                po += npar
        return st[0]

    for c in cases:
        gp = F.GaussianProcess(F.ZeroPrior(), build_kernel(c["ops"], c["params"]), c["noise"], None, c["X"], c["y"])
        L = gp.cholesky_factor()
        n = len(c["y"])
        assert np.all(np.isnan(L[np.triu_indices(n, 1)]))
        assert np.allclose(np.tril(L), np.array(c["L"]), rtol=1e-11, atol=1e-14), c["name"]
        Xq = np.array(c["Xq"])
        assert close(gp.predict(Xq), c["mean"]), c["name"]
        assert close(gp.predict_variance(Xq), c["var"]), c["name"]
        m, v = gp.predict_mean_variance(Xq)
        assert close(m, c["mean"]) and close(v, c["var"]), c["name"]
        assert np.allclose(gp.predict_covariance(Xq), c["cov"], rtol=1e-8, atol=1e-12), c["name"]
        assert abs(gp.likelihood() - c["likelihood"]) < 1e-9 * max(1.0, abs(c["likelihood"])), c["name"]
        assert np.allclose(gp.gradient_marginal_likelihood(), c["grad_unscaled"], rtol=1e-7, atol=1e-9), c["name"]


def test_doctest_dataset_values():
    """The reference's doc-test data (mod.rs:7-8) with the SURVEY §8c literals."""
    F, *_ = _mods()
    X, y = [[0.8], [1.2], [3.8], [4.2]], [3.0, 4.0, -2.0, -2.0]
    gp = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(1.0, 1.0), 0.27726341266023546, None, X, y)
    assert close(gp.predict([[1.0], [2.0], [3.0]]), [3.43460089928164334, 2.61347089019187292, -0.483309215643327428])
    assert close(gp.predict_variance([[1.0], [2.0], [3.0]]),
                 [0.0391981303494769655, 0.405545879974281484, 0.405545879974281298])
    assert abs(gp.likelihood() - (-13.8047231440077)) < 1e-10
    # Vec<f64> input = one sample -> scalar outputs (conversion/mod.rs:95-118)
    assert isinstance(gp.predict([1.0]), float) and abs(gp.predict([1.0]) - 3.43460089928164334) < 1e-9
    m, v = gp.predict_mean_variance([2.0])
    assert abs(m - 2.61347089019187292) < 1e-9 and abs(v - 0.405545879974281484) < 1e-9


# ---------------------------------------------------------------------------------------------------------------------
CASES = [  # n, d, q, kernel factory, oracle desc, noise
    (100, 1, 33, "sqexp"), (129, 3, 128, "matern2"), (300, 8, 100, "sqexp"), (777, 5, 200, "matern2"),
    (1024, 16, 256, "sqexp"), (1500, 32, 300, "sqexp"),
]


def _kern(F, O, name, d):
    ls = math.sqrt(d / 6.0)
    if name == "sqexp":
        return F.SquaredExp(ls, 1.0), O.KernelDesc.make([O.K_SQUARED_EXP], [ls, 1.0])
    return F.Matern2(ls, 1.0), O.KernelDesc.make([O.K_MATERN2], [ls, 1.0])


@pytest.mark.parametrize("n,d,q,kname", CASES)
def test_fit_predict_parity(n, d, q, kname):
    F, N, O, make_dataset, make_inputs = _mods()
    X, y = make_dataset(1000 + n, n, d)
    Xq = make_inputs(2000 + n, q, d)
    kern, kd = _kern(F, O, kname, d)
    gp = F.GaussianProcess(F.ConstantPrior(0.25), kern, 0.1, None, X, y)
    ref = O.OracleGaussianProcess(O.ConstantPrior(0.25), kd, 0.1, None, X, y)
    L, Lr = gp.cholesky_factor(), ref.L
    assert np.all(np.isnan(L[np.triu_indices(n, 1)]))
    assert frob_rel(np.tril(L), np.tril(Lr)) < L_RTOL
    assert close(gp.predict(Xq), ref.predict(Xq))
    assert close(gp.predict_variance(Xq), ref.predict_variance(Xq))
    m, v = gp.predict_mean_variance(Xq)
    mr, vr = ref.predict_mean_variance(Xq)
    assert close(m, mr) and close(v, vr)
    assert abs(gp.likelihood() - ref.likelihood()) < 1e-9 * abs(ref.likelihood())
    # d=1: the posterior covariance of many queries on a line is numerically singular (the reference's
    # MultivariateNormal::new would panic as well), so sample only a handful there
    qs = min(q, 64) if d > 1 else 5
    assert np.allclose(gp.predict_covariance(Xq[:qs]), ref.predict_covariance(Xq[:qs]), rtol=1e-8, atol=1e-11)
    mvn = gp.sample_at(Xq[:qs])
    mean2, cov2 = ref.sample_at_params(Xq[:qs])
    assert close(mvn.mean(), mean2)
    assert frob_rel(mvn.cholesky_covariance, np.linalg.cholesky(cov2)) < 1e-8
    s = mvn.sample(np.random.default_rng(0))
    assert s.shape == (qs,) and np.all(np.isfinite(s))


def test_single_query_and_ragged_leading_dimension():
    """q = 1 (latency case) and host matrices whose leading dimension exceeds the row count (EMatrix slack,
    extendable_matrix.rs:52-55) through the raw ABI."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 333, 4
    X, y = make_dataset(5, n, d)
    ld = n + 57
    Xpad = np.full((ld, d), np.nan, order="F")
    Xpad[:n] = X
    kern, kd = _kern(F, O, "sqexp", d)
    h = N.Handle(0)
    desc = kern.device_desc()
    h.check(N.lib().fgp_fit(h.ptr, N.dptr(Xpad), ld, n, d, N.dptr(y), C.byref(desc), 0.1, 0, 0.0))
    ref = O.OracleGaussianProcess(O.ZeroPrior(), kd, 0.1, None, X, y)
    xq = make_inputs(6, 1, d)
    out = np.zeros(1)
    var = np.zeros(1)
    h.check(N.lib().fgp_predict_mean_var(h.ptr, C.byref(desc), N.dptr(xq), 1, 1, N.dptr(out), N.dptr(var)))
    mr, vr = ref.predict_mean_variance(xq)
    assert close(out, mr) and close(var, vr)
    Lbuf = np.zeros((n + 9, n), order="F")
    h.check(N.lib().fgp_download_factor(h.ptr, N.dptr(Lbuf), n + 9))
    assert frob_rel(np.tril(Lbuf[:n]), np.tril(ref.L)) < L_RTOL
    alpha = np.zeros(n)
    h.check(N.lib().fgp_download_alpha(h.ptr, N.dptr(alpha)))
    assert np.allclose(alpha, O.chol_solve(ref.L, ref.y.reshape(-1, 1)).ravel(), rtol=1e-8, atol=1e-10)


ALL_KERNELS = [
    ("SquaredExp(0.7, 1.3)", lambda O: ([O.K_SQUARED_EXP], [0.7, 1.3])),
    ("Exponential(0.9, 1.1)", lambda O: ([O.K_EXPONENTIAL], [0.9, 1.1])),
    ("Matern1(0.8, 1.2)", lambda O: ([O.K_MATERN1], [0.8, 1.2])),
    ("Matern2(0.8, -1.2)", lambda O: ([O.K_MATERN2], [0.8, -1.2])),
    ("RationalQuadratic(1.5, 0.9)", lambda O: ([O.K_RATIONAL_QUADRATIC], [1.5, 0.9])),
    ("Linear(0.5) + SquaredExp(0.7, 1.0)", lambda O: ([O.K_LINEAR, O.K_SQUARED_EXP, O.K_SUM], [0.5, 0.7, 1.0])),
    ("Polynomial(0.5, 1.0, 2.0) * Matern2(1.0, 1.0)",
     lambda O: ([O.K_POLYNOMIAL, O.K_MATERN2, O.K_PROD], [0.5, 1.0, 2.0, 1.0, 1.0])),
    ("HyperTan(0.1, 0.2) + Matern1(0.6, 2.0)", lambda O: ([O.K_HYPERTAN, O.K_MATERN1, O.K_SUM], [0.1, 0.2, 0.6, 2.0])),
    ("Multiquadric(0.5) * SquaredExp(0.5, 1.0)",
     lambda O: ([O.K_MULTIQUADRIC, O.K_SQUARED_EXP, O.K_PROD], [0.5, 0.5, 1.0])),
]


@pytest.mark.parametrize("expr,desc", ALL_KERNELS)
def test_every_kernel_value_and_gradient(expr, desc):
    """All nine kernels + Sum/Prod (kernel.rs) through Gram -> factor -> predict -> LML gradient."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d, q = 200, 3, 40
    X, y = make_dataset(31337, n, d)
    Xq = make_inputs(4242, q, d)
    kern = eval(expr, {k: getattr(F, k) for k in dir(F)})
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.3, None, X, y)
    ref = O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make(*desc(O)), 0.3, None, X, y)
    assert frob_rel(np.tril(gp.cholesky_factor()), np.tril(ref.L)) < L_RTOL
    assert close(gp.predict(Xq), ref.predict(Xq))
    assert close(gp.predict_variance(Xq), ref.predict_variance(Xq))
    assert abs(gp.likelihood() - ref.likelihood()) < 1e-9 * abs(ref.likelihood())
    assert np.allclose(gp.gradient_marginal_likelihood(), ref.gradient_marginal_likelihood(scaled=False), rtol=1e-8,
                       atol=1e-9)
    if kern.is_scalable():
        s, g = gp.scaled_gradient_marginal_likelihood()
        sr, gr = ref.gradient_marginal_likelihood(scaled=True)
        assert abs(s - sr) < 1e-9 * abs(sr) and np.allclose(g, gr, rtol=1e-8, atol=1e-9)


def test_exponential_kernel_near_coincident_points():
    """|x-y| through the GEMM expansion loses digits for near-coincident points; the direct-difference
    fallback must keep the cusp of the Exponential kernel exact."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 256, 6
    X, y = make_dataset(9, n, d)
    X = np.asfortranarray(X + 10.0)                  # large offset: the expansion on raw coordinates would lose 4 digits
    X[1::2] = X[0::2] + 1e-7 * (X[1::2] - 10.0)      # pairs at distance ~1e-7
    gp = F.GaussianProcess(F.ZeroPrior(), F.Exponential(0.9, 1.0), 0.2, None, X, y)
    ref = O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make([O.K_EXPONENTIAL], [0.9, 1.0]), 0.2, None, X, y)
    assert frob_rel(np.tril(gp.cholesky_factor()), np.tril(ref.L)) < L_RTOL
    Xq = np.asfortranarray(X[:50] + 1e-9)
    assert close(gp.predict_variance(Xq), ref.predict_variance(Xq), rtol=1e-7, atol=1e-10)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n0,k", [(100, 5), (128, 128), (300, 77), (640, 300), (257, 1)])
def test_add_samples_matches_sequential_oracle_and_scratch_fit(n0, k):
    F, N, O, make_dataset, make_inputs = _mods()
    d = 3
    X, y = make_dataset(777, n0 + k, d)
    kern, kd = _kern(F, O, "sqexp", d)
    gp = F.GaussianProcess(F.ConstantPrior(0.1), kern, 0.1, None, X[:n0], y[:n0])
    gp.add_samples(X[n0:], y[n0:])
    ref = O.OracleGaussianProcess(O.ConstantPrior(0.1), kd, 0.1, None, X[:n0], y[:n0])
    ref.add_samples(X[n0:], y[n0:])
    L = gp.cholesky_factor()
    assert L.shape == (n0 + k, n0 + k)
    assert frob_rel(np.tril(L), np.tril(ref.L)) < L_RTOL
    Xq = make_inputs(5, 50, d)
    assert close(gp.predict(Xq), ref.predict(Xq)) and close(gp.predict_variance(Xq), ref.predict_variance(Xq))
    scratch = F.GaussianProcess(F.ConstantPrior(0.1), kern, 0.1, None, X, y)
    assert frob_rel(np.tril(L), np.tril(scratch.cholesky_factor())) < L_RTOL
    assert close(gp.predict(Xq), scratch.predict(Xq))


def test_add_samples_repeated_growth():
    """The reference's only unit test (extendable_matrix.rs:114-130): repeated add_rows must keep working while the
    capacity grows."""
    F, N, O, make_dataset, make_inputs = _mods()
    d = 2
    X, y = make_dataset(3, 700, d)
    kern, kd = _kern(F, O, "matern2", d)
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.2, None, X[:50], y[:50])
    at = 50
    for k in (1, 7, 70, 200, 372):
        gp.add_samples(X[at:at + k], y[at:at + k])
        at += k
    assert at == 700 and gp.n_samples == 700
    scratch = O.OracleGaussianProcess(O.ZeroPrior(), kd, 0.2, None, X, y)
    assert frob_rel(np.tril(gp.cholesky_factor()), np.tril(scratch.L)) < L_RTOL


# ---------------------------------------------------------------------------------------------------------------------
def test_cholesky_failure_and_epsilon_semantics():
    """algebra/mod.rs:81-91: failure reports the failing column; cholesky_epsilon substitutes the pivot."""
    F, N, O, make_dataset, make_inputs = _mods()
    X, y = make_dataset(5, 150, 4)
    Xd = np.asfortranarray(np.vstack([X, X]))
    yd = np.concatenate([y, y])
    kd = O.KernelDesc.make([O.K_SQUARED_EXP], [1.0, 1.0])
    _, fail = O.make_cholesky_cov_matrix(kd, Xd, 0.0, None)
    with pytest.raises(ArithmeticError) as ei:
        F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(1.0, 1.0), 0.0, None, Xd, yd)
    assert "set_cholesky_epsilon" in str(ei.value)
    # the first failing pivot sits among the duplicated rows; blocked and unblocked rounding may disagree on which
    # exactly-singular pivot crosses zero first, but never before the duplicates start
    assert fail - 1 >= 150 and f"column" in str(ei.value)
    gp = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(1.0, 1.0), 0.0, 1e-6, Xd, yd)
    L = gp.cholesky_factor()
    assert np.all(np.isfinite(np.tril(L)))
    Lr, fail2 = O.make_cholesky_cov_matrix(kd, Xd, 0.0, 1e-6)
    assert fail2 == 0
    assert frob_rel(np.tril(L)[:, :150], np.tril(Lr)[:, :150]) < 1e-8
    with pytest.raises(AssertionError):
        F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(1.0, 1.0), -0.1, None, X, y)  # mod.rs:150


def test_cholesky_failure_behind_tcgen05_updates_reports_the_duplicate_column():
    """The same semantics when the singular pivot sits behind several panels whose trailing updates ran on tcgen05 (n = 3200:
    the duplicated point is row 3100; rows 0..3099 are distinct, so with zero noise the first exactly-singular pivot is column
    3100), and cholesky_epsilon carries the factorisation through it."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 3200, 5
    X, y = make_dataset(0x5EED0060, n, d)
    X[3100] = X[7]
    y[3100] = y[7]
    with pytest.raises(ArithmeticError):
        F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(0.3, 1.0), 0.0, None, X, y)
    h = N.Handle(0)
    kd = F.SquaredExp(0.3, 1.0).device_desc()
    rc = N.lib().fgp_fit(h.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.0, 0, 0.0)
    assert rc == N.FGP_ERR_NOT_POSDEF and N.lib().fgp_failed_column(h.ptr) == 3100
    h.close()
    gp = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(0.3, 1.0), 0.0, 1e-6, X, y)
    assert np.all(np.isfinite(np.tril(gp.cholesky_factor())))


def test_bad_arguments_return_status_codes():
    F, N, O, make_dataset, make_inputs = _mods()
    h = N.Handle(0)
    kd = F.SquaredExp().device_desc()
    out = np.zeros(4)
    X = make_inputs(1, 4, 2)
    assert N.lib().fgp_predict_mean(h.ptr, C.byref(kd), N.dptr(X), 4, 4, N.dptr(out)) == N.FGP_ERR_NOT_FITTED
    bad = N.KernelDesc.make([F.kernels.K_SUM], [])
    assert N.lib().fgp_fit(h.ptr, N.dptr(X), 4, 4, 2, N.dptr(out), C.byref(bad), 0.1, 0, 0.0) == N.FGP_ERR_BAD_KERNEL
    assert N.lib().fgp_fit(h.ptr, N.dptr(X), 2, 4, 2, N.dptr(out), C.byref(kd), 0.1, 0, 0.0) == N.FGP_ERR_BAD_ARG
    assert N.lib().fgp_fit(h.ptr, N.dptr(X), 4, 4, 2, N.dptr(out), C.byref(kd), -1.0, 0, 0.0) == N.FGP_ERR_BAD_ARG
    assert b"noise" in N.lib().fgp_last_error(h.ptr)


# ---------------------------------------------------------------------------------------------------------------------
def test_builder_train_with_parameter_fit_follows_oracle_trajectory():
    """builder.rs:189-214 + optimizer.rs:211-283: heuristic fit, prior fit, scaled ADAM loop — the whole parameter
    trajectory against the oracle's (same number of Gram + Cholesky + gradient iterations)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 400, 3
    X, y = make_dataset(11, n, d)
    y = y + 2.0
    gp = (F.GaussianProcessBuilder(X, y).set_kernel(F.Matern2()).fit_kernel().fit_prior()
          .set_fit_parameters(6, 0.05).train())
    ref = O.OracleGaussianProcess.train(X, y, kernel=O.KernelDesc.make([O.K_MATERN2], [1.0, 1.0]), fit_kernel=True,
                                        fit_prior=True, max_iter=6, convergence_fraction=0.05)
    assert len(gp.trace) == len(ref.trace) > 0
    for a, b in zip(gp.trace, ref.trace):
        assert abs(a["scale"] - b["scale"]) < 1e-8 * abs(b["scale"])
        assert np.allclose(a["grads"], b["grads"], rtol=1e-7, atol=1e-8)
        assert np.allclose(a["params"], b["params"], rtol=1e-8)
        assert abs(a["noise"] - b["noise"]) < 1e-8 * b["noise"]
    assert abs(gp.prior.c - ref.prior.c) < 1e-12
    Xq = make_inputs(12, 64, d)
    assert close(gp.predict(Xq), ref.predict(Xq), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("shape", [(5000, 6), (1300, 17), (3000, 40)])
def test_linear_prior_fit_on_device_matches_least_squares(shape):
    """prior.rs:139-159 through fgp_linear_prior_fit (normal equations on the resident centred inputs) and through
    fit_parameters(fit_prior=True) of a LinearPrior model, against the oracle's SVD least squares."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = shape
    X, y = make_dataset(0x5EED0050 + d, n, d)
    rng = np.random.default_rng(d)
    y = y + X @ rng.standard_normal(d) * 3.0 + 1.7
    ref = O.LinearPrior.default(d)
    ref.fit(X, y)
    gp = F.GaussianProcess(F.LinearPrior.default(d), F.SquaredExp(math.sqrt(d / 6.0), 1.0), 0.1, None, X, y)
    w, b = np.zeros(d), C.c_double(0.0)
    gp._h.check(N.lib().fgp_linear_prior_fit(gp._h.ptr, N.dptr(y), N.dptr(w), C.cast(C.byref(b), N._dp)))
    assert np.allclose(w, ref.weights, rtol=1e-9, atol=1e-11) and abs(b.value - ref.intercept) < 1e-9 * max(1.0, abs(ref.intercept))
    gp.fit_parameters(True, False)
    assert np.allclose(gp.prior.weights, ref.weights, rtol=1e-9, atol=1e-11)
    assert abs(gp.prior.intercept - ref.intercept) < 1e-9 * max(1.0, abs(ref.intercept))
    if n <= 1500:
        oref = O.OracleGaussianProcess(ref, O.KernelDesc.make([O.K_SQUARED_EXP], [math.sqrt(d / 6.0), 1.0]), 0.1, None, X, y)
        Xq = make_inputs(99, 50, d)
        assert close(gp.predict(Xq), oref.predict(Xq), rtol=1e-7, atol=1e-9)


def test_unscaled_optimizer_trajectory():
    """optimizer.rs:69-149 (non-scalable kernel: noise fitted in log space)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 300, 2
    X, y = make_dataset(13, n, d)
    kern = F.RationalQuadratic(1.5, 0.9)
    gp = F.GaussianProcessBuilder(X, y).set_kernel(kern).set_prior(F.ZeroPrior()).fit_kernel().set_fit_parameters(4, 0.05).train()
    ref = O.OracleGaussianProcess.train(X, y, kernel=O.KernelDesc.make([O.K_RATIONAL_QUADRATIC], [1.5, 0.9]),
                                        prior=O.ZeroPrior(), fit_kernel=True, max_iter=4)
    assert len(gp.trace) == len(ref.trace) > 0
    for a, b in zip(gp.trace, ref.trace):
        assert np.allclose(a["grads"], b["grads"], rtol=1e-7, atol=1e-8)
        assert np.allclose(a["params"], b["params"], rtol=1e-8)
        assert abs(a["noise"] - b["noise"]) < 1e-8 * b["noise"]


def test_mean_pair_distance_heuristic():
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 900, 7
    X, _ = make_dataset(21, n, d)
    h = N.Handle(0)
    h.check(N.lib().fgp_set_inputs(h.ptr, N.dptr(X), n, n, d))
    out = C.c_double()
    h.check(N.lib().fgp_mean_pair_distance(h.ptr, C.cast(C.byref(out), N._dp)))
    assert abs(out.value - O.fit_bandwidth_mean(X)) < 1e-11 * out.value


# ---------------------------------------------------------------------------------------------------------------------
def test_config2_full_size_parity():
    """BASELINE.json configs[1]: RBF n=4096 d=8 fit + predict 1024 queries, against the oracle (~10 s of CPU)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d, q = 4096, 8, 1024
    X, y = make_dataset(0x5EED0002, n, d)
    Xq = make_inputs(0x5EED0003, q, d)
    kern, kd = _kern(F, O, "sqexp", d)
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y)
    ref = O.OracleGaussianProcess(O.ZeroPrior(), kd, 0.1, None, X, y)
    assert frob_rel(np.tril(gp.cholesky_factor()), np.tril(ref.L)) < L_RTOL
    m, v = gp.predict_mean_variance(Xq[:256])
    mr, vr = ref.predict_mean_variance(Xq[:256])
    assert close(m, mr) and close(v, vr)
    assert close(gp.predict_variance(Xq[:256]), ref.predict_variance(Xq[:256]))


def test_full_size_properties_n16384():
    """Size-independent properties at the metric's size (n=16384, d=16), where the oracle would take ~20 minutes:
    residual ||L L^T x - K x|| on random probes, predicted variance at training points (0 <= var <= noise^2-ish and the
    posterior mean reproduces y within the noise), prediction consistency between the three predict entry points, and
    idempotence of refit."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 16384, 16
    X, y = make_dataset(0x5EED0003, n, d)
    kern, kd = _kern(F, O, "sqexp", d)
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y)
    L = np.tril(gp.cholesky_factor())
    rng = np.random.default_rng(0)
    idx = rng.choice(n, 48, replace=False)
    # rows of K recomputed by the oracle's kernel function on the host (48 x n evaluations)
    Krows = O.make_covariance_matrix(kd, X[idx], X)
    Krows[np.arange(48), idx] += 0.1 ** 2
    rec = L[idx] @ L.T
    assert np.abs(rec - Krows).max() < 1e-12 * n
    Xq = make_inputs(77, 512, d)
    m1 = gp.predict(Xq)
    v1 = gp.predict_variance(Xq)
    m2, v2 = gp.predict_mean_variance(Xq)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
    assert np.all(v1 > 0) and np.all(v1 <= 1.0 + 1e-12)
    # variance at training points is below the noise floor bound noise^2 (k(x,x) - k^T (K+s I)^-1 k <= s)
    vt = gp.predict_variance(X[:256])
    assert np.all(vt >= -1e-10) and np.all(vt <= 0.1 ** 2 + 1e-10)
    # alpha solves K alpha = y: check 48 rows
    alpha = np.zeros(n)
    gp._h.check(N.lib().fgp_download_alpha(gp._h.ptr, N.dptr(alpha)))
    assert np.abs(Krows @ alpha - y[idx]).max() < 1e-9
    gp._refit()
    assert np.array_equal(np.tril(gp.cholesky_factor()), L)  # deterministic, idempotent


def test_wavefront_solves_beyond_one_resident_wave_n24576():
    """192 block rows > 148 SMs: the flag-synchronised triangular solves (csrc/vector_kernels.cuh) must still terminate and
    give alpha = K^-1 y when not every block of the grid is resident at once (blocks only wait for lower block indices)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 24576, 8
    X, y = make_dataset(0x5EED0009, n, d)
    kern, kd = _kern(F, O, "sqexp", d)
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y)
    alpha = np.zeros(n)
    gp._h.check(N.lib().fgp_download_alpha(gp._h.ptr, N.dptr(alpha)))
    idx = np.random.default_rng(3).choice(n, 32, replace=False)
    Krows = O.make_covariance_matrix(kd, X[idx], X)
    Krows[np.arange(32), idx] += 0.1 ** 2
    assert np.abs(Krows @ alpha - y[idx]).max() < 1e-9
    m = gp.predict(X[idx])                       # posterior mean at training points = (K - s^2 I) alpha
    assert close(m, y[idx] - 0.1 ** 2 * alpha[idx])


def test_two_devices_in_one_process():
    """Per-device one-time kernel setup (opt-in shared memory sizes): a second model on another GPU of the same process
    must work like the first.  Skipped on single-GPU boxes."""
    F, N, O, make_dataset, make_inputs = _mods()
    try:
        h1 = N.Handle(1)
    except Exception:
        pytest.skip("needs a second GPU")
    h1.close()
    n, d = 900, 4
    X, y = make_dataset(0x5EED0010, n, d)
    Xq = make_inputs(0x5EED0011, 40, d)
    kern, kd = _kern(F, O, "sqexp", d)
    g0 = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y, device=0)
    g1 = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y, device=1)
    m0, v0 = g0.predict_mean_variance(Xq)
    m1, v1 = g1.predict_mean_variance(Xq)
    assert np.array_equal(np.tril(g0.cholesky_factor()), np.tril(g1.cholesky_factor()))
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2] and configs[4] at (or near) config size, and the state round trip
def test_config5_add_samples_16384_plus_1024():
    """configs[4]: fit n=16384 d=16, add_samples k=1024 (algebra/mod.rs:97-126).  The oracle's sequential append would take
    hours, so the check is the size-independent identity the block update must satisfy: the updated factor equals the
    from-scratch device factor of all 17408 rows (new block rows to 1e-10 Frobenius, old block bit-identical because it is
    not touched), and both models predict the same.  Also exercises the capacity regrowth copy of the 2.1 GB factor."""
    F, N, O, make_dataset, make_inputs = _mods()
    n0, k, d = 16384, 1024, 16
    X, y = make_dataset(0x5EED0005, n0 + k, d)
    kern, kd = _kern(F, O, "sqexp", d)
    gp = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X[:n0], y[:n0])
    L0 = np.tril(gp.cholesky_factor())
    gp.add_samples(X[n0:], y[n0:])
    assert gp.n_samples == n0 + k
    L = np.tril(gp.cholesky_factor())
    assert np.array_equal(L[:n0, :n0], L0)                     # insert_column(end) never changes the old factor
    del L0
    scratch = F.GaussianProcess(F.ZeroPrior(), kern, 0.1, None, X, y)
    Ls = np.tril(scratch.cholesky_factor())
    assert frob_rel(L[n0:], Ls[n0:]) < L_RTOL                  # the 1024 new block rows
    assert frob_rel(L[:n0, :n0], Ls[:n0, :n0]) < L_RTOL
    # rows of K on the host (oracle kernel function) against L L^T on probe rows among the NEW samples
    idx = n0 + np.random.default_rng(5).choice(k, 16, replace=False)
    Krows = O.make_covariance_matrix(kd, X[idx], X)
    Krows[np.arange(16), idx] += 0.1 ** 2
    assert np.abs(L[idx] @ L.T - Krows).max() < 1e-12 * (n0 + k)
    del L, Ls
    Xq = make_inputs(0x5EED0006, 256, d)
    m1, v1 = gp.predict_mean_variance(Xq)
    m2, v2 = scratch.predict_mean_variance(Xq)
    assert close(m1, m2) and close(v1, v2)


def test_config3_matern_scaled_adam_against_oracle_n2048():
    """configs[2] at a size the oracle still finishes (n=2048, d=16, Matern-5/2): three iterations of the scaled ADAM loop
    (optimizer.rs:211-283) — scale, gradients, parameters, noise per iteration.  n=2048 = 16 tile rows runs the K^-1 = U U^T
    GEMM in its 64-row / multi-wave form and the shrinking-row TRSM over 16 block columns (n=400 elsewhere is 4 tiles)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 2048, 16
    X, y = make_dataset(0x5EED0003, n, d)
    ls = math.sqrt(d / 6.0)
    gp = F.GaussianProcess(F.ZeroPrior(), F.Matern2(ls, 1.0), 0.1, None, X, y)
    gp.fit_parameters(False, True, 3, 0.0)
    ref = O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make([O.K_MATERN2], [ls, 1.0]), 0.1, None, X, y)
    ref.fit_parameters(False, True, 3, 0.0)
    assert len(gp.trace) == len(ref.trace) == 3
    for a, b in zip(gp.trace, ref.trace):
        assert abs(a["scale"] - b["scale"]) < 1e-8 * abs(b["scale"])
        assert np.allclose(a["grads"], b["grads"], rtol=1e-7, atol=1e-8)
        assert np.allclose(a["params"], b["params"], rtol=1e-8)
        assert abs(a["noise"] - b["noise"]) < 1e-8 * b["noise"]
    assert frob_rel(np.tril(gp.cholesky_factor()), np.tril(ref.L)) < L_RTOL


def test_config3_inverse_identity_n16384():
    """configs[2] at full size (Matern-5/2, n=16384, d=16): the explicit inverse behind the LML gradient
    (`covmat_cholesky.inverse()`, optimizer.rs:169 -> U = L^-T by the shrinking-row TRSM, K^-1 = U U^T by the k_from_tile
    GEMM) checked by identities the oracle cannot afford here: K[probe rows] K^-1[:, probe cols] = I, K^-1[:, cols]^T y =
    alpha[cols] (alpha comes from the independent wavefront solves), and the scaled gradient's scale = y.alpha / n."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 16384, 16
    X, y = make_dataset(0x5EED0003, n, d)
    ls = math.sqrt(d / 6.0)
    kd = O.KernelDesc.make([O.K_MATERN2], [ls, 1.0])
    gp = F.GaussianProcess(F.ZeroPrior(), F.Matern2(ls, 1.0), 0.1, None, X, y)
    scale, grads = gp.scaled_gradient_marginal_likelihood()
    assert np.all(np.isfinite(grads))
    rng = np.random.default_rng(16384)
    cols = np.sort(rng.choice(n, 32, replace=False))
    cols[0], cols[-1] = 0, n - 1                               # first and last block column included
    Kinv = gp.inverse_columns(cols)
    Krows = O.make_covariance_matrix(kd, X[cols], X)
    Krows[np.arange(32), cols] += 0.1 ** 2
    assert np.abs(Krows @ Kinv - np.eye(32)).max() < 1e-9
    alpha = np.zeros(n)
    gp._h.check(N.lib().fgp_download_alpha(gp._h.ptr, N.dptr(alpha)))
    assert np.allclose(Kinv.T @ y, alpha[cols], rtol=1e-8, atol=1e-10)
    assert abs(scale - float(y @ alpha) / n) < 1e-10 * abs(scale)
    # the gradient's trace term against the same columns: tr(K^-1 G_p) restricted to the probe columns is not available
    # without G_p, but K^-1's diagonal must be positive and below 1/noise^2
    dg = Kinv[cols, np.arange(32)]
    assert np.all(dg > 0) and np.all(dg < 1.0 / 0.1 ** 2 + 1e-9)


def test_state_round_trip_restores_model_without_refit():
    """serde round trip (mod.rs:58): download -> destroy -> upload -> identical predictions, factor and likelihood."""
    import pickle
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 1000, 5
    X, y = make_dataset(0x5EED0020, n, d)
    Xq = make_inputs(0x5EED0021, 300, d)
    kern, kd = _kern(F, O, "matern2", d)
    gp = F.GaussianProcess(F.ConstantPrior(0.3), kern, 0.1, None, X, y)
    m0, v0 = gp.predict_mean_variance(Xq)
    m1 = gp.predict(Xq[:3])                                    # latency path (q <= 16)
    lik0, L0 = gp.likelihood(), gp.cholesky_factor()
    blob = pickle.dumps(gp.to_state())
    gp._h.close()
    del gp
    gp2 = F.GaussianProcess.from_state(pickle.loads(blob))
    assert np.array_equal(np.tril(gp2.cholesky_factor()), np.tril(L0))
    m, v = gp2.predict_mean_variance(Xq)
    assert np.allclose(m, m0, rtol=1e-12, atol=1e-13) and np.allclose(v, v0, rtol=1e-11, atol=1e-13)
    assert np.allclose(gp2.predict(Xq[:3]), m1, rtol=1e-12, atol=1e-13)
    assert abs(gp2.likelihood() - lik0) < 1e-10 * abs(lik0)
    # the restored model keeps working as a model: add_samples and the optimiser's gradient
    Xn, yn = make_dataset(0x5EED0022, 130, d)
    gp2.add_samples(Xn, yn)
    ref = O.OracleGaussianProcess(O.ConstantPrior(0.3), kd, 0.1, None, X, y)
    ref.add_samples(Xn, yn)
    assert frob_rel(np.tril(gp2.cholesky_factor()), np.tril(ref.L)) < L_RTOL
    s, g = gp2.scaled_gradient_marginal_likelihood()
    sr, gr = ref.gradient_marginal_likelihood(scaled=True)
    assert abs(s - sr) < 1e-9 * abs(sr) and np.allclose(g, gr, rtol=1e-8, atol=1e-9)


def test_failing_column_is_exact_when_the_defect_is_structural():
    """nalgebra reports the FIRST column whose pivot is not positive.  With a leading block that is safely positive definite
    and one diagonal entry pushed far negative the failing column is unambiguous for any summation order: the device
    factorisation must report exactly that column (MultivariateNormal::new path, multivariate_normal.rs:54-59)."""
    F, N, O, make_dataset, make_inputs = _mods()
    n = 640
    rng = np.random.default_rng(12)
    M = rng.standard_normal((n, n + 8))
    A0 = np.asfortranarray(M @ M.T / n + 0.5 * np.eye(n))
    for j0 in (0, 37, 128, 300, 639):
        A = A0.copy(order="F")
        A[j0, j0] = -3.0
        ref = A.copy(order="F")
        assert O.cholesky_inplace(ref) == j0 + 1
        failed = C.c_int64(-7)
        rc = N.lib().fgp_cholesky_lower(0, N.dptr(A), n, n, C.byref(failed))
        assert rc == N.FGP_ERR_NOT_POSDEF and failed.value == j0, (j0, failed.value)
    A = A0.copy(order="F")
    failed = C.c_int64(-7)
    assert N.lib().fgp_cholesky_lower(0, N.dptr(A), n, n, C.byref(failed)) == 0 and failed.value == -1
    ref = A0.copy(order="F")
    assert O.cholesky_inplace(ref) == 0
    assert frob_rel(np.tril(A), np.tril(ref)) < L_RTOL and np.all(A[np.triu_indices(n, 1)] == 0.0)


def test_add_samples_failure_leaves_a_refittable_handle():
    """a failed block update must not leave the handle with the new sample count."""
    F, N, O, make_dataset, make_inputs = _mods()
    n0, d = 200, 3
    X, y = make_dataset(91, n0, d)
    gp = F.GaussianProcess(F.ZeroPrior(), F.Exponential(0.9, 1.0), 0.0, None, X, y)   # noise 0: duplicates are singular
    m0 = gp.predict(X[:5])
    with pytest.raises(ArithmeticError):
        gp.add_samples(np.vstack([X[:40], X[:40]]), np.concatenate([y[:40], y[:40]]))
    assert gp.n_samples == n0
    gp._refit()
    assert np.allclose(gp.predict(X[:5]), m0, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("q", [2, 7, 12, 13, 16, 17])
def test_small_query_batches_use_one_wavefront_solve(q):
    """q <= 16 takes the latency path (csrc/fgp_api.cu predict_small): ONE multi-right-hand-side wavefront launch; q = 17 is
    the first batch on the tensor-pipe path.  Same results either way, against the oracle."""
    F, N, O, make_dataset, make_inputs = _mods()
    n, d = 1500, 6
    X, y = make_dataset(0x5EED0030, n, d)
    Xq = make_inputs(0x5EED0031 + q, q, d)
    kern, kd = _kern(F, O, "matern2", d)
    gp = F.GaussianProcess(F.ConstantPrior(0.5), kern, 0.1, None, X, y)
    ref = O.OracleGaussianProcess(O.ConstantPrior(0.5), kd, 0.1, None, X, y)
    m, v = gp.predict_mean_variance(Xq)
    mr, vr = ref.predict_mean_variance(Xq)
    assert close(m, mr) and close(v, vr)
    assert close(gp.predict_variance(Xq), ref.predict_variance(Xq))
    assert gp._h.last_launch_count() <= 12 or q > 16   # one wavefront launch, not one per query
