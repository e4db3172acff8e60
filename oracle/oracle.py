"""ctypes front-end of the CPU oracle (oracle/friedrich_oracle.c).  TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED *** — see the header of friedrich_oracle.c: the Rust reference cannot run in this image, so
the oracle is a restatement pinned against independent known-answer anchors (tests/golden/) and LAPACK.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

`OracleGaussianProcess` mirrors the reference's public flow (src/gaussian_process/{mod,builder,optimizer}.rs) on
top of the C restatement; priors (src/parameters/prior.rs) are O(n d) host work and live here in numpy.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfriedrich_oracle.so")

MAX_OPS = 15
MAX_PARAMS = 24

K_LINEAR, K_POLYNOMIAL, K_SQUARED_EXP, K_EXPONENTIAL, K_MATERN1, K_MATERN2 = 1, 2, 3, 4, 5, 6
K_HYPERTAN, K_MULTIQUADRIC, K_RATIONAL_QUADRATIC, K_SUM, K_PROD = 7, 8, 9, 100, 101
_NPARAMS = {K_LINEAR: 1, K_POLYNOMIAL: 3, K_SQUARED_EXP: 2, K_EXPONENTIAL: 2, K_MATERN1: 2, K_MATERN2: 2,
            K_HYPERTAN: 2, K_MULTIQUADRIC: 1, K_RATIONAL_QUADRATIC: 2, K_SUM: 0, K_PROD: 0}


class KernelDesc(C.Structure):
    """Binary twin of `fgp_kernel_desc` (include/fgp_kernel_desc.h)."""
    _fields_ = [("n_ops", C.c_int32), ("op", C.c_int32 * MAX_OPS), ("param", C.c_double * MAX_PARAMS)]

    @classmethod
    def make(cls, ops, params):
        k = cls()
        k.n_ops = len(ops)
        for i, o in enumerate(ops):
            k.op[i] = o
        for i, p in enumerate(params):
            k.param[i] = float(p)
        return k

    def nparams(self):
        return sum(_NPARAMS[self.op[i]] for i in range(self.n_ops))

    def params(self):
        return [self.param[i] for i in range(self.nparams())]

    def copy(self):
        return KernelDesc.make([self.op[i] for i in range(self.n_ops)], self.params())


class OptRecord(C.Structure):
    _fields_ = [("scale", C.c_double), ("grads", C.c_double * (MAX_PARAMS + 1)),
                ("params", C.c_double * MAX_PARAMS), ("noise", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "friedrich_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_i64 = C.c_int64


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        kd = C.POINTER(KernelDesc)
        L.fo_kernel.restype = C.c_double
        L.fo_kernel.argtypes = [kd, _dp, _i64, _dp, _i64, _i64]
        L.fo_kernel_gradient.restype = C.c_int
        L.fo_kernel_gradient.argtypes = [kd, _dp, _i64, _dp, _i64, _i64, _dp]
        L.fo_desc_nb_parameters.argtypes = [kd]
        L.fo_desc_is_scalable.argtypes = [kd]
        L.fo_desc_rescale.argtypes = [kd, C.c_double]
        L.fo_make_covariance_matrix.restype = None
        L.fo_make_covariance_matrix.argtypes = [kd, _dp, _i64, _i64, _dp, _i64, _i64, _i64, _dp, _i64]
        L.fo_gram_lower.restype = None
        L.fo_gram_lower.argtypes = [kd, _dp, _i64, _i64, _i64, C.c_double, _dp, _i64]
        L.fo_cholesky_inplace.restype = _i64
        L.fo_cholesky_inplace.argtypes = [_dp, _i64, _i64, C.c_int, C.c_double]
        L.fo_make_cholesky_cov_matrix.restype = _i64
        L.fo_make_cholesky_cov_matrix.argtypes = [kd, _dp, _i64, _i64, _i64, C.c_double, C.c_int, C.c_double, _dp, _i64]
        L.fo_solve_lower.restype = C.c_int
        L.fo_solve_lower.argtypes = [_dp, _i64, _i64, _dp, _i64, _i64]
        L.fo_ad_solve_lower.restype = None
        L.fo_ad_solve_lower.argtypes = [_dp, _i64, _i64, _dp, _i64, _i64]
        L.fo_chol_solve.restype = None
        L.fo_chol_solve.argtypes = [_dp, _i64, _i64, _dp, _i64, _i64]
        L.fo_chol_inverse.restype = None
        L.fo_chol_inverse.argtypes = [_dp, _i64, _i64, _dp]
        L.fo_add_rows_cholesky.restype = C.c_int
        L.fo_add_rows_cholesky.argtypes = [kd, _dp, _i64, _dp, _i64, _i64, _i64, _i64, C.c_double]
        L.fo_make_gradient_covariance_matrices.restype = None
        L.fo_make_gradient_covariance_matrices.argtypes = [kd, _dp, _i64, _i64, _i64, _dp]
        L.fo_fit_bandwidth_mean.restype = C.c_double
        L.fo_fit_bandwidth_mean.argtypes = [_dp, _i64, _i64, _i64]
        L.fo_mean.restype = C.c_double
        L.fo_mean.argtypes = [_dp, _i64]
        L.fo_variance.restype = C.c_double
        L.fo_variance.argtypes = [_dp, _i64]
        L.fo_likelihood.restype = C.c_double
        L.fo_likelihood.argtypes = [kd, C.c_double, _dp, _i64, _i64, _i64, _dp, _dp, _i64, C.POINTER(C.c_int)]
        L.fo_predict.restype = None
        L.fo_predict.argtypes = [kd, _dp, _i64, _i64, _i64, _dp, _dp, _i64, _dp, _i64, _i64, _dp]
        L.fo_predict_variance.restype = C.c_int
        L.fo_predict_variance.argtypes = [kd, _dp, _i64, _i64, _i64, _dp, _i64, _dp, _i64, _i64, _dp]
        L.fo_predict_mean_variance.restype = None
        L.fo_predict_mean_variance.argtypes = [kd, _dp, _i64, _i64, _i64, _dp, _dp, _i64, _dp, _i64, _i64, _dp, _dp]
        L.fo_predict_covariance.restype = C.c_int
        L.fo_predict_covariance.argtypes = [kd, _dp, _i64, _i64, _i64, _dp, _dp, _i64, _dp, _i64, _i64, C.c_int, _dp, _dp]
        L.fo_gradient_marginal_likelihood.restype = None
        L.fo_gradient_marginal_likelihood.argtypes = [kd, C.c_double, _dp, _i64, _i64, _i64, _dp, _dp, _i64, C.c_int,
                                                      _dp, _dp]
        L.fo_optimize_parameters.restype = _i64
        L.fo_optimize_parameters.argtypes = [kd, _dp, _dp, _i64, _i64, _i64, _dp, _dp, _i64, C.c_int, C.c_double,
                                             C.c_int, _i64, C.c_double, C.POINTER(OptRecord)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def fcol(a):
    """float64 column-major (Fortran) copy — nalgebra DMatrix layout."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


# ---------------------------------------------------------------------------------------------------------------
# thin functional wrappers (column-major numpy arrays in, numpy out)

def kernel_value(k, x1, x2):
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    return lib().fo_kernel(C.byref(k), _p(x1), 1, _p(x2), 1, len(x1))


def kernel_gradient(k, x1, x2):
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    g = np.zeros(max(k.nparams(), 1))
    lib().fo_kernel_gradient(C.byref(k), _p(x1), 1, _p(x2), 1, len(x1), _p(g))
    return g[:k.nparams()]


def make_covariance_matrix(k, m1, m2):
    m1, m2 = fcol(m1), fcol(m2)
    out = np.zeros((m1.shape[0], m2.shape[0]), order="F")
    lib().fo_make_covariance_matrix(C.byref(k), _p(m1), m1.shape[0], m1.shape[0], _p(m2), m2.shape[0], m2.shape[0],
                                    m1.shape[1], _p(out), out.shape[0])
    return out


def gram_lower(k, X, noise):
    X = fcol(X)
    n = X.shape[0]
    K = np.zeros((n, n), order="F")
    lib().fo_gram_lower(C.byref(k), _p(X), n, n, X.shape[1], noise, _p(K), n)
    return K


def cholesky_inplace(A, substitute=None):
    """A: column-major n x n (lower used). Returns (fail_col_or_0)."""
    assert A.flags.f_contiguous
    n = A.shape[0]
    return lib().fo_cholesky_inplace(_p(A), n, n, int(substitute is not None), float(substitute or 0.0))


def make_cholesky_cov_matrix(k, X, noise, eps=None):
    X = fcol(X)
    n = X.shape[0]
    L = np.zeros((n, n), order="F")
    fail = lib().fo_make_cholesky_cov_matrix(C.byref(k), _p(X), n, n, X.shape[1], noise, int(eps is not None),
                                             float(eps or 0.0), _p(L), n)
    return L, fail


def solve_lower(L, B):
    B = fcol(B).copy(order="F")
    B2 = B.reshape(B.shape[0], -1, order="F")
    ok = lib().fo_solve_lower(_p(L), L.shape[0], L.shape[0], _p(B2), B2.shape[0], B2.shape[1])
    return B, ok


def chol_solve(L, B):
    B = fcol(B).copy(order="F")
    B2 = B.reshape(B.shape[0], -1, order="F")
    lib().fo_chol_solve(_p(L), L.shape[0], L.shape[0], _p(B2), B2.shape[0], B2.shape[1])
    return B


def chol_inverse(L):
    n = L.shape[0]
    out = np.zeros((n, n), order="F")
    lib().fo_chol_inverse(_p(L), n, n, _p(out))
    return out


def fit_bandwidth_mean(X):
    X = fcol(X)
    return lib().fo_fit_bandwidth_mean(_p(X), X.shape[0], X.shape[0], X.shape[1])


def variance(y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().fo_variance(_p(y), len(y))


def gradient_covariance_matrices(k, X):
    X = fcol(X)
    n = X.shape[0]
    P = k.nparams()
    out = np.zeros((max(P, 1), n, n))
    lib().fo_make_gradient_covariance_matrices(C.byref(k), _p(X), n, n, X.shape[1], _p(out))
    return [np.asfortranarray(out[p].T) for p in range(P)]


# ---------------------------------------------------------------------------------------------------------------
# priors (src/parameters/prior.rs) — host-side numpy

class ZeroPrior:
    """prior.rs:43-56"""
    @staticmethod
    def default(d):
        return ZeroPrior()

    def prior(self, X):
        return np.zeros(np.atleast_2d(X).shape[0])

    def fit(self, X, y):
        pass


class ConstantPrior:
    """prior.rs:66-99"""
    def __init__(self, c=0.0):
        self.c = float(c)

    @staticmethod
    def default(d):
        return ConstantPrior(0.0)

    def prior(self, X):
        return np.full(np.atleast_2d(X).shape[0], self.c)

    def fit(self, X, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        self.c = lib().fo_mean(_p(y), len(y))


class LinearPrior:
    """prior.rs:108-160 (fit = SVD least squares on [1 | X])"""
    def __init__(self, weights, intercept=0.0):
        self.weights = np.asarray(weights, dtype=np.float64)
        self.intercept = float(intercept)

    @staticmethod
    def default(d):
        return LinearPrior(np.zeros(d), 0.0)

    def prior(self, X):
        return np.atleast_2d(X) @ self.weights + self.intercept

    def fit(self, X, y):
        A = np.hstack([np.ones((X.shape[0], 1)), X])
        w, *_ = np.linalg.lstsq(A, y, rcond=None)
        self.intercept = float(w[0])
        self.weights = w[1:].copy()


# ---------------------------------------------------------------------------------------------------------------

class OracleGaussianProcess:
    """Reference flow of `GaussianProcess` (mod.rs:58-446) + builder defaults (builder.rs:66-95, :189-214)."""

    def __init__(self, prior, kernel: KernelDesc, noise, cholesky_epsilon, X, y):
        assert noise >= 0.0  # mod.rs:150
        self.prior, self.kernel, self.noise, self.cholesky_epsilon = prior, kernel.copy(), float(noise), cholesky_epsilon
        self.X = fcol(np.atleast_2d(X))
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        assert self.X.shape[0] == y.shape[0]  # mod.rs:153
        self.y = y - prior.prior(self.X)  # mod.rs:156
        self.L, fail = make_cholesky_cov_matrix(self.kernel, self.X, self.noise, cholesky_epsilon)
        if fail:
            raise ArithmeticError(f"Cholesky decomposition failed at column {fail - 1}")

    # ---- builder.rs:66-95 + :189-214 ----
    @classmethod
    def train(cls, X, y, kernel=None, prior=None, noise=None, cholesky_epsilon=None, fit_kernel=False, fit_prior=False,
              max_iter=100, convergence_fraction=0.05):
        X = fcol(np.atleast_2d(X))
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        kernel = (kernel or KernelDesc.make([K_SQUARED_EXP], [1.0, 1.0])).copy()
        prior = prior if prior is not None else ConstantPrior.default(X.shape[1])
        if noise is None:
            noise = 0.1 * math.sqrt(variance(y))  # builder.rs:73
        if fit_kernel:
            heuristic_fit(kernel, X, y)  # builder.rs:193-196
        gp = cls(prior, kernel, noise, cholesky_epsilon, X, y)
        gp.fit_parameters(fit_prior, fit_kernel, max_iter, convergence_fraction)
        return gp

    @property
    def n(self):
        return self.X.shape[0]

    def _refit(self):
        self.L, fail = make_cholesky_cov_matrix(self.kernel, self.X, self.noise, self.cholesky_epsilon)
        if fail:
            raise ArithmeticError(f"Cholesky decomposition failed at column {fail - 1}")

    def add_samples(self, Xn, yn):  # mod.rs:173-190
        Xn = fcol(np.atleast_2d(Xn))
        yn = np.asarray(yn, dtype=np.float64).reshape(-1) - self.prior.prior(Xn)
        n_old, k = self.n, Xn.shape[0]
        self.X = fcol(np.vstack([self.X, Xn]))
        self.y = np.concatenate([self.y, yn])
        Lnew = np.full((n_old + k, n_old + k), np.nan, order="F")
        Lnew[:n_old, :n_old] = self.L
        ok = lib().fo_add_rows_cholesky(C.byref(self.kernel), _p(Lnew), n_old + k, _p(self.X), n_old + k, n_old, k,
                                        self.X.shape[1], self.noise)
        assert ok, "Cholesky::insert_column : Unable to solve lower triangular system!"
        self.L = Lnew

    def likelihood(self):  # mod.rs:196-220
        ok = C.c_int(0)
        v = lib().fo_likelihood(C.byref(self.kernel), self.noise, _p(self.X), self.n, self.n, self.X.shape[1],
                                _p(self.y), _p(self.L), self.n, C.byref(ok))
        assert ok.value, "likelihood : solve failed"
        return v

    def predict(self, Xq):  # mod.rs:226-244
        Xq = fcol(np.atleast_2d(Xq))
        mean = np.ascontiguousarray(self.prior.prior(Xq), dtype=np.float64)
        lib().fo_predict(C.byref(self.kernel), _p(self.X), self.n, self.n, self.X.shape[1], _p(self.y), _p(self.L),
                         self.n, _p(Xq), Xq.shape[0], Xq.shape[0], _p(mean))
        return mean

    def predict_variance(self, Xq):  # mod.rs:248-273
        Xq = fcol(np.atleast_2d(Xq))
        var = np.zeros(Xq.shape[0])
        ok = lib().fo_predict_variance(C.byref(self.kernel), _p(self.X), self.n, self.n, self.X.shape[1], _p(self.L),
                                       self.n, _p(Xq), Xq.shape[0], Xq.shape[0], _p(var))
        assert ok, "predict_covariance : solve failed"
        return var

    def predict_mean_variance(self, Xq):  # mod.rs:290-326
        Xq = fcol(np.atleast_2d(Xq))
        mean = np.ascontiguousarray(self.prior.prior(Xq), dtype=np.float64)
        var = np.zeros(Xq.shape[0])
        lib().fo_predict_mean_variance(C.byref(self.kernel), _p(self.X), self.n, self.n, self.X.shape[1], _p(self.y),
                                       _p(self.L), self.n, _p(Xq), Xq.shape[0], Xq.shape[0], _p(mean), _p(var))
        return mean, var

    def predict_covariance(self, Xq):  # mod.rs:329-350
        Xq = fcol(np.atleast_2d(Xq))
        q = Xq.shape[0]
        cov = np.zeros((q, q), order="F")
        ok = lib().fo_predict_covariance(C.byref(self.kernel), _p(self.X), self.n, self.n, self.X.shape[1], _p(self.y),
                                         _p(self.L), self.n, _p(Xq), q, q, 0, _p(cov), None)
        assert ok, "predict_covariance : solve failed"
        return cov

    def sample_at_params(self, Xq):  # mod.rs:371-392 -> (mean, cov) fed to MultivariateNormal::new
        Xq = fcol(np.atleast_2d(Xq))
        q = Xq.shape[0]
        cov = np.zeros((q, q), order="F")
        mean = np.ascontiguousarray(self.prior.prior(Xq), dtype=np.float64)
        lib().fo_predict_covariance(C.byref(self.kernel), _p(self.X), self.n, self.n, self.X.shape[1], _p(self.y),
                                    _p(self.L), self.n, _p(Xq), q, q, 1, _p(cov), _p(mean))
        return mean, cov

    def gradient_marginal_likelihood(self, scaled):  # optimizer.rs:24-60 / :159-203
        P = self.kernel.nparams()
        grads = np.zeros(P + 1)
        scale = C.c_double(1.0)
        lib().fo_gradient_marginal_likelihood(C.byref(self.kernel), self.noise, _p(self.X), self.n, self.n,
                                              self.X.shape[1], _p(self.y), _p(self.L), self.n, int(scaled),
                                              C.cast(C.byref(scale), _dp), _p(grads))
        return (scale.value, grads[:P]) if scaled else grads

    def fit_parameters(self, fit_prior, fit_kernel, max_iter=100, convergence_fraction=0.05):  # mod.rs:406-445
        self.trace = []
        if fit_prior:
            y_raw = self.y + self.prior.prior(self.X)
            self.prior.fit(self.X, y_raw)
            self.y = y_raw - self.prior.prior(self.X)
            if not fit_kernel:
                self._refit()
        if fit_kernel:
            scaled = lib().fo_desc_is_scalable(C.byref(self.kernel)) == 1
            noise = np.array([self.noise])
            trace = (OptRecord * max(max_iter, 1))()
            eps = self.cholesky_epsilon
            it = lib().fo_optimize_parameters(C.byref(self.kernel), _p(noise), _p(self.X), self.n, self.n,
                                              self.X.shape[1], _p(self.y), _p(self.L), self.n, int(eps is not None),
                                              float(eps or 0.0), int(scaled), max_iter, convergence_fraction, trace)
            self.noise = float(noise[0])
            if it < 0:
                raise ArithmeticError(f"Cholesky decomposition failed at column {-it - 1}")
            P = self.kernel.nparams()
            np_ = P if scaled else P + 1
            self.trace = [dict(scale=trace[i].scale, grads=[trace[i].grads[p] for p in range(np_)],
                               params=[trace[i].params[p] for p in range(P)], noise=trace[i].noise)
                          for i in range(it)]


def heuristic_fit(kernel: KernelDesc, X, y):
    """`heuristic_fit` of every leaf (kernel.rs:594-600, :699-705, :806-812, :918-924; Sum/Prod :194-200, :290-296).
    Only the four ls/ampl kernels implement it; the others keep the default no-op (kernel.rs:81-85)."""
    po = 0
    ls = ampl = None
    for i in range(kernel.n_ops):
        tag = kernel.op[i]
        if tag in (K_SQUARED_EXP, K_EXPONENTIAL, K_MATERN1, K_MATERN2):
            if ls is None:
                ls, ampl = fit_bandwidth_mean(X), variance(y)
            kernel.param[po], kernel.param[po + 1] = ls, ampl
        po += _NPARAMS[tag]
