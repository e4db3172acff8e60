"""Host-side mirror of the reference's public API (`GaussianProcess`, `GaussianProcessBuilder`, `MultivariateNormal`,
priors) on top of the C-ABI of libfgp_sm100.so.

Same names, argument meaning and error behaviour as src/gaussian_process/{mod,builder,optimizer,multivariate_normal}.rs
and src/parameters/prior.rs.  Everything O(n^2) or larger happens on the GPU behind `friedrich_b200._native`; this file
only keeps what the reference also keeps outside `algebra`/nalgebra: the prior (O(n d)), the ADAM scalar loop, argument
checks and type adaptation (`Input`, src/conversion/mod.rs).  There is no CPU fallback for the device work.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np

from . import _native as N
from .kernels import Gaussian, Kernel

__all__ = ["GaussianProcess", "GaussianProcessBuilder", "MultivariateNormal", "ZeroPrior", "ConstantPrior",
           "LinearPrior"]


# ---------------------------------------------------------------------------------------------------------------------
# src/conversion/mod.rs — `Input`: a 2-D array-like is a matrix with one sample per row; a 1-D array-like is ONE sample
# (Vec<f64>, conversion/mod.rs:95-118) and makes the scalar-returning variants of the predict functions.

def _as_matrix(inputs):
    a = np.asarray(inputs, dtype=np.float64)
    single = a.ndim == 1
    if single:
        a = a.reshape(1, -1)
    if a.ndim != 2:
        raise ValueError("inputs must be a matrix (one sample per row) or a single sample")
    return N.fcol(a), single


def _as_vector(outputs):
    return np.ascontiguousarray(np.asarray(outputs, dtype=np.float64).reshape(-1))


# ---------------------------------------------------------------------------------------------------------------------
# src/parameters/prior.rs

class ZeroPrior:
    """prior.rs:43-56"""

    @staticmethod
    def default(input_dimension):
        return ZeroPrior()

    def prior(self, X):
        return np.zeros(X.shape[0])

    def fit(self, X, y):
        pass


class ConstantPrior:
    """prior.rs:66-99"""

    def __init__(self, c=0.0):
        self.c = float(c)

    @staticmethod
    def default(input_dimension):
        return ConstantPrior(0.0)

    def prior(self, X):
        return np.full(X.shape[0], self.c)

    def fit(self, X, y):
        self.c = float(np.sum(y) / len(y))  # nalgebra mean(): sum / len


class LinearPrior:
    """prior.rs:108-160; fit = least squares on [1 | X] through an SVD (prior.rs:144-148)."""

    def __init__(self, weights, intercept=0.0):
        self.weights = np.asarray(weights, dtype=np.float64).reshape(-1)
        self.intercept = float(intercept)

    @staticmethod
    def default(input_dimension):
        return LinearPrior(np.zeros(input_dimension), 0.0)

    def prior(self, X):
        return X @ self.weights + self.intercept

    def fit(self, X, y):
        A = np.hstack([np.ones((X.shape[0], 1)), X])
        w, *_ = np.linalg.lstsq(A, y, rcond=None)
        self.intercept = float(w[0])
        self.weights = w[1:].copy()


def _population_variance(y):
    """nalgebra `variance` / `row_variance` (population, two-pass) — fit_amplitude_var kernel.rs:116-119, builder.rs:73."""
    mean = float(np.sum(y) / len(y))
    return float(np.sum((y - mean) * (y - mean)) / len(y))


# ---------------------------------------------------------------------------------------------------------------------

class MultivariateNormal:
    """src/gaussian_process/multivariate_normal.rs:20-74.  The q x q Cholesky factor is computed by the same device
    factorisation as the model's (a scratch handle on the model's GPU)."""

    def __init__(self, mean, covariance, single=False, device=0):
        self._mean = np.asarray(mean, dtype=np.float64)
        self._single = single
        L = N.fcol(covariance, copy=True)
        q = L.shape[0]
        failed = C.c_int64(-1)
        rc = N.lib().fgp_cholesky_lower(int(device), N.dptr(L), q, q, C.byref(failed))  # covariance.cholesky().unpack()
        if rc == N.FGP_ERR_NOT_POSDEF:
            raise N.NotPositiveDefinite(rc, f"MultivariateNormal: Cholesky decomposition failed! (column {failed.value})")
        if rc != N.FGP_OK:
            raise N.FgpError(rc, "fgp_cholesky_lower failed")
        self.cholesky_covariance = L

    def mean(self):
        return float(self._mean[0]) if self._single else self._mean.copy()

    def sample(self, rng):
        normal = rng.standard_normal(self._mean.shape[0])
        s = self._mean + self.cholesky_covariance @ normal  # O(q^2) host product, as in the reference (:71)
        return float(s[0]) if self._single else s


# ---------------------------------------------------------------------------------------------------------------------

class GaussianProcess:
    """src/gaussian_process/mod.rs:58-446.

    Public fields as in the reference: `prior`, `kernel`, `noise`, `cholesky_epsilon` (mod.rs:62-73).  The private
    ones (training_inputs, training_outputs, covmat_cholesky) live on the GPU inside the native handle.
    """

    # -- constructors -----------------------------------------------------------------------------------------------
    def __init__(self, prior, kernel: Kernel, noise, cholesky_epsilon, training_inputs, training_outputs, device=0):
        """`GaussianProcess::new` (mod.rs:142-167)."""
        X, _ = _as_matrix(training_inputs)
        y = _as_vector(training_outputs)
        assert noise >= 0.0, f"The noise parameter should non-negative but we tried to set it to {noise}"  # mod.rs:150
        assert X.shape[0] == y.shape[0]  # mod.rs:153
        self.prior, self.kernel, self.noise, self.cholesky_epsilon = prior, kernel, float(noise), cholesky_epsilon
        self._h = N.Handle(device)
        self._d = X.shape[1]
        y_resid = np.ascontiguousarray(y - prior.prior(X))  # mod.rs:156
        self._X_host, self._y_host = X, y_resid  # host twins of training_inputs / training_outputs (O(n d); prior refits)
        self._fit(X, y_resid)

    @classmethod
    def default(cls, training_inputs, training_outputs, device=0):
        """mod.rs:96-102: default builder with kernel and prior fitting switched on."""
        return GaussianProcessBuilder(training_inputs, training_outputs, device=device).fit_kernel().fit_prior().train()

    @classmethod
    def builder(cls, training_inputs, training_outputs, device=0):
        """mod.rs:130-136"""
        return GaussianProcessBuilder(training_inputs, training_outputs, device=device)

    @classmethod
    def _from_handle(cls, handle, prior, kernel, noise, cholesky_epsilon, X, y):
        """builder.train(): the inputs are already resident (heuristic_fit ran on them)."""
        self = cls.__new__(cls)
        assert noise >= 0.0, f"The noise parameter should non-negative but we tried to set it to {noise}"
        assert X.shape[0] == y.shape[0]
        self.prior, self.kernel, self.noise, self.cholesky_epsilon = prior, kernel, float(noise), cholesky_epsilon
        self._h = handle
        self._d = X.shape[1]
        y_resid = np.ascontiguousarray(y - prior.prior(X))
        self._X_host, self._y_host = X, y_resid
        self._h.check(N.lib().fgp_set_outputs(self._h.ptr, N.dptr(y_resid), len(y_resid)))
        self._refit()
        return self

    # -- native plumbing ----------------------------------------------------------------------------------------------
    def _eps(self):
        e = self.cholesky_epsilon
        return (0, 0.0) if e is None else (1, float(e))

    def _desc(self):
        return self.kernel.device_desc()

    def _fit(self, X, y_resid):
        has_eps, eps = self._eps()
        kd = self._desc()
        rc = N.lib().fgp_fit(self._h.ptr, N.dptr(X), X.shape[0], X.shape[0], X.shape[1], N.dptr(y_resid), C.byref(kd),
                             self.noise, has_eps, eps)
        self._check_fit(rc)

    def _refit(self):
        """make_cholesky_cov_matrix on the resident inputs (mod.rs:426-429, optimizer.rs:133-136, :267-270)."""
        has_eps, eps = self._eps()
        kd = self._desc()
        self._check_fit(N.lib().fgp_refit(self._h.ptr, C.byref(kd), self.noise, has_eps, eps))

    def _check_fit(self, rc):
        if rc == N.FGP_ERR_NOT_POSDEF:
            col = N.lib().fgp_failed_column(self._h.ptr)
            if self.cholesky_epsilon is None:  # algebra/mod.rs:90
                raise N.NotPositiveDefinite(rc, "Cholesky decomposition failed! Try using the `set_cholesky_epsilon` "
                                                f"method on the builder (column {col})")
            raise N.NotPositiveDefinite(rc, f"Cholesky decomposition failed! (column {col})")  # algebra/mod.rs:85
        self._h.check(rc)

    def _queries(self, inputs):
        Xq, single = _as_matrix(inputs)
        assert Xq.shape[1] == self._d  # mod.rs:231, :253, :293, :334, :374
        return Xq, single

    @property
    def n_samples(self):
        return int(N.lib().fgp_num_samples(self._h.ptr))

    def cholesky_factor(self):
        """The matrix held by `covmat_cholesky` (lower valid, strict upper NaN: algebra/mod.rs:67) — serde / tests."""
        n = self.n_samples
        L = np.zeros((n, n), order="F")
        self._h.check(N.lib().fgp_download_factor(self._h.ptr, N.dptr(L), n))
        return L

    # -- serde (feature friedrich_serde: mod.rs:58, extendable_matrix.rs:14,62, kernel.rs:506, prior.rs:42) ------------
    def to_state(self):
        """What `#[derive(Serialize)]` writes for a GaussianProcess: prior, kernel, noise, cholesky_epsilon, training inputs
        and (residual) outputs, and covmat_cholesky as nalgebra keeps it (n x n, factor in the lower triangle, NaN above).
        Plain Python / numpy objects: pickle, json or npz them as you like."""
        return dict(prior=self.prior, kernel=self.kernel, noise=self.noise, cholesky_epsilon=self.cholesky_epsilon,
                    training_inputs=self._X_host.copy(order="F"), training_outputs=self._y_host.copy(),
                    covmat_cholesky=self.cholesky_factor())

    @classmethod
    def from_state(cls, state, device=0):
        """`Deserialize`: the model goes back to the GPU without refitting (fgp_upload_state rebuilds only what the device
        path caches on top of the reference's state)."""
        self = cls.__new__(cls)
        self.prior, self.kernel = state["prior"], state["kernel"]
        self.noise, self.cholesky_epsilon = float(state["noise"]), state["cholesky_epsilon"]
        X = N.fcol(state["training_inputs"])
        y = np.ascontiguousarray(state["training_outputs"], dtype=np.float64)
        L = N.fcol(state["covmat_cholesky"])
        assert X.shape[0] == y.shape[0] == L.shape[0] == L.shape[1]
        self._h = N.Handle(device)
        self._d = X.shape[1]
        self._X_host, self._y_host = X, y
        self._h.check(N.lib().fgp_upload_state(self._h.ptr, N.dptr(X), X.shape[0], X.shape[0], X.shape[1], N.dptr(y),
                                               N.dptr(L), L.shape[0]))
        return self

    def inverse_columns(self, cols):
        """Columns of K^-1 left on the device by the last (scaled_)gradient_marginal_likelihood call (diagnostics)."""
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        out = np.zeros((self.n_samples, len(cols)), order="F")
        self._h.check(N.lib().fgp_inverse_columns(self._h.ptr, cols.ctypes.data_as(C.POINTER(C.c_int64)), len(cols),
                                                  N.dptr(out), out.shape[0]))
        return out

    # -- mod.rs:173-190 -------------------------------------------------------------------------------------------------
    def add_samples(self, inputs, outputs):
        X, _ = _as_matrix(inputs)
        y = _as_vector(outputs)
        assert X.shape[0] == y.shape[0]  # mod.rs:177
        assert X.shape[1] == self._d     # mod.rs:178
        y_resid = np.ascontiguousarray(y - self.prior.prior(X))  # mod.rs:180
        has_eps, eps = self._eps()
        kd = self._desc()
        rc = N.lib().fgp_add_samples(self._h.ptr, N.dptr(X), X.shape[0], X.shape[0], N.dptr(y_resid), C.byref(kd),
                                     self.noise, has_eps, eps)
        self._check_fit(rc)
        self._X_host = N.fcol(np.vstack([self._X_host, X]))  # mod.rs:181-182
        self._y_host = np.concatenate([self._y_host, y_resid])

    # -- mod.rs:196-220 -------------------------------------------------------------------------------------------------
    def likelihood(self):
        out = C.c_double(0.0)
        kd = self._desc()
        self._h.check(N.lib().fgp_likelihood(self._h.ptr, C.byref(kd), self.noise, C.cast(C.byref(out), N._dp)))
        return out.value

    # -- mod.rs:226-244 -------------------------------------------------------------------------------------------------
    def predict(self, inputs):
        Xq, single = self._queries(inputs)
        mean = np.zeros(Xq.shape[0])
        kd = self._desc()
        self._h.check(N.lib().fgp_predict_mean(self._h.ptr, C.byref(kd), N.dptr(Xq), Xq.shape[0], Xq.shape[0],
                                               N.dptr(mean)))
        mean += self.prior.prior(Xq)  # mod.rs:238-241
        return float(mean[0]) if single else mean

    # -- mod.rs:248-273 -------------------------------------------------------------------------------------------------
    def predict_variance(self, inputs):
        Xq, single = self._queries(inputs)
        var = np.zeros(Xq.shape[0])
        kd = self._desc()
        self._h.check(N.lib().fgp_predict_var(self._h.ptr, C.byref(kd), N.dptr(Xq), Xq.shape[0], Xq.shape[0],
                                              N.dptr(var)))
        return float(var[0]) if single else var

    # -- mod.rs:290-326 -------------------------------------------------------------------------------------------------
    def predict_mean_variance(self, inputs):
        Xq, single = self._queries(inputs)
        mean, var = np.zeros(Xq.shape[0]), np.zeros(Xq.shape[0])
        kd = self._desc()
        self._h.check(N.lib().fgp_predict_mean_var(self._h.ptr, C.byref(kd), N.dptr(Xq), Xq.shape[0], Xq.shape[0],
                                                   N.dptr(mean), N.dptr(var)))
        mean += self.prior.prior(Xq)
        return (float(mean[0]), float(var[0])) if single else (mean, var)

    # -- mod.rs:329-350 -------------------------------------------------------------------------------------------------
    def predict_covariance(self, inputs):
        Xq, _ = self._queries(inputs)
        q = Xq.shape[0]
        cov = np.zeros((q, q), order="F")
        kd = self._desc()
        self._h.check(N.lib().fgp_predict_cov(self._h.ptr, C.byref(kd), N.dptr(Xq), q, q, 0, N.dptr(cov), q, None))
        return cov

    # -- mod.rs:371-392 -------------------------------------------------------------------------------------------------
    def sample_at(self, inputs):
        Xq, single = self._queries(inputs)
        q = Xq.shape[0]
        cov = np.zeros((q, q), order="F")
        mean = np.zeros(q)
        kd = self._desc()
        self._h.check(N.lib().fgp_predict_cov(self._h.ptr, C.byref(kd), N.dptr(Xq), q, q, 1, N.dptr(cov), q,
                                              N.dptr(mean)))
        mean += self.prior.prior(Xq)
        return MultivariateNormal(mean, cov, single=single, device=self._h.device)

    # -- optimizer.rs:24-60 / :159-203 ---------------------------------------------------------------------------------
    def gradient_marginal_likelihood(self):
        P = self.kernel.nb_parameters()
        grads = np.zeros(P + 1)
        kd = self._desc()
        self._h.check(N.lib().fgp_lml_gradient(self._h.ptr, C.byref(kd), self.noise, 0, None, N.dptr(grads)))
        return list(grads)

    def scaled_gradient_marginal_likelihood(self):
        P = self.kernel.nb_parameters()
        grads = np.zeros(P + 1)
        scale = C.c_double(1.0)
        kd = self._desc()
        self._h.check(N.lib().fgp_lml_gradient(self._h.ptr, C.byref(kd), self.noise, 1, C.cast(C.byref(scale), N._dp),
                                               N.dptr(grads)))
        return scale.value, list(grads[:P])

    # -- mod.rs:406-445 -------------------------------------------------------------------------------------------------
    def fit_parameters(self, fit_prior, fit_kernel, max_iter=100, convergence_fraction=0.05, max_time=3600.0):
        self.trace = []
        if fit_prior:
            X, y_resid = self._X_host, self._y_host
            y_raw = y_resid + self.prior.prior(X)                 # mod.rs:415
            if isinstance(self.prior, LinearPrior) and X.shape[1] <= 44:
                # prior.rs:139-159 on the resident inputs (fgp_linear_prior_fit): normal equations on the device, d x d solve
                w = np.zeros(X.shape[1])
                b = C.c_double(0.0)
                self._h.check(N.lib().fgp_linear_prior_fit(self._h.ptr, N.dptr(np.ascontiguousarray(y_raw)), N.dptr(w),
                                                           C.cast(C.byref(b), N._dp)))
                self.prior.weights, self.prior.intercept = w, b.value
            else:
                self.prior.fit(X, y_raw)                           # mod.rs:417
            y_resid = np.ascontiguousarray(y_raw - self.prior.prior(X))   # mod.rs:419-420
            self._y_host = y_resid
            self._h.check(N.lib().fgp_set_outputs(self._h.ptr, N.dptr(y_resid), len(y_resid)))
            if not fit_kernel:
                self._refit()                                      # mod.rs:423-430
        if fit_kernel:
            if self.kernel.is_scalable():
                self._scaled_optimize_parameters(max_iter, convergence_fraction, max_time)   # mod.rs:436-438
            else:
                self._optimize_parameters(max_iter, convergence_fraction, max_time)          # mod.rs:440-443

    # -- optimizer.rs:69-149 ----------------------------------------------------------------------------------------------
    def _optimize_parameters(self, max_iter, convergence_fraction, max_time):
        beta1, beta2, epsilon, learning_rate = 0.9, 0.999, 1e-8, 0.1
        parameters = [epsilon if p == 0.0 else p for p in self.kernel.get_parameters()]
        parameters.append(math.log(self.noise))
        mean_grad = [0.0] * len(parameters)
        var_grad = [0.0] * len(parameters)
        t0 = time.monotonic()
        for i in range(1, max_iter + 1):
            gradients = self.gradient_marginal_likelihood()
            gradients[-1] *= self.noise                                            # optimizer.rs:106-110
            progress = False
            for p in range(len(parameters)):
                mean_grad[p] = beta1 * mean_grad[p] + (1.0 - beta1) * gradients[p]
                var_grad[p] = beta2 * var_grad[p] + (1.0 - beta2) * gradients[p] ** 2
                bcm = mean_grad[p] / (1.0 - beta1 ** i)
                bcv = var_grad[p] / (1.0 - beta2 ** i)
                delta = learning_rate * bcm / (math.sqrt(bcv) + epsilon)
                progress |= abs(delta) > convergence_fraction
                parameters[p] *= 1.0 + delta
            self.kernel.set_parameters(parameters[:-1])
            self.noise = math.exp(parameters[-1])
            self._refit()
            self.trace.append(dict(scale=1.0, grads=list(gradients), params=self.kernel.get_parameters(),
                                   noise=self.noise))
            if (not progress) or (time.monotonic() - t0 > max_time):
                break

    # -- optimizer.rs:211-283 ---------------------------------------------------------------------------------------------
    def _scaled_optimize_parameters(self, max_iter, convergence_fraction, max_time):
        beta1, beta2, epsilon, learning_rate = 0.9, 0.999, 1e-8, 0.1
        parameters = [epsilon if p == 0.0 else p for p in self.kernel.get_parameters()]
        mean_grad = [0.0] * len(parameters)
        var_grad = [0.0] * len(parameters)
        t0 = time.monotonic()
        for i in range(1, max_iter + 1):
            scale, gradients = self.scaled_gradient_marginal_likelihood()
            progress = False
            for p in range(len(parameters)):
                mean_grad[p] = beta1 * mean_grad[p] + (1.0 - beta1) * gradients[p]
                var_grad[p] = beta2 * var_grad[p] + (1.0 - beta2) * gradients[p] ** 2
                bcm = mean_grad[p] / (1.0 - beta1 ** i)
                bcv = var_grad[p] / (1.0 - beta2 ** i)
                delta = learning_rate * bcm / (math.sqrt(bcv) + epsilon)
                progress |= abs(delta) > convergence_fraction
                parameters[p] *= 1.0 + delta
            self.kernel.set_parameters(parameters)
            self.kernel.rescale(scale)                                             # optimizer.rs:262
            self.noise *= scale                                                    # optimizer.rs:263
            parameters = self.kernel.get_parameters()                              # optimizer.rs:264
            self._refit()
            self.trace.append(dict(scale=scale, grads=list(gradients), params=self.kernel.get_parameters(),
                                   noise=self.noise))
            if (not progress) or (time.monotonic() - t0 > max_time):
                break


# ---------------------------------------------------------------------------------------------------------------------

class GaussianProcessBuilder:
    """src/gaussian_process/builder.rs:36-215"""

    def __init__(self, training_inputs, training_outputs, device=0):
        X, _ = _as_matrix(training_inputs)
        y = _as_vector(training_outputs)
        self.training_inputs, self.training_outputs = X, y
        self.prior = ConstantPrior.default(X.shape[1])          # builder.rs:71
        self.kernel = Gaussian()                                 # builder.rs:72
        self.noise = 0.1 * math.sqrt(_population_variance(y))    # builder.rs:73
        self.should_fit_kernel = False
        self.should_fit_prior = False
        self.max_iter = 100
        self.convergence_fraction = 0.05
        self.max_time = 3600.0
        self.cholesky_epsilon = None
        self.device = device

    def set_prior(self, prior):
        self.prior = prior
        return self

    def set_noise(self, noise):
        assert noise >= 0.0, f"The noise parameter should non-negative but we tried to set it to {noise}"  # builder.rs:123
        self.noise = float(noise)
        return self

    def set_kernel(self, kernel):
        self.kernel = kernel
        return self

    def set_cholesky_epsilon(self, cholesky_epsilon):
        self.cholesky_epsilon = cholesky_epsilon
        return self

    def set_fit_parameters(self, max_iter, convergence_fraction):
        self.max_iter, self.convergence_fraction = int(max_iter), float(convergence_fraction)
        return self

    def fit_kernel(self):
        self.should_fit_kernel = True
        return self

    def fit_prior(self):
        self.should_fit_prior = True
        return self

    def train(self):
        """builder.rs:189-214"""
        X, y = self.training_inputs, self.training_outputs
        handle = N.Handle(self.device)
        handle.check(N.lib().fgp_set_inputs(handle.ptr, N.dptr(X), X.shape[0], X.shape[0], X.shape[1]))
        if self.should_fit_kernel:  # builder.rs:193-196: heuristic on the RAW outputs, before the prior is subtracted
            def bandwidth_mean():
                out = C.c_double(0.0)
                handle.check(N.lib().fgp_mean_pair_distance(handle.ptr, C.cast(C.byref(out), N._dp)))
                return out.value

            self.kernel.heuristic_fit(bandwidth_mean, lambda: _population_variance(y))
        gp = GaussianProcess._from_handle(handle, self.prior, self.kernel, self.noise, self.cholesky_epsilon, X, y)
        gp.fit_parameters(self.should_fit_prior, self.should_fit_kernel, self.max_iter, self.convergence_fraction,
                          self.max_time)
        return gp
