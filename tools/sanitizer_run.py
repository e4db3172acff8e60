"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): the head kernel alone on 1..4 tiles, a fit
through the head schedule (two panels + a partial one), predict (batched and latency paths), add_samples, LML gradient."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import friedrich_b200 as F  # noqa: E402
from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

for nt in (1, 4):
    n = 128 * nt
    rng = np.random.default_rng(nt)
    M = rng.standard_normal((n, n + 16))
    A = np.asfortranarray(M @ M.T / n + 0.05 * np.eye(n))
    W = np.zeros((n, n), order="F")
    info = C.c_int(-1)
    rc = N.lib().fgp_dbg_potrf_head(0, N.dptr(A), nt, N.dptr(W), 0, 0.0, C.byref(info), 0, None)
    print("head nt", nt, "rc", rc, "info", info.value)
n, d = 1200, 4
X, y = make_dataset(3, n, d)
gp = F.GaussianProcess(F.ZeroPrior(), F.Matern2(0.8, 1.0), 0.1, None, X[:1000], y[:1000])
Xq = make_inputs(4, 200, d)
m, v = gp.predict_mean_variance(Xq)
m2, v2 = gp.predict_mean_variance(Xq[:5])
gp.add_samples(X[1000:], y[1000:])
s, g = gp.scaled_gradient_marginal_likelihood()
cov = gp.predict_covariance(Xq[:40])
print("ok", float(m[0]), float(v2[0]), s, g, float(cov[0, 0]))
