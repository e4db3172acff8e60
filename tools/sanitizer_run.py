"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): the head kernel alone on 1..4 tiles, a fit
through the head schedule (two panels + a partial one), predict (batched and latency paths), add_samples, LML gradient; the
tcgen05 update kernel and its digit slicing alone (single-CTA and CTA-pair kernels) and inside a fit + predict + LML gradient
large enough to route through them (n = 2560: 2048 rows below the first panel, threshold 1024)."""
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

# tcgen05 path
rng = np.random.default_rng(9)
for flags in (0, 32):
    N.lib().fgp_dbg_ozaki_experiment(flags)
    for (M, K, lower, skip) in [(384, 256, 1, 0), (256, 128, 0, 0), (640, 512, 1, 2)]:
        P = np.asfortranarray(rng.standard_normal((M, K)))
        Cm = np.asfortranarray(rng.standard_normal((M, M)))
        rc = N.lib().fgp_dbg_ozaki_syrk(0, N.dptr(Cm), M, N.dptr(P), M, M, K, lower, skip, 0, 0, 0)
        print("ozaki", "pairs" if flags else "single", M, K, lower, skip, "rc", rc)
N.lib().fgp_dbg_ozaki_experiment(0)
if "--small" not in sys.argv:
    n2 = 2560
    X2, y2 = make_dataset(5, n2, d)
    gp2 = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(0.8, 1.0), 0.1, None, X2, y2)
    m3, v3 = gp2.predict_mean_variance(Xq)
    s2, g2 = gp2.scaled_gradient_marginal_likelihood()
    print("ok tcgen05 fit", float(m3[0]), float(v3[0]), s2, g2)
    # single-rank sharded fit on the row-piece schedule (csrc/sharded.cu factor_sharded_pipe): n = 2560 -> panels 0..2 travel in
    # pieces (512, 512, rest), the last two in one piece; then predict from the kept digit slices
    from friedrich_b200 import sharded  # noqa: E402
    from friedrich_b200.kernels import SquaredExp  # noqa: E402
    hs = N.Handle(0)
    sharded.comm_init(hs, 0, 1)
    assert N.lib().fgp_set_option(hs.ptr, N.FGP_OPT_SHARD_PIPE, 1) == 0   # (automatic = one piece below 3 ranks)
    kd = SquaredExp(0.8, 1.0).device_desc()
    sharded.fit_sharded(hs, X2, y2, kd, 0.1)
    mean, var = np.zeros(200), np.zeros(200)
    hs.check(N.lib().fgp_predict_mean_var(hs.ptr, C.byref(kd), N.dptr(N.fcol(Xq)), 200, 200, N.dptr(mean), N.dptr(var)))
    print("ok sharded pipe fit", float(mean[0]), float(var[0]), "plain", float(m3[0]), float(v3[0]))
    hs.close()
