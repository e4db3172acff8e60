"""GPU diagnostic: runs every C-ABI entry point at small sizes against the CPU oracle and prints the errors.
Usage (on the GPU box): python tools/gpu_diag.py [--big]
Never exits non-zero on a numeric mismatch: it is a report, the assertions live in tests/."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200 import (GaussianProcess, Matern2, SquaredExp, ZeroPrior, ConstantPrior, Exponential, Linear,  # noqa: E402
                            RationalQuadratic, Matern1, Polynomial, HyperTan, Multiquadric)
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def maxrel(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def section(name):
    print(f"\n=== {name}", flush=True)


def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        print("!!! section failed", flush=True)


def diag_gemm():
    section("gemm_nt hook")
    rng = np.random.default_rng(0)
    for (M, Nn, K, lower) in [(128, 128, 16, 0), (256, 128, 128, 0), (384, 256, 272, 0), (512, 512, 64, 1)]:
        A = np.asfortranarray(rng.standard_normal((M, K)))
        B = np.asfortranarray(rng.standard_normal((Nn, K)))
        Cm = np.asfortranarray(rng.standard_normal((M, Nn)))
        ref = Cm - 0.5 * A @ B.T
        out = Cm.copy(order="F")
        rc = N.lib().fgp_dbg_gemm_nt(0, N.dptr(out), M, N.dptr(A), M, N.dptr(B), Nn, M, Nn, K, -0.5, 1, lower)
        if lower:
            mask = np.tril(np.ones((M, Nn), dtype=bool))
            e = rel(out[mask], ref[mask])
            untouched = np.array_equal(out[~mask], Cm[~mask])
            print(f"M={M} N={Nn} K={K} lower rc={rc} rel={e:.2e} upper_untouched={untouched}")
        else:
            print(f"M={M} N={Nn} K={K} rc={rc} rel={rel(out, ref):.2e}")


def oracle_gp(kdesc, noise, X, y, eps=None):
    return O.OracleGaussianProcess(O.ZeroPrior(), O.KernelDesc.make(*kdesc), noise, eps, X, y)


def diag_fit_predict(n, d, q, kernel, kdesc, noise=0.1, tag=""):
    section(f"fit/predict n={n} d={d} q={q} {tag}")
    X, y = make_dataset(1234 + n, n, d)
    Xq = make_inputs(99 + n, q, d)
    t0 = time.time()
    gp = GaussianProcess(ZeroPrior(), kernel, noise, None, X, y)
    t1 = time.time()
    print(f"gpu fit wall {t1 - t0:.3f}s device {gp._h.last_device_ms():.3f} ms launches {gp._h.last_launch_count()}")
    L = gp.cholesky_factor()
    t0 = time.time()
    ref = oracle_gp(kdesc, noise, X, y)
    print(f"oracle fit {time.time() - t0:.3f}s")
    Lr = ref.L
    print(f"L normwise {rel(np.tril(L), np.tril(Lr)):.3e}  elementwise max {np.nanmax(np.abs(np.tril(L) - np.tril(Lr))):.3e}"
          f"  upper NaN: {bool(np.all(np.isnan(L[np.triu_indices(n, 1)])))}")
    K = O.gram_lower(ref.kernel, X, noise)
    Kf = np.tril(K) + np.tril(K, -1).T
    Lt = np.tril(L)
    print(f"backward error |LL^T-K|/|K| = {rel(Lt @ Lt.T, Kf):.3e}")
    for name in ("predict", "predict_variance"):
        a = getattr(gp, name)(Xq)
        print(f"  [{name}] device {gp._h.last_device_ms():.3f} ms launches {gp._h.last_launch_count()}")
        b = getattr(ref, name)(Xq)
        print(f"{name}: maxrel {maxrel(a, b):.3e} normwise {rel(a, b):.3e}")
    m, v = gp.predict_mean_variance(Xq)
    mr, vr = ref.predict_mean_variance(Xq)
    print(f"predict_mean_variance: mean maxrel {maxrel(m, mr):.3e} var maxrel {maxrel(v, vr):.3e}")
    print(f"likelihood: gpu {gp.likelihood():.12f} oracle {ref.likelihood():.12f}")
    qs = min(q, 96)
    cg = gp.predict_covariance(Xq[:qs])
    cr = ref.predict_covariance(Xq[:qs])
    print(f"predict_covariance: normwise {rel(cg, cr):.3e} max abs {np.abs(cg - cr).max():.3e}")
    mvn = gp.sample_at(Xq[:qs])
    mean2, cov2 = ref.sample_at_params(Xq[:qs])
    Lc = np.linalg.cholesky(cov2)
    print(f"sample_at: mean maxrel {maxrel(mvn.mean(), mean2):.3e} chol(cov) normwise {rel(mvn.cholesky_covariance, Lc):.3e}")
    return gp, ref, X, y


def diag_gradient(gp, ref):
    section("lml gradient")
    g = gp.gradient_marginal_likelihood()
    print(f"  device {gp._h.last_device_ms():.3f} ms launches {gp._h.last_launch_count()}")
    gr = ref.gradient_marginal_likelihood(scaled=False)
    print("unscaled gpu   ", g)
    print("unscaled oracle", list(gr))
    print(f"unscaled maxrel {maxrel(g, gr, 1e-9):.3e}")
    if gp.kernel.is_scalable():
        s, g2 = gp.scaled_gradient_marginal_likelihood()
        sr, g2r = ref.gradient_marginal_likelihood(scaled=True)
        print(f"scaled: scale {s:.15g} vs {sr:.15g}; grads maxrel {maxrel(g2, g2r, 1e-9):.3e}")


def diag_add_samples():
    section("add_samples")
    d = 3
    for (n0, k) in [(100, 5), (128, 128), (300, 77), (640, 300)]:
        X, y = make_dataset(777, n0 + k, d)
        ls = float(np.sqrt(d / 6.0))
        gp = GaussianProcess(ZeroPrior(), SquaredExp(ls, 1.0), 0.1, None, X[:n0], y[:n0])
        gp.add_samples(X[n0:], y[n0:])
        ref = oracle_gp(([O.K_SQUARED_EXP], [ls, 1.0]), 0.1, X[:n0], y[:n0])
        ref.add_samples(X[n0:], y[n0:])
        L, Lr = gp.cholesky_factor(), ref.L
        Xq = make_inputs(5, 50, d)
        print(f"n0={n0} k={k}: L normwise {rel(np.tril(L), np.tril(Lr)):.3e} mean maxrel "
              f"{maxrel(gp.predict(Xq), ref.predict(Xq)):.3e} var maxrel "
              f"{maxrel(gp.predict_variance(Xq), ref.predict_variance(Xq)):.3e}")


def diag_kernels():
    section("all kernels (n=200, d=3): Gram via L, predict, gradient")
    n, d, q = 200, 3, 40
    X, y = make_dataset(31337, n, d)
    Xq = make_inputs(4242, q, d)
    cases = [
        (SquaredExp(0.7, 1.3), ([O.K_SQUARED_EXP], [0.7, 1.3])),
        (Exponential(0.9, 1.1), ([O.K_EXPONENTIAL], [0.9, 1.1])),
        (Matern1(0.8, 1.2), ([O.K_MATERN1], [0.8, 1.2])),
        (Matern2(0.8, -1.2), ([O.K_MATERN2], [0.8, -1.2])),
        (RationalQuadratic(1.5, 0.9), ([O.K_RATIONAL_QUADRATIC], [1.5, 0.9])),
        (Linear(0.5) + SquaredExp(0.7, 1.0), ([O.K_LINEAR, O.K_SQUARED_EXP, O.K_SUM], [0.5, 0.7, 1.0])),
        (Polynomial(0.5, 1.0, 2.0) * Matern2(1.0, 1.0), ([O.K_POLYNOMIAL, O.K_MATERN2, O.K_PROD], [0.5, 1.0, 2.0, 1.0, 1.0])),
        (HyperTan(0.1, 0.2) + Matern1(0.6, 2.0), ([O.K_HYPERTAN, O.K_MATERN1, O.K_SUM], [0.1, 0.2, 0.6, 2.0])),
        (Multiquadric(0.5) * SquaredExp(0.5, 1.0), ([O.K_MULTIQUADRIC, O.K_SQUARED_EXP, O.K_PROD], [0.5, 0.5, 1.0])),
    ]
    for kern, kdesc in cases:
        try:
            gp = GaussianProcess(ZeroPrior(), kern, 0.3, None, X, y)
            ref = oracle_gp(kdesc, 0.3, X, y)
            L, Lr = gp.cholesky_factor(), ref.L
            g = gp.gradient_marginal_likelihood()
            gr = ref.gradient_marginal_likelihood(scaled=False)
            print(f"{kern!r}: L {rel(np.tril(L), np.tril(Lr)):.2e} mean {maxrel(gp.predict(Xq), ref.predict(Xq)):.2e} "
                  f"var {maxrel(gp.predict_variance(Xq), ref.predict_variance(Xq)):.2e} lik "
                  f"{abs(gp.likelihood() - ref.likelihood()):.2e} grad {maxrel(g, gr, 1e-9):.2e}")
        except Exception as e:  # noqa: BLE001
            print(f"{kern!r}: FAILED {type(e).__name__}: {e}")


def diag_misc():
    section("misc: mean pair distance, failure path, epsilon")
    n, d = 500, 4
    X, y = make_dataset(5, n, d)
    h = N.Handle(0)
    h.check(N.lib().fgp_set_inputs(h.ptr, N.dptr(X), n, n, d))
    out = C.c_double()
    h.check(N.lib().fgp_mean_pair_distance(h.ptr, C.cast(C.byref(out), N._dp)))
    print(f"mean pair distance gpu {out.value:.15g} oracle {O.fit_bandwidth_mean(X):.15g}")
    # duplicate points + zero noise -> singular
    Xd = np.asfortranarray(np.vstack([X[:150], X[:150]]))
    yd = np.concatenate([y[:150], y[:150]])
    try:
        GaussianProcess(ZeroPrior(), SquaredExp(1.0, 1.0), 0.0, None, Xd, yd)
        print("singular fit: no failure reported (!)")
    except ArithmeticError as e:
        Lr, fail = O.make_cholesky_cov_matrix(O.KernelDesc.make([O.K_SQUARED_EXP], [1.0, 1.0]), Xd, 0.0, None)
        print(f"singular fit: gpu says: {e} | oracle fail col {fail - 1}")
    try:
        gp = GaussianProcess(ZeroPrior(), SquaredExp(1.0, 1.0), 0.0, 1e-6, Xd, yd)
        Lr, fail = O.make_cholesky_cov_matrix(O.KernelDesc.make([O.K_SQUARED_EXP], [1.0, 1.0]), Xd, 0.0, 1e-6)
        L = gp.cholesky_factor()
        print(f"epsilon fit: oracle fail={fail} finite={bool(np.all(np.isfinite(np.tril(L))))} "
              f"first 150 cols normwise {rel(np.tril(L)[:, :150], np.tril(Lr)[:, :150]):.3e}")
    except Exception as e:  # noqa: BLE001
        print(f"epsilon fit FAILED {type(e).__name__}: {e}")


def main():
    print(N.lib().fgp_version().decode())
    guarded(diag_gemm)
    ls = float(np.sqrt(8 / 6.0))
    res = []
    guarded(lambda: res.append(diag_fit_predict(300, 8, 100, SquaredExp(ls, 1.0), ([O.K_SQUARED_EXP], [ls, 1.0]), tag="sqexp")))
    if res:
        guarded(lambda: diag_gradient(res[0][0], res[0][1]))
    res2 = []
    guarded(lambda: res2.append(diag_fit_predict(1000, 16, 300, Matern2(1.6, 1.0), ([O.K_MATERN2], [1.6, 1.0]), tag="matern2")))
    if res2:
        guarded(lambda: diag_gradient(res2[0][0], res2[0][1]))
    guarded(diag_add_samples)
    guarded(diag_kernels)
    guarded(diag_misc)
    if "--big" in sys.argv:
        guarded(lambda: diag_fit_predict(4096, 8, 1024, SquaredExp(ls, 1.0), ([O.K_SQUARED_EXP], [ls, 1.0]), tag="config2"))
    section("timing (device ms): fit at growing n, d=8, RBF")
    for n in (1024, 4096, 8192, 16384):
        X, y = make_dataset(1, n, 8)
        gp = GaussianProcess(ZeroPrior(), SquaredExp(ls, 1.0), 0.1, None, X, y)
        gp._refit()
        ms = gp._h.last_device_ms()
        fl = n ** 3 / 3 + n * n * 8
        print(f"n={n}: refit {ms:.2f} ms -> {fl / ms * 1e-9:.2f} TFLOP/s, launches {gp._h.last_launch_count()}", flush=True)
        del gp


if __name__ == "__main__":
    main()
