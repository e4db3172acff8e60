"""Device timings of the other §8 rows at their BASELINE.json config sizes (one JSON line each; profiles/rows_*.jsonl):
  C3  Matern-5/2 n=16384 d=16: one scaled LML-gradient evaluation (a12+a13) and one ADAM iteration = refit + gradient (a14)
  C5  add_samples: n=16384 base + 1024 new points (a11), against a from-scratch fit of the 17408 rows
  a15 likelihood, a10 predict_covariance q=512, a16 mean pair distance
Times are CUDA-event device times of the C-ABI calls (fgp_last_device_ms), after one warm-up call (buffers allocated, capacity
grown: add_samples is timed on a model whose capacity already holds the new rows — the base model is fitted on n - k rows, a
first add_samples of k rows grows the capacity, the SECOND one is exactly "n base + k new").
Usage: python tools/bench_rows.py [n] [d]"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import friedrich_b200 as F  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
PEAK = 37.07


def emit(row, ms, flops=None, **kw):
    line = {"row": row, "n": n, "d": d, "device_ms": ms}
    if flops:
        line["algorithmic_flops"] = flops
        line["tflops"] = flops / (ms * 1e-3) * 1e-12
        line["frac_of_fp64_peak"] = line["tflops"] / PEAK
    line.update(kw)
    print(json.dumps(line), flush=True)


X, y = make_dataset(0x5EED0003, n, d)
ls = math.sqrt(d / 6.0)

# ---- C3: Matern2 LML gradient -----------------------------------------------------------------------------------------
gp = F.GaussianProcess(F.ZeroPrior(), F.Matern2(ls, 1.0), 0.1, None, X, y)
gp._refit()
emit("fit (Matern2)", gp._h.last_device_ms(), n ** 3 / 3.0 + float(n) * n * d)
for rep in range(2):
    t0 = time.perf_counter()
    scale, grads = gp.scaled_gradient_marginal_likelihood()
    wall = (time.perf_counter() - t0) * 1e3
    ms = gp._h.last_device_ms()
# algorithmic: U = L^-T (n^3/3) + K^-1 = U U^T lower (n^3/3) + gradient Gram reductions 2 n^2 d
emit("scaled LML gradient (a12+a13)", ms, 2.0 * n ** 3 / 3.0 + 2.0 * n * n * d, wall_ms=wall, scale=float(scale),
     grads=[float(g) for g in grads])
gp._refit()
emit("ADAM iteration = refit + gradient (a14)", gp._h.last_device_ms() + ms, n ** 3 + 3.0 * n * n * d)
for rep in range(2):
    lik = gp.likelihood()
emit("likelihood (a15)", gp._h.last_device_ms(), None, value=float(lik))
del gp

# ---- C5: add_samples --------------------------------------------------------------------------------------------------
k = 1024
Xa, ya = make_dataset(0x5EED0005, n + k, d)
gp = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(ls, 1.0), 0.1, None, Xa[:n - k], ya[:n - k])
gp.add_samples(Xa[n - k:n], ya[n - k:n])
emit("add_samples k=1024 onto n-k rows, first call (grows the capacity: allocations + copy of L inside)", gp._h.last_device_ms())
gp.add_samples(Xa[n:], ya[n:])
ms_add = gp._h.last_device_ms()
emit("add_samples k=1024 (a11)", ms_add, float(n) * n * k + float(n) * k * k + k ** 3 / 3.0 + 2.0 * n * k * d)
La = np.tril(gp.cholesky_factor()[n:, :][:, : n + k])  # the new block rows only (the first n rows are untouched)
gp2 = F.GaussianProcess(F.ZeroPrior(), F.SquaredExp(ls, 1.0), 0.1, None, Xa, ya)
gp2._refit()
emit("from-scratch fit of n+k rows", gp2._h.last_device_ms(), (n + k) ** 3 / 3.0 + float(n + k) ** 2 * d)
Lb = np.tril(gp2.cholesky_factor()[n:, :][:, : n + k])
print(json.dumps({"row": "add_samples vs from-scratch: new block rows of L", "frob_rel": float(np.linalg.norm(La - Lb) /
                                                                                               np.linalg.norm(Lb))}), flush=True)
Xq = make_inputs(77, 512, d)
for rep in range(2):
    cov = gp2.predict_covariance(Xq)
emit("predict_covariance q=512 (a10)", gp2._h.last_device_ms(), float(n + k) ** 2 * 512 + float(n + k) * 512 * 512)
