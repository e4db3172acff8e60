"""Latency path of predict (csrc/fgp_api.cu predict_small): device time of fgp_predict_mean_var for a handful of queries at
n = 16384, d = 16 (one multi-right-hand-side wavefront launch up to q = 16, the tensor-pipe path above).
    python tools/predict_small_times.py [n] [d]"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200.kernels import SquaredExp  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
X, y = make_dataset(0x5EED0001, n, d)
h = N.Handle(0)
kd = SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
lib = N.lib()
h.check(lib.fgp_fit(h.ptr, N.dptr(N.fcol(X)), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
out = {"n": n, "d": d, "fit_ms": h.last_device_ms(), "predict_mean_var_ms": {}}
for q in (1, 2, 4, 8, 12, 16, 17, 32):
    Xq = N.fcol(make_inputs(0x5EED0100 + q, q, d))
    mean, var = np.zeros(q), np.zeros(q)
    best = 1e9
    for _ in range(4):
        h.check(lib.fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(Xq), q, q, N.dptr(mean), N.dptr(var)))
        best = min(best, h.last_device_ms())
    out["predict_mean_var_ms"][str(q)] = round(best, 4)
print(json.dumps(out))
