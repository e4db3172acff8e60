"""Profiling driver (run under ncu on the GPU box): one fit at the given size through the C-ABI, optionally followed by
one predict_mean_variance.  Usage: python tools/prof_fit.py [n] [d] [q]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402
from friedrich_b200.kernels import SquaredExp  # noqa: E402
from friedrich_b200.synthetic import make_dataset, make_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
q = int(sys.argv[3]) if len(sys.argv) > 3 else 0
X, y = make_dataset(0x5EED0001, n, d)
h = N.Handle(0)
kd = SquaredExp(math.sqrt(d / 6.0), 1.0).device_desc()
lib = N.lib()
h.check(lib.fgp_fit(h.ptr, N.dptr(X), n, n, d, N.dptr(y), C.byref(kd), 0.1, 0, 0.0))
print("fit device ms", h.last_device_ms(), "launches", h.last_launch_count())
if q:
    import numpy as np
    Xq = make_inputs(0x5EED0002, q, d)
    mean, var = np.zeros(q), np.zeros(q)
    h.check(lib.fgp_predict_mean_var(h.ptr, C.byref(kd), N.dptr(Xq), q, q, N.dptr(mean), N.dptr(var)))
    print("predict device ms", h.last_device_ms(), "launches", h.last_launch_count())
