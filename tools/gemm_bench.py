"""Isolated timing of the production GEMM kernel (gemm_nt_kernel) on device-resident data: SYRK-shaped launches of the
sizes the blocked Cholesky issues.  Prints one JSON line per shape.  Usage: python tools/gemm_bench.py [reps]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from friedrich_b200 import _native as N  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
lib = N.lib()
shapes = [
    # (M, N, K, lower)  trailing updates of the 512-column panel schedule at several remaining sizes
    (32768, 32768, 512, 1), (16384, 16384, 512, 1), (8192, 8192, 512, 1), (4096, 4096, 512, 1), (2048, 2048, 512, 1),
    (16384, 16384, 1024, 1), (16384, 16384, 256, 1), (16384, 16384, 128, 1),
    # panel-internal products: block-column update (N = 128, K <= 384) and the panel solve (K = 128)
    (16384, 128, 384, 0), (16384, 128, 128, 0), (4096, 128, 384, 0), (4096, 128, 128, 0),
    # multi-RHS solve step of predict (M = q)
    (1024, 16384, 128, 0), (1024, 8192, 128, 0),
]
for (M, Nn, K, lower) in shapes:
    ms, fl = C.c_double(0), C.c_double(0)
    rc = lib.fgp_dbg_gemm_bench(0, M, Nn, K, lower, 1, reps, C.byref(ms), C.byref(fl))
    assert rc == 0, rc
    print(json.dumps({"M": M, "N": Nn, "K": K, "lower": lower, "ms": ms.value, "tflops": fl.value / ms.value * 1e-9,
                      "frac_of_37.07": fl.value / ms.value * 1e-9 / 37.07}), flush=True)
