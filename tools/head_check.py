#!/usr/bin/env python
"""GPU diagnostic for csrc/potrf_head.cu: the head kernel alone on random SPD blocks of 1..4 tiles, against LAPACK.
Prints one JSON line per case: ||L - chol(A)||/||L||, ||W L - I||, the info word and the kernel's CUDA-event time."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from friedrich_b200 import _native as N  # noqa: E402


def run(nt, seed, cond=0.05, reps=5):
    n = 128 * nt
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n + 16))
    A0 = np.asfortranarray(M @ M.T / n + cond * np.eye(n))
    A = A0.copy(order="F")
    W = np.zeros((n, n), order="F")
    info = C.c_int(-1)
    ms = C.c_double(0.0)
    rc = N.lib().fgp_dbg_potrf_head(0, N.dptr(A), nt, N.dptr(W), 0, 0.0, C.byref(info), reps, C.cast(C.byref(ms), N._dp))
    L = np.tril(A)
    Lref = np.linalg.cholesky(A0)
    Wl = np.tril(W)
    out = {"nt": nt, "rc": rc, "info": info.value, "us": ms.value * 1e3,
           "L_rel": float(np.linalg.norm(L - Lref) / np.linalg.norm(Lref)),
           "LLt_rel": float(np.linalg.norm(L @ L.T - A0) / np.linalg.norm(A0)),
           "WL_minus_I": float(np.abs(Wl @ Lref - np.eye(n)).max()),
           "W_upper_max": float(np.abs(np.triu(W, 1)).max())}
    # per-tile breakdown when something is off
    if not (out["L_rel"] < 1e-12 and out["WL_minus_I"] < 1e-9):
        tiles = {}
        for i in range(nt):
            for j in range(i + 1):
                bl = np.s_[128 * i:128 * (i + 1), 128 * j:128 * (j + 1)]
                tiles[f"L{i}{j}"] = float(np.abs(L[bl] - Lref[bl]).max())
                tiles[f"W{i}{j}"] = float(np.abs(Wl[bl] - np.linalg.inv(Lref)[bl]).max())
        out["tiles"] = tiles
    return out


if __name__ == "__main__":
    for nt in (1, 2, 3, 4):
        print(json.dumps(run(nt, 100 + nt)), flush=True)
    print(json.dumps(run(4, 7, cond=1e-3)), flush=True)
