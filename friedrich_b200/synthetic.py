"""Deterministic synthetic workloads shared by tests and bench.py.

X ~ U[0,1)^{n x d} from splitmix64(seed) -> (u >> 11) * 2^-53, filled in column-major order;
y_i = sum_j sin(2 pi x_ij) / sqrt(d) + 0.1 z_i with z from Box-Muller on the same stream.
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_uniform(seed: int, count: int, skip: int = 0) -> np.ndarray:
    """`count` doubles in [0,1) from the splitmix64 stream started at `seed`, skipping `skip` outputs."""
    with np.errstate(over="ignore"):
        idx = np.arange(skip + 1, skip + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_inputs(seed: int, n: int, d: int) -> np.ndarray:
    """n x d column-major (Fortran order) matrix of U[0,1) samples."""
    return np.asfortranarray(splitmix64_uniform(seed, n * d).reshape((n, d), order="F"))


def make_dataset(seed: int, n: int, d: int):
    """(X, y) as described in the module docstring."""
    X = make_inputs(seed, n, d)
    u = splitmix64_uniform(seed, 2 * n, skip=n * d)
    u1 = np.maximum(u[:n], 1e-300)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u[n:])
    y = np.sin(2.0 * np.pi * X).sum(axis=1) / np.sqrt(d) + 0.1 * z
    return X, np.ascontiguousarray(y)


def config_seed(config_number: int) -> int:
    return 0x5EED0000 + config_number
