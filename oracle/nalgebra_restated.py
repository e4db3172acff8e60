"""Second, independent restatement of nalgebra 0.31.4's dense f64 routines that friedrich's hot path calls — TEST
INFRASTRUCTURE ONLY (same rule as friedrich_oracle.c: only tests/ may import it).

Why it exists: the C oracle is the only arbiter of the Cholesky factor at 1e-10 and the reference cannot be built here
("parity unpinned").  This file re-derives the same loops a second time, in a different language and a
different style (numpy column slices instead of pointer loops), from nalgebra's published algorithms, so that
tests/test_oracle_restatement.py can assert BIT equality between the two restatements on random SPD inputs: a
transcription slip in either (loop bound, operand order, where the rounding happens) shows up as a differing bit.

Call sites in the reference: Cholesky::new / new_with_substitute src/algebra/mod.rs:83,90 · solve_lower_triangular
src/gaussian_process/mod.rs:203,260-263,342-345 · Cholesky::solve_mut mod.rs:235,298,379 · Cholesky::inverse
src/gaussian_process/optimizer.rs:32,169 · insert_column src/algebra/mod.rs:124.

Rounding model: Rust never contracts a*b+c into an FMA, so every product and every sum rounds once (numpy array
arithmetic does exactly that); nalgebra's `dotx` keeps eight running sums over chunks of eight and folds them as
(0+4)+(1+5)+(2+6)+(3+7) before the sequential tail; `axcpy` with beta = 1 is y[i] = a*x[i] + y[i].
"""
from __future__ import annotations

import math

import numpy as np


def dotx(a: np.ndarray, b: np.ndarray) -> float:
    """base/blas.rs dotx, unit stride."""
    n = len(a)
    m = n - n % 8
    prod = a[:m] * b[:m]                       # each product rounded once
    acc = np.zeros(8)
    for row in prod.reshape(-1, 8):            # acc_k += a[i+k]*b[i+k], chunk after chunk
        acc = acc + row
    res = 0.0
    res += acc[0] + acc[4]
    res += acc[1] + acc[5]
    res += acc[2] + acc[6]
    res += acc[3] + acc[7]
    for i in range(m, n):
        res += a[i] * b[i]
    return float(res)


def cholesky_new(A: np.ndarray, substitute=None):
    """linalg/cholesky.rs Cholesky::new_internal: left-looking COLUMN algorithm on the lower triangle.
    Returns (L in place of a copy of A — strict upper untouched, 0) or (partial, j+1) when column j's pivot fails."""
    L = np.array(A, dtype=np.float64, order="F", copy=True)
    n = L.shape[0]
    for j in range(n):
        for k in range(j):
            factor = -L[j, k]
            L[j:, j] = factor * L[j:, k] + L[j:, j]          # axpy(factor, col_k[j..], 1)
        diag = L[j, j]
        if diag != 0.0 and diag >= 0.0:                       # !is_zero && try_sqrt succeeds (NaN fails the >=)
            denom = math.sqrt(diag)
        elif substitute is not None and substitute != 0.0 and substitute >= 0.0:
            denom = math.sqrt(substitute)
        else:
            return L, j + 1
        L[j, j] = denom
        L[j + 1:, j] = L[j + 1:, j] / denom
    return L, 0


def solve_lower_triangular(L: np.ndarray, B: np.ndarray):
    """linalg/solve.rs solve_lower_triangular_mut: per right-hand side, column-oriented forward substitution.
    Returns None when a diagonal entry is exactly zero (the reference then panics through `expect`)."""
    X = np.array(B, dtype=np.float64, order="F", copy=True).reshape(L.shape[0], -1)
    n = L.shape[0]
    for c in range(X.shape[1]):
        b = X[:, c]
        for i in range(n):
            if L[i, i] == 0.0:
                return None
            coeff = b[i] / L[i, i]
            b[i] = coeff
            b[i + 1:] = (-coeff) * L[i + 1:, i] + b[i + 1:]
    return X


def ad_solve_lower_triangular(L: np.ndarray, B: np.ndarray):
    """linalg/solve.rs ad_solve_lower_triangular_unchecked_mut: L^T x = b, backwards, dot-product form."""
    X = np.array(B, dtype=np.float64, order="F", copy=True).reshape(L.shape[0], -1)
    n = L.shape[0]
    for c in range(X.shape[1]):
        b = X[:, c]
        for i in range(n - 1, -1, -1):
            d = dotx(np.ascontiguousarray(L[i + 1:, i]), np.ascontiguousarray(b[i + 1:]))
            b[i] = (b[i] - d) / L[i, i]
    return X


def cholesky_solve(L, B):
    """Cholesky::solve_mut: forward then adjoint."""
    return ad_solve_lower_triangular(L, solve_lower_triangular(L, B))


def cholesky_inverse(L):
    """Cholesky::inverse: solve_mut on a dense identity."""
    return cholesky_solve(L, np.eye(L.shape[0]))


def insert_last_column(L: np.ndarray, col: np.ndarray):
    """Cholesky::insert_column(j = n, col) — the only form add_rows_cholesky_cov_matrix uses (algebra/mod.rs:108-125):
    new (n+1)^2 zeroed matrix, old factor copied, L11 r = col[..n], row n = r^T, diag = sqrt(col[n] - ||r||^2) UNCHECKED."""
    n = L.shape[0]
    out = np.zeros((n + 1, n + 1), order="F")
    out[:n, :n] = np.where(np.tril(np.ones((n, n), dtype=bool)), L, 0.0)
    r = solve_lower_triangular(L, col[:n]) if n else np.zeros((0, 1))
    r = r[:, 0]
    out[n, :n] = r
    v = col[n] - dotx(np.ascontiguousarray(r), np.ascontiguousarray(r))
    out[n, n] = math.sqrt(v) if v >= 0.0 else float("nan")
    return out
