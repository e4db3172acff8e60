"""CPU: the C oracle's nalgebra restatement against a second, independently written one (oracle/nalgebra_restated.py):
BIT equality on random SPD inputs.  Two restatements of the same published loops that agree to the last bit rule out a
transcription slip in either (the reference itself cannot run here: SURVEY.md §8c, "parity unpinned")."""
import numpy as np
import pytest

from oracle import nalgebra_restated as R
from oracle import oracle as O


def _spd(n, seed, cond_boost=0.0):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n + 3))
    A = M @ M.T / n + (0.05 + cond_boost) * np.eye(n)
    return np.asfortranarray(A)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 16, 33, 64, 127, 150])
def test_cholesky_new_bit_equal(n):
    A = _spd(n, 100 + n)
    Lc = A.copy(order="F")
    assert O.cholesky_inplace(Lc) == 0
    Lp, fail = R.cholesky_new(A)
    assert fail == 0
    assert np.array_equal(np.tril(Lc), np.tril(Lp))
    # the strict upper triangle is never touched by either (nalgebra reads/writes the lower triangle only)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(Lc[iu], A[iu]) and np.array_equal(Lp[iu], A[iu])


def test_cholesky_failure_column_and_substitute_bit_equal():
    n = 40
    A = _spd(n, 7)
    A[25:, 25:] = A[:15, :15]          # rows 25.. duplicate rows 0..14 of a different block: not positive definite
    A = np.asfortranarray(np.tril(A) + np.tril(A, -1).T)
    Lc = A.copy(order="F")
    fail_c = O.cholesky_inplace(Lc)
    Lp, fail_p = R.cholesky_new(A)
    assert fail_c == fail_p and fail_c > 0
    j = fail_c - 1
    assert np.array_equal(np.tril(Lc)[:, :j], np.tril(Lp)[:, :j])   # columns before the failure are final and equal
    # exactly singular by construction: duplicated row/column -> zero pivot at the duplicate, substitute takes over
    B = _spd(20, 9)
    B2 = np.zeros((21, 21), order="F")
    B2[:20, :20] = B
    B2[20, :20] = B[5, :]
    B2[:20, 20] = B[:, 5]
    B2[20, 20] = B[5, 5]
    Lc = B2.copy(order="F")
    fc = O.cholesky_inplace(Lc)
    Lp, fp = R.cholesky_new(B2)
    assert fc == fp
    Lc = B2.copy(order="F")
    fc = O.cholesky_inplace(Lc, substitute=1e-6)
    Lp, fp = R.cholesky_new(B2, substitute=1e-6)
    assert fc == fp == 0 and np.array_equal(np.tril(Lc), np.tril(Lp))


@pytest.mark.parametrize("n,q", [(5, 1), (8, 3), (17, 4), (64, 5), (131, 2)])
def test_solves_and_inverse_bit_equal(n, q):
    A = _spd(n, 300 + n)
    L = A.copy(order="F")
    assert O.cholesky_inplace(L) == 0
    B = np.asfortranarray(np.random.default_rng(n).standard_normal((n, q)))
    Xc, ok = O.solve_lower(L, B)
    assert ok
    assert np.array_equal(Xc, R.solve_lower_triangular(L, B))
    assert np.array_equal(O.chol_solve(L, B), R.cholesky_solve(L, B))
    if n <= 64:
        assert np.array_equal(O.chol_inverse(L), R.cholesky_inverse(L))


def test_dotx_accumulator_order_matters_and_matches():
    """The 8-accumulator fold is observable: a naive sequential sum differs in the last bits, the restated dotx does not
    (checked through the adjoint solve, the only consumer of dotx on this path besides the column norms)."""
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(1003), rng.standard_normal(1003)
    naive = 0.0
    for x, y in zip(a, b):
        naive += x * y
    # a lower-triangular system whose adjoint solve at row 0 is exactly one dotx over 1003 elements
    n = 1004
    L = np.asfortranarray(np.eye(n))
    L[1:, 0] = a
    rhs = np.zeros((n, 1), order="F")
    rhs[1:, 0] = b
    from ctypes import c_int64  # noqa: F401
    lib = O.lib()
    X = rhs.copy(order="F")
    lib.fo_ad_solve_lower(O._p(L), n, n, O._p(X), n, 1)
    assert X[0, 0] == -R.dotx(a, b)
    assert abs(naive - R.dotx(a, b)) < 1e-9  # same value up to rounding ...
    # ... and the restated order is the one the C oracle uses (bit equality above); the naive order is a different sum
    # in general, which is what makes the equality above a meaningful check


@pytest.mark.parametrize("n0,k", [(1, 3), (9, 8), (40, 17)])
def test_insert_column_bit_equal(n0, k):
    from friedrich_b200.synthetic import make_dataset
    d = 3
    X, y = make_dataset(4242 + n0, n0 + k, d)
    kd = O.KernelDesc.make([O.K_MATERN2], [0.9, 1.3])
    noise = 0.2
    gp = O.OracleGaussianProcess(O.ZeroPrior(), kd, noise, None, X[:n0], y[:n0])
    gp.add_samples(X[n0:], y[n0:])
    # second restatement: the same Gram columns (kernel(x_t, x_new), + noise^2 on the last entry: algebra/mod.rs:115-121)
    # pushed through insert_last_column one by one
    K0 = O.gram_lower(kd, X[:n0], noise)
    L, fail = R.cholesky_new(K0)
    assert fail == 0
    for i in range(k):
        j = n0 + i
        col = O.make_covariance_matrix(kd, X[:j + 1], X[j:j + 1])[:, 0].copy()
        col[j] += noise * noise
        L = R.insert_last_column(L, col)
    assert np.array_equal(np.tril(gp.L), np.tril(L))
