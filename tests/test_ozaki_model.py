"""CPU tests of the arithmetic behind the tcgen05 trailing update (oracle/ozaki_model.py restates csrc/ozaki.cu in numpy):
digit range, int32 headroom, and the error of the 36-product truncation against the exact product."""
import numpy as np
import pytest

from oracle import ozaki_model as OM


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_digits_reconstruct_the_panel_to_55_bits_of_the_row_maximum(seed):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((64, 96)) * np.exp2(rng.integers(-30, 30, size=(64, 1)).astype(np.float64))
    P[5] = 0.0                                  # an all-zero row (padding rows of the factor are)
    D, scale = OM.digits(P)
    assert np.abs(D).max() <= 65 and np.abs(D[1:]).max() <= 64
    X = sum(D[i].astype(object) * 128 ** (7 - i) for i in range(8))          # exact integers
    e, nz = OM.row_exponents(P)
    back = np.array([[float(X[r, k]) for k in range(P.shape[1])] for r in range(P.shape[0])]) * np.ldexp(1.0, (e - 55).astype(np.int32))[:, None]
    rowmax = np.abs(P).max(axis=1, keepdims=True)
    assert np.all(np.abs(back - P) <= np.ldexp(rowmax, -54) + 0.0)     # half a unit of 2^(e-55), e <= log2(max) + 1
    assert scale[5] == 0.0 and np.all(D[:, 5] == 0)


def test_group_sums_fit_int32_for_the_largest_contraction():
    # worst case by construction: all digits at their extreme, K = 512
    K = 512
    worst = 8 * K * 65 * 64
    assert worst < 2 ** 31
    rng = np.random.default_rng(7)
    P = rng.standard_normal((128, K))
    D, _ = OM.digits(P)
    G = OM.group_products(D, D)
    assert max(int(np.abs(g).max()) for g in G) < 2 ** 27


@pytest.mark.parametrize("scaled", [False, True])
def test_update_matches_the_exact_product_like_an_f64_gemm(scaled):
    rng = np.random.default_rng(11)
    M, K = 256, 512
    P = rng.standard_normal((M, K))
    if scaled:
        P *= np.exp2(rng.integers(-40, 40, size=(M, 1)).astype(np.float64))
    C = rng.standard_normal((M, M))
    got = OM.update(C, P)
    exact = (C.astype(np.longdouble) - P.astype(np.longdouble) @ P.T.astype(np.longdouble))
    rowmax = np.abs(P).max(axis=1)
    # dropped slice pairs (i + j > 7) and the 55-bit digit cut: < 2^-52 of rowmax * colmax per term; plus two roundings of C
    budget = np.outer(rowmax, rowmax) * K * 2.0 ** -52 + 2.0 * np.abs(np.asarray(exact, dtype=np.float64)) * 2.0 ** -52
    assert np.all(np.abs(np.asarray(got - exact, dtype=np.float64)) <= budget)
    # and it is at least as accurate as a plain float64 GEMM on the same data
    plain = C - P @ P.T
    err_model = np.abs(np.asarray(got - exact, dtype=np.float64)).max()
    err_plain = np.abs(np.asarray(plain - exact, dtype=np.float64)).max()
    assert err_model <= 2.0 * err_plain + 1e-300
