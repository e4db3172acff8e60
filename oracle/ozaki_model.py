"""numpy restatement of the exact-integer trailing update of csrc/ozaki.cu (test infrastructure only: imported by tests/ alone).

The device path cuts every row of the panel P into 8 balanced base-128 digits under a per-row power-of-two scale, forms the
36 digit-slice products with i + j <= 7 in int32 on the tcgen05 tensor cores, combines the groups g = i + j exactly and
adds the result to C in two steps (groups 0..3, then 4..7).  Every step is exact except the two additions into C, so this
model — the same steps in numpy int64 / float64 — must reproduce the kernel's output BIT FOR BIT."""
import numpy as np

SLICES = 8


def row_exponents(P):
    """e_r with max_k |P[r, k]| < 2^e_r (0 for an all-zero row)."""
    m = np.abs(P).max(axis=1)
    _, e = np.frexp(m)          # m = f * 2^e, 0.5 <= f < 1
    return np.where(m > 0, e, 0).astype(np.int64), m > 0


def digits(P):
    """(D, scale): D[i] int64 digit slices (|d| <= 65), scale[r] = 2^(e_r - 30) (0 for zero rows); P ~ 2^(e-55) sum_i D[i] 128^(7-i)."""
    e, nz = row_exponents(P)
    up = np.where(nz, np.ldexp(1.0, (55 - e).astype(np.int32)), 0.0)
    X = np.rint(P * up[:, None]).astype(np.int64)
    D = np.zeros((SLICES,) + P.shape, dtype=np.int64)
    for i in range(SLICES - 1, 0, -1):
        d = ((X & 127) ^ 64) - 64
        D[i] = d
        X = (X - d) >> 7
    D[0] = X
    scale = np.where(nz, np.ldexp(1.0, (e - 30).astype(np.int32)), 0.0)
    return D, scale


def group_products(DA, DB):
    """G[g] = sum_{i+j=g} DA[i] DB[j]^T for g = 0..7 (int64; the device accumulates each in int32)."""
    return [sum(DA[i] @ DB[g - i].T for i in range(g + 1)) for g in range(SLICES)]


def update(C, PA, PB=None):
    """C - PA PB^T as the device computes it (full matrix; the caller masks the tiles a launch covers)."""
    PB = PA if PB is None else PB
    DA, sa = digits(PA)
    DB, sb = digits(PB)
    G = group_products(DA, DB)
    assert max(int(np.abs(g).max()) for g in G) < 2 ** 31
    v0 = ((G[0] * 128 + G[1]) * 128 + G[2]) * 128 + G[3]
    v1 = ((G[4] * 128 + G[5]) * 128 + G[6]) * 128 + G[7]
    f0 = (sa * -134217728.0)[:, None] * sb[None, :]      # -2^27 * 2^(e_r-30) * 2^(e_c-30) = -2^(e_r+e_c-110) * 128^7 * 128^4
    f1 = (sa * -0.5)[:, None] * sb[None, :]
    out = C + v0.astype(np.float64) * f0
    return out + v1.astype(np.float64) * f1


def launch_mask(M, lower, row_skip):
    """elements a lower-mode launch with the first row_skip tile rows left out touches (diagonal tiles: r >= c only)"""
    T = M // 128
    mask = np.zeros((M, M), bool)
    for ti in range(T):
        for tj in range(T):
            if (not lower) or (tj <= ti and ti >= row_skip):
                blk = np.ones((128, 128), bool)
                if lower and ti == tj:
                    blk = np.tril(blk)
                mask[ti * 128:(ti + 1) * 128, tj * 128:(tj + 1) * 128] = blk
    return mask
