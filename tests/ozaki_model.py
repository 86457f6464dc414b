"""NumPy model of the int8 slice arithmetic of cheq_b200/csrc/kernels_ozaki.cu (test infrastructure, CPU only).

The device code writes every operand entry as  x = 2^(e - B) * sum_p d_p 256^(S - p),  B = 8 S - 2, with one exponent e
per ROW (rowmax < 2^e) and S balanced base-256 digits (oz_slice_kernel), multiplies the digit planes exactly in int32
(tcgen05.mma.kind::i8, groups g = p + q kept for g <= S + 1) and recombines  C = 2^(ea - 6) 2^(eb - 6) sum_g 256^(2 - g) G_g
by Horner in fp64 (oz_gemm_kernel epilogue).  This module restates those three steps with the same integer operations.
"""
import numpy as np


def row_exponents(X):
    """e with rowmax < 2^e, from the exponent field of the largest magnitude (oz_exponent); 0 for an all-zero row."""
    mx = np.abs(X).max(axis=1)
    bits = mx.view(np.uint64) if mx.dtype == np.float64 else np.asarray(mx, dtype=np.float64).view(np.uint64)
    E = ((bits >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64)
    e = np.maximum(E - 1022, -900)
    return np.where(bits == 0, 0, e)


def slice_rows(X, S, signs=None):
    """(planes [S][rows][K] int8, scale [rows]): digits exactly as oz_slice_kernel produces them."""
    X = np.asarray(X, dtype=np.float64)
    if signs is not None:
        X = X * np.asarray(signs, dtype=np.float64)[None, :]
    e = row_exponents(np.abs(X))
    B = 8 * S - 2
    v = np.rint(X * np.exp2((B - e).astype(np.float64))[:, None]).astype(np.int64)   # __double2ll_rn
    planes = np.zeros((S,) + X.shape, dtype=np.int8)
    for p in range(S - 1, 0, -1):
        d = ((v + 128) & 255) - 128            # balanced digit in [-128, 127]
        v = (v - d) >> 8                       # exact: v - d is a multiple of 256
        planes[p] = d.astype(np.int8)
    assert np.all(np.abs(v) <= 64), "top digit must fit [-64, 64]"
    planes[0] = v.astype(np.int8)
    return planes, np.exp2((e - 6).astype(np.float64))


def gemm(A, Bm, S, signs=None):
    """A S B^T through the slice products (int64 here; the device accumulates in int32, see max_group_sum)."""
    pa, sa = slice_rows(A, S)
    pb, sb = slice_rows(Bm, S, signs)
    acc = np.zeros((A.shape[0], Bm.shape[0]))
    worst = 0
    for g in range(S + 1, 1, -1):              # Horner from the least significant group kept
        G = np.zeros(acc.shape, dtype=np.int64)
        for p in range(1, S + 1):
            q = g - p
            if 1 <= q <= S:
                G += pa[p - 1].astype(np.int64) @ pb[q - 1].astype(np.int64).T
        worst = max(worst, int(np.abs(G).max()))
        acc = acc * (1.0 / 256.0) + G.astype(np.float64)
    return acc * sa[:, None] * sb[None, :], worst
