"""CPU tests of the int8 slice arithmetic behind the tcgen05 updates (tests/ozaki_model.py restates the device code)."""
import numpy as np
import pytest

import ozaki_model as oz


def _operands(rng, m, n, k):
    A = rng.normal(size=(m, k)) * np.exp(2.0 * rng.normal(size=(m, 1))) * np.exp2(-rng.integers(0, 8, size=(m, k)))
    B = rng.normal(size=(n, k)) * np.exp(2.0 * rng.normal(size=(n, 1)))
    A[1, :] = 0.0
    A[2, 3] = 1e12
    return A, B


@pytest.mark.parametrize("S", [4, 5, 6, 7, 8])
def test_digits_reconstruct_the_rounded_entry(S):
    rng = np.random.default_rng(S)
    A, _ = _operands(rng, 64, 8, 256)
    planes, scale = oz.slice_rows(A, S)
    assert planes.dtype == np.int8 and np.all(np.abs(planes[0].astype(int)) <= 64)
    v = np.zeros(A.shape, dtype=np.int64)
    for p in range(S):
        v = v * 256 + planes[p].astype(np.int64)
    back = v.astype(np.longdouble) * (scale.astype(np.longdouble) * np.longdouble(2.0) ** (8 - 8 * S))[:, None]
    rowmax = np.abs(A).max(axis=1)[:, None] + 1e-300
    # every entry is kept to 8 S - 2 bits below the power of two above its row maximum
    assert np.max(np.abs(back - A) / rowmax) <= 2.0 ** -(8 * S - 2)


@pytest.mark.parametrize("S,k", [(4, 512), (5, 512), (6, 1024), (7, 1024), (8, 1024), (7, 4096)])
def test_slice_product_error_bound(S, k):
    rng = np.random.default_rng(10 * S + k)
    A, B = _operands(rng, 96, 80, k)
    sgn = np.where(rng.random(k) < 0.3, -1.0, 1.0)
    C, worst = oz.gemm(A, B, S, sgn)
    ref = (A * sgn).astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :] * k + 1e-300
    err = float(np.max(np.abs(C - ref) / scale))
    assert err <= max(2.0 ** -(8 * S - 6), 4e-16), err
    assert worst < 2 ** 31


def test_int32_group_sums_cannot_overflow_within_one_launch():
    """kOzakiMaxK = 8192: at most S products of magnitude <= 2^14 per k and group."""
    for S in range(4, 9):
        assert 8192 * S * 128 * 128 < 2 ** 31
    # the extreme digits really occur: a row of values just below a power of two gives top digit 64, -128 below it
    planes, _ = oz.slice_rows(np.array([[1.0 - 2.0 ** -40, -0.5 - 2.0 ** -9]]), 7)
    assert planes.max() <= 127 and planes.min() >= -128 and abs(int(planes[0, 0, 0])) == 64
