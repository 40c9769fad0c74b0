"""tcgen05 GEMM kernel vs float64 numpy, every operand-major / tile / cluster-split variant (GPU only)."""
import numpy as np
import pytest
from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = []
for a_mn, b_mn in ((0, 0), (0, 1), (1, 1)):
    for bn in (64, 128):
        for splits in (1, 2, 4, 8):
            if (bn // splits) % 16:
                continue
            CASES.append((a_mn, b_mn, bn, splits))


@pytest.mark.parametrize("a_mn,b_mn,bn,splits", CASES)
def test_gemm_variants(pkg, a_mn, b_mn, bn, splits):
    from se_ml_b200.bp_gpu import debug_gemm
    rng = np.random.RandomState(1000 + a_mn * 100 + b_mn * 10 + bn + splits)
    I, J, R = 256, 2 * bn + 64 * (bn == 64), 64 * 9
    A = rng.randn(I, R).astype(np.float32)
    B = rng.randn(J, R).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    Am = np.ascontiguousarray(A.T) if a_mn else A
    Bm = np.ascontiguousarray(B.T) if b_mn else B
    got = debug_gemm(a_mn, b_mn, I, J, R, bn, splits, Am, Bm)
    err = rel_err(got, ref)
    assert err < 3e-5, "rel err %g (a_mn=%d b_mn=%d bn=%d splits=%d)" % (err, a_mn, b_mn, bn, splits)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
def test_gemm_ragged(pkg, a_mn, b_mn):
    """extents that are not tile multiples: zero padding must not leak into the result"""
    from se_ml_b200.bp_gpu import debug_gemm
    rng = np.random.RandomState(7)
    I, J, R = 130, 257, 1799
    A = rng.randn(I, R).astype(np.float32)
    B = rng.randn(J, R).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    Am = np.ascontiguousarray(A.T) if a_mn else A
    Bm = np.ascontiguousarray(B.T) if b_mn else B
    got = debug_gemm(a_mn, b_mn, I, J, R, 64, 4, Am, Bm)
    assert rel_err(got, ref) < 3e-5


def test_gemm_one_hot_exact(pkg):
    """identity-like operands: every product is exact, so the result must be bit-exact (catches swizzle /
    descriptor mistakes that random data could hide behind the tolerance)"""
    from se_ml_b200.bp_gpu import debug_gemm
    I, J, R = 128, 128, 128
    A = np.zeros((I, R), np.float32); B = np.zeros((J, R), np.float32)
    for i in range(I):
        A[i, (i * 7) % R] = 1.0 + i
    for j in range(J):
        B[j, (j * 11 + 3) % R] = 2.0 + j
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 1)):
        Am = np.ascontiguousarray(A.T) if a_mn else A
        Bm = np.ascontiguousarray(B.T) if b_mn else B
        got = debug_gemm(a_mn, b_mn, I, J, R, 128, 1, Am, Bm)
        assert np.array_equal(got.astype(np.float64), ref), (a_mn, b_mn)
