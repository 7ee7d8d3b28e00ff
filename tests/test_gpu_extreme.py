"""GPU: the CUDA path (through the C-ABI) against the same cv2 known answers at parameter extremes and awkward shapes."""
import numpy as np
import pytest

import extreme_cases as X

pytestmark = pytest.mark.gpu


def test_cuda_lk_extreme(ibt, golden):
    X.check_lk_extreme(ibt, golden("kat_extreme.npz"))


def test_cuda_gftt_extreme(ibt, golden):
    X.check_gftt_extreme(ibt, golden("kat_extreme.npz"))


def test_cuda_tiny(ibt, golden):
    X.check_tiny(ibt, golden("kat_extreme.npz"))


def test_cuda_vs_oracle_random_shapes(ibt, oracle):
    """Randomised shapes / parameters, CUDA vs oracle (integer stages bit-exact)."""
    rng = np.random.default_rng(123)
    for trial in range(12):
        h, w = int(rng.integers(8, 300)), int(rng.integers(8, 400))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(ibt.cvtColor(rgb, 6), oracle.cvtColor(rgb)), (h, w)
        assert np.array_equal(ibt.pyrDown(a), oracle.pyrDown(a)), (h, w)
        win = (int(rng.integers(3, 40)), int(rng.integers(3, 40)))
        ml, pyr = ibt.buildOpticalFlowPyramid(a, win, 4, True)
        ml_o, pyr_o = oracle.buildOpticalFlowPyramid(a, win, 4, True)
        assert ml == ml_o
        for x, y in zip(pyr, pyr_o):
            assert np.array_equal(x, y), (h, w, win)
        bs = int(rng.integers(2, 12))          # blockSize 1 is degenerate: lambda_min of a rank-1 tensor is rounding noise
        e, eo = ibt.cornerMinEigenVal(a, bs), oracle.cornerMinEigenVal(a, bs)
        assert np.abs(e - eo).max() <= 2e-4 * np.abs(eo).max(), (h, w, bs)
