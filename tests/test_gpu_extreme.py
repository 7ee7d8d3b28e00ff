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


def test_pyramid_cp_async_path_without_tma(tmp_path):
    """The cp.async staging paths of K1 and K3 (used when no tensor map can be made) stay exact: run in a subprocess
    with IBT_NO_TMA=1 against the oracle."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from iceberg_tracking_code_b200 import cv\n"
        "from oracle import oracle as orc\n"
        "rng = np.random.default_rng(3)\n"
        "for hw in [(1080, 1920), (333, 640), (64, 48), (37, 1024)]:\n"
        "    a = rng.integers(0, 256, hw, dtype=np.uint8)\n"
        "    ml, p = cv.buildOpticalFlowPyramid(a, (21, 21), 4, True)\n"
        "    mlo, po = orc.buildOpticalFlowPyramid(a, (21, 21), 4, True)\n"
        "    assert ml == mlo and all(np.array_equal(x, y) for x, y in zip(p, po)), hw\n"
        "g = dict(np.load(%r))\n"
        "for lp in (dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01)), dict(winSize=(35, 35), maxLevel=4, criteria=(3, 25, 0.03))):\n"
        "    p1, st, err = cv.calcOpticalFlowPyrLK(g['f0'], g['f1'], g['lk_pts'], None, **lp)\n"
        "    po, so, eo = orc.calcOpticalFlowPyrLK(g['f0'], g['f1'], g['lk_pts'], None, **lp)\n"
        "    assert np.mean(st == so) >= 0.995 and np.abs(p1 - po).reshape(-1, 2).max(1)[(st == 1).ravel() & (so == 1).ravel()].max() <= 0.01\n"
        "print('cp.async path ok')\n" % (root, os.path.join(root, "tests", "golden", "kat_texture.npz")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, IBT_NO_TMA="1"))
    assert r.returncode == 0 and "cp.async path ok" in r.stdout, r.stdout + r.stderr
