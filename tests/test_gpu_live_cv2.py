"""GPU: the CUDA path (through the C-ABI) against the cv2 wheel itself on seeded random cases -- the same generator as
tests/test_oracle_live_cv2.py (random sizes, windows, levels, criteria, masks, border points), so oracle, CUDA path and the
dependency the reference calls (s1_lucaskanade_tracking.py:311,323,326,437) are compared on identical inputs.  Skipped (not
passed) where cv2 does not import."""
import numpy as np
import pytest

from parity import assert_lk_parity, as_corners, corner_overlap, CORNER_OVERLAP, ERR_TOL
import test_oracle_live_cv2 as L

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def test_gray_and_pyramid_live(ibt):
    rng = np.random.default_rng(1100)
    for h, w in L.SIZES:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(ibt.cvtColor(rgb, ibt.COLOR_BGR2GRAY), cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)), (h, w)
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        for code in (cv2.COLOR_BGR2GRAY, cv2.COLOR_RGB2GRAY, cv2.COLOR_BGRA2GRAY, cv2.COLOR_RGBA2GRAY):
            assert np.array_equal(ibt.cvtColor(rgba, code), cv2.cvtColor(rgba, code)), (h, w, code)
    for h, w in [s for s in L.SIZES if min(s) >= 3]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        win = (int(rng.choice([3, 9, 21, 31, 35])), int(rng.choice([3, 9, 21, 31, 35])))
        ml_req = int(rng.integers(0, 6))
        ml_cv, pyr_cv = cv2.buildOpticalFlowPyramid(a, win, ml_req, withDerivatives=True)
        ml, pyr = ibt.buildOpticalFlowPyramid(a, win, ml_req, True)
        assert ml == ml_cv, (h, w, win, ml_req)
        for l in range(ml + 1):
            assert np.array_equal(pyr[2 * l], pyr_cv[2 * l]), (h, w, l)
            assert np.array_equal(pyr[2 * l + 1], pyr_cv[2 * l + 1]), (h, w, l)


@pytest.mark.parametrize("seed", range(4))
def test_gftt_live(ibt, seed):
    """corner lists of cv2 for random sizes, parameters and masks: same count, >= 99 % same set, >= 99 % same rank"""
    rng = np.random.default_rng(300 + seed)                 # the seeds of the oracle's live test
    h, w = int(rng.integers(40, 260)), int(rng.integers(40, 330))
    img = L._texture(rng, h, w, smooth=int(rng.integers(1, 4)))
    mask = (rng.random((h, w)) > 0.3).astype(np.uint8) * 255
    mask[: h // 5] = 0
    for _ in range(4):
        gp = dict(maxCorners=int(rng.choice([0, 25, 400, 50000000])), qualityLevel=float(rng.choice([0.007, 0.01, 0.05, 0.2])),
                  minDistance=float(rng.choice([0, 1, 3.5, 10, 25])), blockSize=int(rng.choice([3, 5, 10])))
        for m in (None, mask):
            ref = as_corners(cv2.goodFeaturesToTrack(img, mask=m, **gp))
            got = as_corners(ibt.goodFeaturesToTrack(img, mask=m, **gp))
            assert got.shape == ref.shape, (seed, gp, m is not None, got.shape, ref.shape)
            assert corner_overlap(got, ref) >= CORNER_OVERLAP, (seed, gp, m is not None)
            if len(ref):
                assert np.mean(np.all(got == ref, axis=(1, 2))) >= 0.99, (seed, gp, m is not None)


@pytest.mark.parametrize("seed", range(4))
def test_lk_live(ibt, seed):
    """forward + backward LK on displaced frames against cv2: BASELINE's status / position criteria, err, FB decisions"""
    rng = np.random.default_rng(400 + seed)
    h, w = int(rng.integers(70, 260)), int(rng.integers(70, 330))
    f0 = L._texture(rng, h, w, smooth=2)
    dx, dy = rng.uniform(-2.5, 2.5, 2)
    f1 = L._shifted(f0, float(dx), float(dy))
    f1 = np.clip(f1.astype(np.int16) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    n = 160
    pts = np.stack([rng.uniform(-3, w + 3, n), rng.uniform(-3, h + 3, n)], 1).astype(np.float32)
    pts[:20] = np.round(pts[:20])
    more = np.stack([rng.uniform(0, w, 640), rng.uniform(0, h, 640)], 1).astype(np.float32)
    pts = np.concatenate([pts, more]).reshape(-1, 1, 2)      # 800 points: one flipped status is 0.125 %
    for _ in range(3):
        win = (int(rng.choice([5, 9, 15, 21, 31, 35])), int(rng.choice([5, 9, 15, 21, 31, 35])))
        lp = dict(winSize=win, maxLevel=int(rng.integers(0, 5)),
                  criteria=(int(rng.choice([1, 2, 3])), int(rng.integers(1, 31)), float(rng.choice([0.0, 0.01, 0.03, 0.3]))))
        r_p1, r_st, r_err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        p1, st, err = ibt.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        assert_lk_parity(p1, st, r_p1, r_st, "fwd %r" % (lp,))
        # err is the residual AT the returned position: compare it where the positions agree to 1e-4 px (a point whose last
        # Newton step lands on the other side of epsilon stops one iteration apart, inside the 0.01 px criterion, and its
        # residual moves with it: 0.003 px -> 0.007 in err on seed 400's (35, 15) window)
        same = np.abs(np.asarray(p1) - r_p1).reshape(-1, 2).max(1) <= 1e-4
        ok = (st.ravel() == 1) & (r_st.ravel() == 1) & same
        # (err is an integer sum of |J - I| in 1/32 grey levels over the window: allow 16 such units on small windows)
        assert np.abs(err.ravel()[ok] - r_err.ravel()[ok]).max(initial=0) <= ERR_TOL + 16.0 / (32 * win[0] * win[1]), lp
        assert np.mean(same[(st.ravel() == 1) & (r_st.ravel() == 1)]) >= 0.9, lp
        r_p0r, r_st0, _ = cv2.calcOpticalFlowPyrLK(f1, f0, r_p1, None, **lp)
        p0r, st0, _ = ibt.calcOpticalFlowPyrLK(f1, f0, r_p1, None, **lp)
        assert_lk_parity(p0r, st0, r_p0r, r_st0, "bwd %r" % (lp,))
        # the fused launch (s1:323-333 in one kernel) against cv2's two calls + numpy
        r = ibt.calcOpticalFlowPyrLK_FB(f0, f1, pts, fb_threshold=1.0, **lp)
        assert_lk_parity(r["p1"], r["st1"], r_p1, r_st, "fused fwd %r" % (lp,))
        d_ref = np.hypot(*(pts.reshape(-1, 2) - r_p0r.reshape(-1, 2)).T)
        both = (r["st0"].ravel() == 1) & (r_st0.ravel() == 1) & (r["st1"].ravel() == 1) & (r_st.ravel() == 1)
        clear = both & (np.abs(d_ref - 1.0) > 0.02)          # decisions within the position tolerance of the threshold may flip
        assert np.mean((d_ref < 1.0)[clear] == r["valid"][clear]) >= 0.99, lp


@pytest.mark.parametrize("seed", range(3))
def test_harris_live(ibt, seed):
    """the Harris branch of the detector (cv2's useHarrisDetector=True) against cv2: response maps and corner lists"""
    import harris_cases as HC
    rng = np.random.default_rng(900 + seed)                 # the seeds of the oracle's live test
    h, w = int(rng.integers(40, 260)), int(rng.integers(40, 330))
    img = L._texture(rng, h, w, smooth=int(rng.integers(1, 4)))
    mask = (rng.random((h, w)) > 0.3).astype(np.uint8) * 255
    for bs in (2, 3, 5, 10):
        k = float(rng.choice([0.0, 0.04, 0.1]))
        ref = cv2.cornerHarris(img, bs, 3, k)
        assert np.abs(ibt.cornerHarris(img, bs, 3, k) - ref).max() <= HC.MAP_TOL * np.abs(ref).max(), (bs, k)
        for q, md in ((0.01, 0), (0.05, 5), (0.01, 10)):
            for m in (None, mask):
                gp = dict(maxCorners=int(rng.choice([0, 300])), qualityLevel=q, minDistance=md, blockSize=bs,
                          useHarrisDetector=True, k=k)
                HC.check_lists(ibt.goodFeaturesToTrack(img, mask=m, **gp), cv2.goodFeaturesToTrack(img, mask=m, **gp),
                               (seed, gp, m is not None))


@pytest.mark.parametrize("seed", range(3))
def test_lk_multichannel_live(ibt, seed):
    """3-channel frames (cv2's signature takes them; the reference converts to gray first): window sums run over pixels and
    channels -- status, positions and err against cv2 for random windows, levels and criteria"""
    import multichannel_cases as MC
    rng = np.random.default_rng(700 + seed)
    h, w = int(rng.integers(70, 260)), int(rng.integers(70, 330))
    f0 = np.dstack([L._texture(rng, h, w) for _ in range(3)])
    dx, dy = rng.uniform(-2.5, 2.5, 2)
    f1 = np.dstack([L._shifted(f0[..., c], float(dx), float(dy)) for c in range(3)])
    f1 = np.clip(f1.astype(np.int16) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    pts = np.stack([rng.uniform(-3, w + 3, 500), rng.uniform(-3, h + 3, 500)], 1).astype(np.float32).reshape(-1, 1, 2)
    for _ in range(3):
        win = (int(rng.choice([5, 9, 15, 21, 31, 35])), int(rng.choice([5, 9, 15, 21, 31, 35])))
        lp = dict(winSize=win, maxLevel=int(rng.integers(0, 5)),
                  criteria=(3, int(rng.integers(1, 31)), float(rng.choice([0.0, 0.01, 0.03]))))
        r_p1, r_st, r_err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        p1, st, err = ibt.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        assert_lk_parity(p1, st, r_p1, r_st, "3-channel %r" % (lp,))
        MC.check_err(err, st, p1, r_err, r_st, r_p1, win, lp)
