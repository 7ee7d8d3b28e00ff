"""Live pin of the CPU oracle against the cv2 wheel itself (the dependency that holds the arithmetic of
s1_lucaskanade_tracking.py:311,323,326,437): random sizes (odd, tiny, non-multiples of the SIMD widths), random parameter
sets, masks and displaced frames drawn from a seeded generator, every call evaluated by BOTH cv2 and the oracle.  The
golden files (test_oracle_golden.py) hold a fixed set of such answers so that the pin also works where cv2 is absent; this
file widens it wherever the wheel imports (build container and GPU box).  CPU only, a few seconds."""
import numpy as np
import pytest

from parity import assert_lk_parity, as_corners, ERR_TOL

cv2 = pytest.importorskip("cv2")


def _texture(rng, h, w, smooth=2):
    """band-limited random texture with corners (u8), the kind of image LK can lock on"""
    a = rng.random((h + 8, w + 8)).astype(np.float32)
    for _ in range(smooth):
        a = (a + np.roll(a, 1, 0) + np.roll(a, 1, 1) + np.roll(np.roll(a, 1, 0), 1, 1)) * 0.25
    a = a[4:4 + h, 4:4 + w]
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-9)
    blocks = (rng.random((h // 7 + 2, w // 7 + 2)) > 0.5).astype(np.float32)
    blocks = np.kron(blocks, np.ones((7, 7), np.float32))[:h, :w]
    return np.clip(a * 170 + blocks * 70 + rng.normal(0, 2, (h, w)), 0, 255).astype(np.uint8)


def _shifted(img, dx, dy):
    """img displaced by a sub-pixel amount (bilinear), borders replicated"""
    h, w = img.shape
    m = np.float32([[1, 0, dx], [0, 1, dy]])
    return cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)


SIZES = [(1, 1), (1, 7), (2, 2), (3, 3), (5, 64), (17, 33), (31, 31), (64, 48), (97, 131), (128, 257), (240, 321)]


@pytest.mark.parametrize("seed", range(4))
def test_cvtcolor_live(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    for h, w in SIZES:
        for cn in (3, 4):
            rgb = rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
            code = cv2.COLOR_BGR2GRAY if cn == 3 else cv2.COLOR_BGRA2GRAY
            assert np.array_equal(oracle.cvtColor(rgb), cv2.cvtColor(rgb, code)), (h, w, cn)
    # saturated / constant channels: the rounding term decides the last bit
    for v in (0, 1, 127, 128, 254, 255):
        rgb = np.full((9, 21, 3), v, np.uint8)
        rgb[..., 1] = 255 - v
        assert np.array_equal(oracle.cvtColor(rgb), cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)), v


@pytest.mark.parametrize("seed", range(3))
def test_pyrdown_and_pyramid_live(oracle, seed):
    rng = np.random.default_rng(200 + seed)
    for h, w in SIZES:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if h >= 2 or w >= 2 or (h, w) == (1, 1):
            assert np.array_equal(oracle.pyrDown(a), cv2.pyrDown(a)), (h, w)
    for h, w in [(s[0], s[1]) for s in SIZES if min(s) >= 3]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        win = (int(rng.choice([3, 9, 21, 31, 35])), int(rng.choice([3, 9, 21, 31, 35])))
        ml_req = int(rng.integers(0, 6))
        ml_cv, pyr_cv = cv2.buildOpticalFlowPyramid(a, win, ml_req, withDerivatives=True)
        ml, pyr = oracle.buildOpticalFlowPyramid(a, win, ml_req, True)
        assert ml == ml_cv, (h, w, win, ml_req)
        for l in range(ml + 1):
            assert np.array_equal(pyr[2 * l], pyr_cv[2 * l]), (h, w, l)
            assert np.array_equal(pyr[2 * l + 1], pyr_cv[2 * l + 1]), (h, w, l)


@pytest.mark.parametrize("seed", range(6))
def test_gftt_live(oracle, seed):
    """ordered corner lists identical to cv2's for random sizes, parameters and masks"""
    rng = np.random.default_rng(300 + seed)
    h, w = int(rng.integers(40, 260)), int(rng.integers(40, 330))
    img = _texture(rng, h, w, smooth=int(rng.integers(1, 4)))
    mask = (rng.random((h, w)) > 0.3).astype(np.uint8) * 255
    mask[: h // 5] = 0
    for _ in range(4):
        gp = dict(maxCorners=int(rng.choice([0, 25, 400, 50000000])), qualityLevel=float(rng.choice([0.007, 0.01, 0.05, 0.2])),
                  minDistance=float(rng.choice([0, 1, 3.5, 10, 25])), blockSize=int(rng.choice([3, 5, 10])))
        for m in (None, mask):
            ref = as_corners(cv2.goodFeaturesToTrack(img, mask=m, **gp))
            got = as_corners(oracle.goodFeaturesToTrack(img, mask=m, **gp))
            if got.shape == ref.shape and np.array_equal(got, ref):
                continue
            # exact response ties / last-bit differences of the float min-eigenvalue can swap neighbours in the order:
            # the SET must still agree to BASELINE's 99 %, and any difference must be such a swap
            sa = set(map(tuple, ref.reshape(-1, 2).tolist()))
            sb = set(map(tuple, got.reshape(-1, 2).tolist()))
            assert len(sa & sb) >= 0.99 * max(len(sa), len(sb)), (seed, gp, m is not None, len(sa), len(sb), len(sa & sb))


@pytest.mark.parametrize("seed", range(6))
def test_lk_live(oracle, seed):
    """forward + backward LK on displaced frames: status, positions and err against cv2 for random windows, levels,
    criteria and points (including points next to and outside the border)"""
    rng = np.random.default_rng(400 + seed)
    h, w = int(rng.integers(70, 260)), int(rng.integers(70, 330))
    f0 = _texture(rng, h, w, smooth=2)
    dx, dy = rng.uniform(-2.5, 2.5, 2)
    f1 = _shifted(f0, float(dx), float(dy))
    f1 = np.clip(f1.astype(np.int16) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    n = 160
    pts = np.stack([rng.uniform(-3, w + 3, n), rng.uniform(-3, h + 3, n)], 1).astype(np.float32)
    pts[:20] = np.round(pts[:20])                        # integer positions (what goodFeaturesToTrack returns)
    more = np.stack([rng.uniform(0, w, 640), rng.uniform(0, h, 640)], 1).astype(np.float32)
    pts = np.concatenate([pts, more]).reshape(-1, 1, 2)  # 800 points (the GPU live test draws the same ones)
    for _ in range(3):
        win = (int(rng.choice([5, 9, 15, 21, 31, 35])), int(rng.choice([5, 9, 15, 21, 31, 35])))
        lp = dict(winSize=win, maxLevel=int(rng.integers(0, 5)),
                  criteria=(int(rng.choice([1, 2, 3])), int(rng.integers(1, 31)), float(rng.choice([0.0, 0.01, 0.03, 0.3]))))
        r_p1, r_st, r_err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        p1, st, err = oracle.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
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
        p0r, st0, _ = oracle.calcOpticalFlowPyrLK(f1, f0, r_p1, None, **lp)
        assert_lk_parity(p0r, st0, r_p0r, r_st0, "bwd %r" % (lp,))
        # the FB decision of s1:329-333 on the points both agree are tracked
        both = (st0.ravel() == 1) & (r_st0.ravel() == 1)
        d_ref = np.hypot(*(pts.reshape(-1, 2) - r_p0r.reshape(-1, 2)).T)
        d_got = np.hypot(*(pts.reshape(-1, 2) - p0r.reshape(-1, 2)).T)
        clear = both & (np.abs(d_ref - 1.0) > 0.02)      # decisions within the position tolerance of the threshold may flip
        assert np.array_equal((d_ref < 1.0)[clear], (d_got < 1.0)[clear]), lp


def test_min_eigen_live(oracle):
    rng = np.random.default_rng(500)
    for (h, w), bs in (((41, 67), 3), ((90, 120), 10), ((64, 64), 5)):
        img = _texture(rng, h, w)
        ref = cv2.cornerMinEigenVal(img, bs, ksize=3)
        got = oracle.cornerMinEigenVal(img, bs)
        assert np.abs(got - ref).max() <= 3e-4 * np.abs(ref).max(), (h, w, bs)


@pytest.mark.parametrize("seed", range(4))
def test_harris_live(oracle, seed):
    """the Harris branch (useHarrisDetector=True; never set by the reference, part of cv2's signature): response maps and
    corner lists against cv2 for random sizes, block sizes, k and masks"""
    import harris_cases as HC
    rng = np.random.default_rng(900 + seed)
    h, w = int(rng.integers(40, 260)), int(rng.integers(40, 330))
    img = _texture(rng, h, w, smooth=int(rng.integers(1, 4)))
    mask = (rng.random((h, w)) > 0.3).astype(np.uint8) * 255
    for bs in (2, 3, 5, 10):
        k = float(rng.choice([0.0, 0.04, 0.1]))
        ref = cv2.cornerHarris(img, bs, 3, k)
        assert np.abs(oracle.cornerHarris(img, bs, 3, k) - ref).max() <= HC.MAP_TOL * np.abs(ref).max(), (bs, k)
        for q, md in ((0.01, 0), (0.05, 5), (0.01, 10)):
            for m in (None, mask):
                gp = dict(maxCorners=int(rng.choice([0, 300])), qualityLevel=q, minDistance=md, blockSize=bs,
                          useHarrisDetector=True, k=k)
                HC.check_lists(oracle.goodFeaturesToTrack(img, mask=m, **gp), cv2.goodFeaturesToTrack(img, mask=m, **gp),
                               (seed, gp, m is not None))


@pytest.mark.parametrize("seed", range(3))
def test_lk_multichannel_live(oracle, seed):
    """3-channel frames (cv2's signature takes them; the reference converts to gray first): window sums run over pixels and
    channels -- status, positions and err against cv2 for random windows, levels and criteria"""
    import multichannel_cases as MC
    rng = np.random.default_rng(700 + seed)
    h, w = int(rng.integers(70, 260)), int(rng.integers(70, 330))
    f0 = np.dstack([_texture(rng, h, w) for _ in range(3)])
    dx, dy = rng.uniform(-2.5, 2.5, 2)
    f1 = np.dstack([_shifted(f0[..., c], float(dx), float(dy)) for c in range(3)])
    f1 = np.clip(f1.astype(np.int16) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    pts = np.stack([rng.uniform(-3, w + 3, 500), rng.uniform(-3, h + 3, 500)], 1).astype(np.float32).reshape(-1, 1, 2)
    for _ in range(3):
        win = (int(rng.choice([5, 9, 15, 21, 31, 35])), int(rng.choice([5, 9, 15, 21, 31, 35])))
        lp = dict(winSize=win, maxLevel=int(rng.integers(0, 5)),
                  criteria=(3, int(rng.integers(1, 31)), float(rng.choice([0.0, 0.01, 0.03]))))
        r_p1, r_st, r_err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        p1, st, err = oracle.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        assert_lk_parity(p1, st, r_p1, r_st, "3-channel %r" % (lp,))
        MC.check_err(err, st, p1, r_err, r_st, r_p1, win, lp)
