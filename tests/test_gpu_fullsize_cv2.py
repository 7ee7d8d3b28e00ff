"""Full-size parity against cv2 ITSELF (the dependency that holds the reference's arithmetic, s1:323,326,437) at BASELINE's
headline sizes: every point, not a sample.  cv2 is importable on the GPU box (bench.py's cpu_baseline uses it); the tests
skip -- they do not pass -- when it is absent.  /root/reference is not touched."""
import numpy as np
import pytest
import torch

from parity import CORNER_OVERLAP, POS_FRAC, POS_TOL_PX, STATUS_FRAC, lk_agreement

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")

H, W = 4000, 6000


@pytest.fixture(scope="module")
def frames24():
    from iceberg_tracking_code_b200 import synthetic as syn
    base = syn.base_texture(H, W, 7, device="cuda")
    f0, f1 = syn.frame_gray(base, 0), syn.frame_gray(base, 1)
    del base
    torch.cuda.empty_cache()
    return f0, f1, f0.cpu().numpy(), f1.cpu().numpy()


def test_config2_all_corners_and_all_points_vs_cv2(ibt, frames24):
    """BASELINE config 2 (6000x4000, top-20 000 Shi-Tomasi corners, win 31, L4, (3,30,0.01), FB < 1 px):
    the whole ordered corner list and all 20 000 forward / backward points, statuses, errs and FB decisions vs cv2."""
    f0, f1, g0, g1 = frames24
    gp = dict(maxCorners=20000, qualityLevel=0.007, minDistance=10, blockSize=10)
    lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
    ours = ibt.goodFeaturesToTrack(f0, **gp).cpu().numpy().reshape(-1, 2)
    ref = cv2.goodFeaturesToTrack(g0, **gp).reshape(-1, 2)
    assert ours.shape == ref.shape == (20000, 2)
    sa, sb = set(map(tuple, ours.astype(np.int64).tolist())), set(map(tuple, ref.astype(np.int64).tolist()))
    assert len(sa & sb) / 20000.0 >= CORNER_OVERLAP
    # same rank for (nearly) every corner: the min-eigenvalue map is reproducible to ~1e-6 relative only, so
    # neighbours in the response order may swap
    assert np.mean(np.all(ours == ref, axis=1)) >= 0.99
    # the complete list (all ~117 k corners, no top-k cut)
    gp0 = dict(gp, maxCorners=0)
    ours0 = ibt.goodFeaturesToTrack(f0, **gp0).cpu().numpy().reshape(-1, 2)
    ref0 = cv2.goodFeaturesToTrack(g0, **gp0).reshape(-1, 2)
    s0, s1 = set(map(tuple, ours0.astype(np.int64).tolist())), set(map(tuple, ref0.astype(np.int64).tolist()))
    assert len(s0 & s1) / max(len(s0), len(s1)) >= CORNER_OVERLAP

    p0 = ref.reshape(-1, 1, 2).astype(np.float32)
    r = ibt.calcOpticalFlowPyrLK_FB(f0, f1, p0, **lp)
    p1c, st1c, err1c = cv2.calcOpticalFlowPyrLK(g0, g1, p0, None, **lp)
    p0rc, st0c, err0c = cv2.calcOpticalFlowPyrLK(g1, g0, p1c, None, **lp)
    for what, pa, sa_, pb, sb_ in (("forward", r["p1"], r["st1"], p1c, st1c), ("backward", r["p0r"], r["st0"], p0rc, st0c)):
        agree, frac, mx = lk_agreement(pa, sa_, pb, sb_)
        assert agree >= STATUS_FRAC and frac >= POS_FRAC, (what, agree, frac, mx)
        assert mx <= POS_TOL_PX, (what, mx)                      # on this scene: every point, not 99 %
    ok = (r["st1"].reshape(-1) == 1) & (st1c.reshape(-1) == 1)
    assert np.abs(r["err1"].reshape(-1) - err1c.reshape(-1))[ok].max() <= 2e-3
    distc = np.hypot(*np.abs(p0.reshape(-1, 2) - p0rc.reshape(-1, 2)).T)
    assert np.mean((r["dist"] < 1) == (distc < 1)) >= STATUS_FRAC
    assert np.abs(r["dist"] - distc).max() <= 2 * POS_TOL_PX
    # bit-exact pyramid + Scharr planes at full size against cv2's own builder
    ml, pyr = cv2.buildOpticalFlowPyramid(g1, (31, 31), 4, withDerivatives=True)
    mlo, po = ibt.buildOpticalFlowPyramid(f1, (31, 31), 4, True)
    assert ml == mlo
    for l in range(ml + 1):
        assert np.array_equal(po[2 * l].cpu().numpy(), pyr[2 * l]), l
        assert np.array_equal(po[2 * l + 1].cpu().numpy(), pyr[2 * l + 1]), l


def test_config4_all_grid_points_vs_cv2(ibt, frames24):
    """BASELINE config 4 (197 835 grid points, maxLevel 5, 30 iterations): every point, forward and backward, vs cv2."""
    from iceberg_tracking_code_b200 import synthetic as syn
    f0, f1, g0, g1 = frames24
    lp = dict(winSize=(31, 31), maxLevel=5, criteria=(3, 30, 0.01))
    p0 = syn.grid_points(H, W, step=11, start=10).numpy()
    assert p0.shape[0] == 197835
    r = ibt.calcOpticalFlowPyrLK_FB(f0, f1, p0, **lp)
    p1c, st1c, _ = cv2.calcOpticalFlowPyrLK(g0, g1, p0, None, **lp)
    p0rc, st0c, _ = cv2.calcOpticalFlowPyrLK(g1, g0, p1c, None, **lp)
    agree, frac, mx = lk_agreement(r["p1"], r["st1"], p1c, st1c)
    assert agree >= STATUS_FRAC and frac >= POS_FRAC, (agree, frac, mx)
    agree, frac, mx = lk_agreement(r["p0r"], r["st0"], p0rc, st0c)
    assert agree >= STATUS_FRAC and frac >= POS_FRAC, (agree, frac, mx)
    # failed / textureless points keep the propagated guess (the reference consumes them: status is never read, s1:326-333)
    d = np.abs(r["p1"].reshape(-1, 2) - p1c.reshape(-1, 2)).max(1)
    assert np.mean(d <= POS_TOL_PX) >= POS_FRAC
    distc = np.hypot(*np.abs(p0.reshape(-1, 2) - p0rc.reshape(-1, 2)).T)
    assert np.mean((r["dist"] < 1) == (distc < 1)) >= STATUS_FRAC
