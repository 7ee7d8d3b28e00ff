"""Pins the CPU oracle (oracle/ibt_oracle.c) against the cv2-generated known answers and the outputs of the
reference's own code recorded in tests/golden/ (make_golden.py).  CPU only."""
import numpy as np
import pytest

from parity import GFTT_SETS, LK_SETS, ERR_TOL, assert_lk_parity, as_corners

SCENES = ["kat_texture.npz", "kat_iceberg.npz"]


@pytest.mark.parametrize("scene", SCENES)
def test_gray_bit_exact(oracle, golden, scene):
    g = golden(scene)
    assert np.array_equal(oracle.cvtColor(g["rgb"]), g["gray"])
    rgba = np.dstack([g["rgb"], g["rgb"][..., 0]])
    assert np.array_equal(oracle.cvtColor(rgba), g["gray4"])


@pytest.mark.parametrize("scene", SCENES)
def test_pyramid_bit_exact(oracle, golden, scene):
    g = golden(scene)
    assert np.array_equal(oracle.pyrDown(g["f0"]), g["pyrdown"])
    for si, win in enumerate([(21, 21), (35, 35)]):
        ml, pyr = oracle.buildOpticalFlowPyramid(g["f0"], win, 4, True)
        assert ml == int(g["pyr%d_maxlevel" % si])
        for l in range(ml + 1):
            assert np.array_equal(pyr[2 * l], g["pyr%d_L%d" % (si, l)]), (si, l)
            assert np.array_equal(pyr[2 * l + 1], g["pyr%d_D%d" % (si, l)]), (si, l)


def test_edge_cases(oracle, golden):
    g = golden("kat_edge.npz")
    for k in g:
        if k.startswith("rnd_"):
            shp = k[4:]
            a = g[k]
            assert np.array_equal(oracle.pyrDown(a), g["pyrdown_" + shp]), shp
            if "scharr_" + shp in g:
                assert np.array_equal(oracle.scharr_deriv(a), g["scharr_" + shp]), shp
        if k.startswith("levels_"):
            h, w = [int(v) for v in k.split("_")[1].split("x")]
            win = int(k.split("_w")[1])
            ml, sizes = oracle.pyramid_sizes(h, w, (win, win), 4)
            assert [ml] + [x for s in sizes for x in s] == g[k].tolist(), k
    flat = np.full((40, 50), 77, np.uint8)
    assert int(g["flat_gftt_none"]) == 1 and oracle.goodFeaturesToTrack(flat, 100, 0.01, 5) is None
    a = g["emptymask_img"]
    assert int(g["emptymask_none"]) == 1
    assert oracle.goodFeaturesToTrack(a, 100, 0.01, 5, mask=np.zeros_like(a)) is None


@pytest.mark.parametrize("scene", SCENES)
def test_mineig_close(oracle, golden, scene):
    g = golden(scene)
    for bs, tol in ((3, 3e-4), (10, 4e-6)):
        ref = g["mineig_bs%d" % bs]
        got = oracle.cornerMinEigenVal(g["f0"], bs)
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= tol * scale, bs


@pytest.mark.parametrize("scene", SCENES)
def test_gftt_identical_lists(oracle, golden, scene):
    g = golden(scene)
    for gi, gp in enumerate(GFTT_SETS):
        for mi, m in enumerate([None, g["mask"]]):
            ref = g["gftt%d_m%d" % (gi, mi)]
            got = as_corners(oracle.goodFeaturesToTrack(g["f0"], mask=m, **gp))
            assert got.shape == ref.shape and np.array_equal(got, ref), (gi, mi, got.shape, ref.shape)


@pytest.mark.parametrize("scene", SCENES)
def test_lk_parity(oracle, golden, scene):
    g = golden(scene)
    pts = g["lk_pts"]
    for li, lp in enumerate(LK_SETS):
        p1, st, err = oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], pts, None, **lp)
        assert_lk_parity(p1, st, g["lk%d_p1" % li], g["lk%d_st" % li], "fwd set %d" % li)
        ok = (st == 1) & (g["lk%d_st" % li] == 1)
        assert np.abs(err - g["lk%d_err" % li])[ok].max(initial=0) <= ERR_TOL
        # backward pass from the golden p1 (the input the reference's second call sees, s1:326)
        p0r, st0, err0 = oracle.calcOpticalFlowPyrLK(g["f1"], g["f0"], g["lk%d_p1" % li], None, **lp)
        assert_lk_parity(p0r, st0, g["lk%d_p0r" % li], g["lk%d_st0" % li], "bwd set %d" % li)
        # failed points keep the propagated guess: compare positions on ALL points too (A.7 relies on it)
        d = np.abs(p1 - g["lk%d_p1" % li]).reshape(-1, 2).max(1)
        assert np.mean(d <= 0.01) >= 0.99, li


def test_lk_flags(oracle, golden):
    g = golden("kat_texture.npz")
    p1, st, _ = oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["lk_pts"], g["lkinit_guess"],
                                            flags=oracle.OPTFLOW_USE_INITIAL_FLOW, **LK_SETS[0])
    assert_lk_parity(p1, st, g["lkinit_p1"], g["lkinit_st"], "initial flow")
    p1, st, err = oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["lk_pts"], None,
                                              flags=oracle.OPTFLOW_LK_GET_MIN_EIGENVALS, **LK_SETS[0])
    assert_lk_parity(p1, st, g["lkeig_p1"], g["lkeig_st"], "min eigenvals")
    ok = (st == 1) & (g["lkeig_st"] == 1)
    ref = g["lkeig_err"]
    assert np.abs(err - ref)[ok].max() <= 1e-5 * max(1.0, np.abs(ref[ok]).max())


def test_lk_empty_and_dtype(oracle, golden):
    g = golden("kat_texture.npz")
    assert oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], np.zeros((0, 1, 2), np.float32)) == (None, None, None)
    with pytest.raises(AssertionError):
        oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], np.zeros((3, 1, 2), np.float64))


def test_photo_to_utm_vs_reference(oracle, golden):
    g = golden("utm_expected.npz")
    got = oracle.photo_to_utm(g["xy"].astype(np.float64), g["cam"])
    assert np.abs(got - g["EN"]).max() <= 1e-6          # metres


def test_harris_branch_golden(oracle, golden):
    """goodFeaturesToTrack(useHarrisDetector=True) / cornerHarris vs cv2's recorded answers (make_harris_golden.py)"""
    import harris_cases as HC
    HC.check_harris_golden(oracle, golden("kat_harris.npz"),
                           {"texture": golden("kat_texture.npz"), "iceberg": golden("kat_iceberg.npz")})


def test_multichannel_lk_golden(oracle, golden):
    """calcOpticalFlowPyrLK on 3- / 4-channel frames vs cv2's recorded answers (make_multichannel_golden.py)"""
    import multichannel_cases as MC
    g = golden("kat_multichannel.npz")
    MC.check_multichannel_golden(oracle, g)
    cv2 = pytest.importorskip("cv2")                     # OPTFLOW_LK_GET_MIN_EIGENVALS on 3 channels: still divided by 2*w*h, no cn
    lp = MC.MC_SETS[0]
    r = cv2.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["pts"], None, flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS, **lp)
    o = oracle.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["pts"], None, flags=oracle.OPTFLOW_LK_GET_MIN_EIGENVALS, **lp)
    assert_lk_parity(o[0], o[1], r[0], r[1], "3-channel, min eigenvalues")
    ok = (r[1].ravel() == 1) & (o[1].ravel() == 1)
    assert np.abs(r[2].ravel()[ok] - o[2].ravel()[ok]).max() <= 1e-5
