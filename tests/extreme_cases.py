"""Shared body of the extreme-parameter checks: `impl` is either the CPU oracle or the CUDA path (same cv2 signatures)."""
import numpy as np

from parity import ERR_TOL, assert_lk_parity, as_corners, corner_overlap, CORNER_OVERLAP

LK_EXTREME = [
    dict(winSize=(3, 3), maxLevel=0, criteria=(3, 30, 0.01)),
    dict(winSize=(63, 63), maxLevel=7, criteria=(3, 30, 0.01)),
    dict(winSize=(5, 61), maxLevel=2, criteria=(3, 200, 0.0)),
    dict(winSize=(33, 32), maxLevel=3, criteria=(3, 0, 20.0)),
    dict(winSize=(21, 21), maxLevel=3, criteria=(0, 5, 0.5)),
    dict(winSize=(45, 47), maxLevel=1, criteria=(3, 30, 0.01), minEigThreshold=0.01),
]
GFTT_EXTREME = [
    dict(maxCorners=0, qualityLevel=0.01, minDistance=1, blockSize=4),
    dict(maxCorners=0, qualityLevel=0.01, minDistance=0.5, blockSize=2),
    dict(maxCorners=1, qualityLevel=0.5, minDistance=3, blockSize=3),
    dict(maxCorners=0, qualityLevel=0.001, minDistance=33.3, blockSize=31),
    dict(maxCorners=40, qualityLevel=0.05, minDistance=2.5, blockSize=17),
    dict(maxCorners=0, qualityLevel=0.9, minDistance=4, blockSize=7),
]
TINY = [(24, 24), (9, 130), (130, 9), (40, 37)]


def check_lk_extreme(impl, g):
    for li, lp in enumerate(LK_EXTREME):
        p1, st, err = impl.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["lk_pts"], None, **lp)
        assert_lk_parity(p1, st, g["lk%d_p1" % li], g["lk%d_st" % li], "extreme LK set %d" % li)
        ok = (st == 1) & (g["lk%d_st" % li] == 1)
        assert np.abs(err - g["lk%d_err" % li])[ok].max(initial=0) <= ERR_TOL, li
        d = np.abs(p1 - g["lk%d_p1" % li]).reshape(-1, 2).max(1)          # failed points keep the propagated guess
        assert np.mean(d <= 0.01) >= 0.99, li


def check_gftt_extreme(impl, g):
    for gi, gp in enumerate(GFTT_EXTREME):
        for mi, m in enumerate([None, g["mask"]]):
            ref = g["gftt%d_m%d" % (gi, mi)]
            got = as_corners(impl.goodFeaturesToTrack(g["f0"], mask=m, **gp))
            assert got.shape == ref.shape, (gi, mi, got.shape, ref.shape)
            assert corner_overlap(got, ref) >= CORNER_OVERLAP, (gi, mi)
            if len(ref) and gp["blockSize"] >= 3:
                assert np.mean(np.all(got == ref, axis=(1, 2))) >= 0.99, (gi, mi)      # same order too


def check_tiny(impl, g):
    for (h, w) in TINY:
        k = "tiny_%dx%d_" % (h, w)
        assert np.array_equal(impl.cvtColor(g[k + "rgb"], 6), g[k + "gray"]), k
        ml, pyr = impl.buildOpticalFlowPyramid(g[k + "a"], (5, 5), 3, True)
        assert ml == int(g[k + "ml"]), k
        for l in range(ml + 1):
            assert np.array_equal(pyr[2 * l], g[k + "L%d" % l]) and np.array_equal(pyr[2 * l + 1], g[k + "D%d" % l]), (k, l)
        p1, st, _ = impl.calcOpticalFlowPyrLK(g[k + "a"], g[k + "b"], g[k + "pts"], None, winSize=(21, 21), maxLevel=3,
                                              criteria=(3, 30, 0.01))
        assert_lk_parity(p1, st, g[k + "p1"], g[k + "st"], k)
        ref = g[k + "gftt"]
        got = as_corners(impl.goodFeaturesToTrack(g[k + "a"], 0, 0.05, 3, blockSize=3))
        assert got.shape == ref.shape and corner_overlap(got, ref) >= CORNER_OVERLAP, k
