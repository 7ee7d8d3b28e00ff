"""Shared body of the Harris-branch checks (goodFeaturesToTrack(useHarrisDetector=True), cornerHarris): `impl` is either the
CPU oracle or the CUDA path (same cv2 signatures); `g` = tests/golden/kat_harris.npz (cv2's answers), `scenes` = the golden
scenes that hold the input frame and mask."""
import numpy as np

from parity import as_corners, corner_overlap, CORNER_OVERLAP

HARRIS_SETS = [
    dict(maxCorners=0, qualityLevel=0.01, minDistance=10, blockSize=10, k=0.04),
    dict(maxCorners=500, qualityLevel=0.02, minDistance=5, blockSize=3, k=0.04),
    dict(maxCorners=0, qualityLevel=0.05, minDistance=0, blockSize=5, k=0.1),
    dict(maxCorners=200, qualityLevel=0.01, minDistance=7.5, blockSize=2, k=0.0),
]
# the response subtracts two nearly equal products: last-bit differences of the window sums (OpenCV adds float products, the
# CUDA path exact integers) show as ~1e-6 of the map's range, and may swap neighbours of almost equal response in the list
MAP_TOL = 2e-6
SAME_RANK = 0.98


def check_lists(got, ref, what):
    got, ref = as_corners(got), as_corners(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert corner_overlap(got, ref) >= CORNER_OVERLAP, what
    if len(ref):
        assert np.mean(np.all(got == ref, axis=(1, 2))) >= SAME_RANK, what


def check_harris_golden(impl, g, scenes):
    for name, sc in scenes.items():
        f0, mask = sc["f0"], sc["mask"]
        for bs, k in ((3, 0.04), (10, 0.04), (5, 0.1)):
            ref = g["%s_harris_bs%d_k%g" % (name, bs, k)]
            got = impl.cornerHarris(f0, bs, 3, k)
            assert got.shape == ref.shape and got.dtype == np.float32
            assert np.abs(got - ref).max() <= MAP_TOL * np.abs(ref).max(), (name, bs, k)
        for si, gp in enumerate(HARRIS_SETS):
            for mi, m in enumerate((None, mask)):
                got = impl.goodFeaturesToTrack(f0, mask=m, useHarrisDetector=True, **gp)
                check_lists(got, g["%s_gftt%d_m%d" % (name, si, mi)], (name, si, mi))
    flat = np.full((40, 50), 77, np.uint8)                   # response 0 everywhere -> None, like cv2
    assert impl.goodFeaturesToTrack(flat, 100, 0.01, 5, useHarrisDetector=True) is None
    # an image of straight edges only: every Harris response is <= 0 -> no corner, whatever the quality level
    edges = np.zeros((60, 80), np.uint8)
    edges[:, 40:] = 200
    assert impl.goodFeaturesToTrack(edges, 100, 0.01, 5, blockSize=3, useHarrisDetector=True, k=0.04) is None
