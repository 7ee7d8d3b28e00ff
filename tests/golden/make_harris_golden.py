"""Generates tests/golden/kat_harris.npz: what the cv2 wheel (opencv-python-headless 4.13.0.92) returns for the Harris branch
of the corner detector -- cv2.cornerHarris(img, blockSize, 3, k) and cv2.goodFeaturesToTrack(..., useHarrisDetector=True, k=k)
-- on the two golden scenes already stored in kat_texture.npz / kat_iceberg.npz (their f0 frame and mask).  The reference
leaves useHarrisDetector at its default (s1_lucaskanade_tracking.py:240-243); cv2's signature, which the drop-in functions
mirror, carries it.

Run:  python tests/golden/make_harris_golden.py      (needs cv2; not needed to RUN the tests)
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

HARRIS_SETS = [
    dict(maxCorners=0, qualityLevel=0.01, minDistance=10, blockSize=10, k=0.04),
    dict(maxCorners=500, qualityLevel=0.02, minDistance=5, blockSize=3, k=0.04),
    dict(maxCorners=0, qualityLevel=0.05, minDistance=0, blockSize=5, k=0.1),
    dict(maxCorners=200, qualityLevel=0.01, minDistance=7.5, blockSize=2, k=0.0),
]


def main():
    out = {}
    for scene in ("texture", "iceberg"):
        g = np.load(os.path.join(HERE, "kat_%s.npz" % scene))
        f0, mask = g["f0"], g["mask"]
        for bs, k in ((3, 0.04), (10, 0.04), (5, 0.1)):
            out["%s_harris_bs%d_k%g" % (scene, bs, k)] = cv2.cornerHarris(f0, bs, 3, k)
        for si, gp in enumerate(HARRIS_SETS):
            for mi, m in enumerate((None, mask)):
                p = cv2.goodFeaturesToTrack(f0, mask=m, useHarrisDetector=True, **gp)
                out["%s_gftt%d_m%d" % (scene, si, mi)] = np.zeros((0, 1, 2), np.float32) if p is None else p
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "kat_harris.npz"), **out)
    print("kat_harris.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
