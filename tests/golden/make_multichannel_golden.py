"""Generates tests/golden/kat_multichannel.npz: cv2.calcOpticalFlowPyrLK (opencv-python-headless 4.13.0.92) on 3-CHANNEL frames.
cv2's signature takes them; the reference converts to gray first (s1_lucaskanade_tracking.py:311), so this pins a
part of the boundary the reference itself never exercises.  Inputs are stored next to the outputs.

Run:  python tests/golden/make_multichannel_golden.py      (needs cv2; not needed to RUN the tests)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_oracle_live_cv2 as L  # noqa: E402  (the seeded texture generator)

MC_SETS = [
    dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01)),
    dict(winSize=(35, 35), maxLevel=4, criteria=(3, 25, 0.03)),       # the reference's window (s1:246-248)
    dict(winSize=(9, 15), maxLevel=1, criteria=(1, 6, 0.0)),
]


def main():
    rng = np.random.default_rng(4242)
    h, w = 150, 212
    f0 = np.dstack([L._texture(rng, h, w) for _ in range(3)])
    f1 = np.dstack([L._shifted(f0[..., c], 1.3125, -0.84375) for c in range(3)])
    f1 = np.clip(f1.astype(np.int16) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    pts = np.stack([rng.uniform(-3, w + 3, 400), rng.uniform(-3, h + 3, 400)], 1).astype(np.float32).reshape(-1, 1, 2)
    out = dict(f0=f0, f1=f1, pts=pts, cv2_version=np.array(cv2.__version__))
    for i, lp in enumerate(MC_SETS):
        p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        out["lk%d_p1" % i], out["lk%d_st" % i], out["lk%d_err" % i] = p1, st, err
    guess = (pts + np.float32([1.0, -1.0])).astype(np.float32)
    p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, guess.copy(), flags=cv2.OPTFLOW_USE_INITIAL_FLOW, **MC_SETS[0])
    out["init_guess"], out["init_p1"], out["init_st"] = guess, p1, st
    # a 4-channel pair (cv2 takes any channel count) and the (H,W,1) form of a gray pair
    f0c4, f1c4 = np.dstack([f0, f0[..., 0]]), np.dstack([f1, f1[..., 1]])
    p1, st, err = cv2.calcOpticalFlowPyrLK(f0c4, f1c4, pts, None, **MC_SETS[0])
    out["c4_p1"], out["c4_st"], out["c4_err"] = p1, st, err
    p1, st, err = cv2.calcOpticalFlowPyrLK(f0[..., :1].copy(), f1[..., :1].copy(), pts, None, **MC_SETS[0])
    out["c1_p1"], out["c1_st"], out["c1_err"] = p1, st, err
    np.savez_compressed(os.path.join(HERE, "kat_multichannel.npz"), **out)
    print("kat_multichannel.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
