"""Shared body of the multi-channel calcOpticalFlowPyrLK checks: `impl` is either the CPU oracle or the CUDA path (same cv2
signature); `g` = tests/golden/kat_multichannel.npz (cv2's answers on 3- and 4-channel frames)."""
import numpy as np

from parity import ERR_TOL, assert_lk_parity

MC_SETS = [
    dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01)),
    dict(winSize=(35, 35), maxLevel=4, criteria=(3, 25, 0.03)),
    dict(winSize=(9, 15), maxLevel=1, criteria=(1, 6, 0.0)),
]
OPTFLOW_USE_INITIAL_FLOW = 4


def check_err(err, st, p1, r_err, r_st, r_p1, win, what):
    same = np.abs(np.asarray(p1) - r_p1).reshape(-1, 2).max(1) <= 1e-4        # err is the residual AT the returned position
    ok = (np.asarray(st).ravel() == 1) & (r_st.ravel() == 1) & same
    assert np.abs(np.asarray(err).ravel()[ok] - r_err.ravel()[ok]).max(initial=0) <= ERR_TOL + 16.0 / (32 * win[0] * win[1]), what


def check_multichannel_golden(impl, g):
    f0, f1, pts = g["f0"], g["f1"], g["pts"]
    for i, lp in enumerate(MC_SETS):
        p1, st, err = impl.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        assert p1.shape == pts.shape and st.shape == (len(pts), 1) and err.shape == (len(pts), 1)
        assert_lk_parity(p1, st, g["lk%d_p1" % i], g["lk%d_st" % i], "3-channel set %d" % i)
        check_err(err, st, p1, g["lk%d_err" % i], g["lk%d_st" % i], g["lk%d_p1" % i], lp["winSize"], i)
        d = np.abs(p1 - g["lk%d_p1" % i]).reshape(-1, 2).max(1)               # failed points keep the propagated guess
        assert np.mean(d <= 0.01) >= 0.99, i
    p1, st, _ = impl.calcOpticalFlowPyrLK(f0, f1, pts, g["init_guess"].copy(), flags=OPTFLOW_USE_INITIAL_FLOW, **MC_SETS[0])
    assert_lk_parity(p1, st, g["init_p1"], g["init_st"], "3-channel, initial flow")
    f0c4, f1c4 = np.dstack([f0, f0[..., 0]]), np.dstack([f1, f1[..., 1]])
    p1, st, err = impl.calcOpticalFlowPyrLK(f0c4, f1c4, pts, None, **MC_SETS[0])
    assert_lk_parity(p1, st, g["c4_p1"], g["c4_st"], "4-channel")
    check_err(err, st, p1, g["c4_err"], g["c4_st"], g["c4_p1"], MC_SETS[0]["winSize"], "c4")
    p1, st, err = impl.calcOpticalFlowPyrLK(f0[..., :1].copy(), f1[..., :1].copy(), pts, None, **MC_SETS[0])
    assert_lk_parity(p1, st, g["c1_p1"], g["c1_st"], "(H,W,1)")
    # three identical channels: every window sum is three times the gray one, so the track is the single-channel track
    gray3 = np.dstack([f0[..., 0]] * 3), np.dstack([f1[..., 0]] * 3)
    a1, s1, _ = impl.calcOpticalFlowPyrLK(gray3[0], gray3[1], pts, None, **MC_SETS[0])
    b1, s2, _ = impl.calcOpticalFlowPyrLK(f0[..., 0].copy(), f1[..., 0].copy(), pts, None, **MC_SETS[0])
    assert_lk_parity(a1, s1, b1, s2, "3 x gray vs gray")
