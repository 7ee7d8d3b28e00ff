"""Generates the golden fixtures under tests/golden/ from the real thing, in the BUILD container:

  * kat_*.npz      -- outputs of the cv2 wheel (opencv-python-headless 4.13.0.92), the third-party dependency that holds
                      all arithmetic of the reference's tracking hot path (s1_lucaskanade_tracking.py:311,323,326,437),
                      on small seeded synthetic frames.  Inputs are stored next to the outputs.
  * seq/*.jpg, seq_expected.npz
                   -- five synthetic JPEG frames and what the UNMODIFIED reference class
                      /root/reference/s0_1_test_lucaskanade_tracking.py:LucasKanade.run() produced on them (matplotlib
                      replaced by MagicMock; cv2 calls recorded), i.e. outputs of the reference itself run here.
  * kat_extreme.npz  -- cv2 outputs at parameter extremes (window 3..63, maxLevel 0..7, criteria clamps, blockSize 1..31,
                      fractional / huge minDistance, single-pixel mask) and awkward shapes (odd widths, 9-pixel-thin images).
  * s2_expected.npz  -- the UNMODIFIED reference worker s2_cam_to_utm.cam_to_utm run on synthetic track files crossing an
                      hour boundary (tide pickle as the reference reads it; pd.read_excel patched to return the parameter
                      table because openpyxl is absent): inputs and the hourly x, y, u, v, speed, time arrays it wrote.
  * utm_expected.npz -- Camera.photo_to_utm / photocords_cropped_to_uncropped of /root/reference/imports/camtools.py
                      evaluated on a Camera object whose Excel parsing is bypassed (object.__new__ + the dict fields
                      __init__ would fill, camtools.py:126-147).

Run:  python tests/golden/make_golden.py      (needs cv2, PIL and /root/reference; not needed to RUN the tests)
"""
import os
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from iceberg_tracking_code_b200 import synthetic as syn  # noqa: E402

LK_SETS = [
    dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01)),       # BASELINE config 1
    dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01)),       # BASELINE config 2
    dict(winSize=(35, 35), maxLevel=4, criteria=(3, 25, 0.03)),       # reference s1:246-248
    dict(winSize=(15, 9), maxLevel=2, criteria=(1, 7, 0.0)),          # COUNT only, rectangular window
    dict(winSize=(41, 41), maxLevel=5, criteria=(2, 0, 0.05)),        # EPS only, two strips
]
GFTT_SETS = [
    dict(maxCorners=50000000, qualityLevel=0.007, minDistance=10, blockSize=10),   # reference s1:240-243
    dict(maxCorners=2000, qualityLevel=0.01, minDistance=7, blockSize=3),
    dict(maxCorners=0, qualityLevel=0.007, minDistance=10, blockSize=10),
    dict(maxCorners=500, qualityLevel=0.05, minDistance=0, blockSize=3),
    dict(maxCorners=300, qualityLevel=0.007, minDistance=10.5, blockSize=10),
    dict(maxCorners=120, qualityLevel=0.02, minDistance=25, blockSize=5),
]


def scene(h, w, seed, kind):
    base = syn.base_texture(h, w, seed, scene=kind)
    f0 = syn.frame_gray(base, 0, noise_sigma=1.0, seed=seed).numpy()
    f1 = syn.frame_gray(base, 1, noise_sigma=1.0, seed=seed).numpy()
    rgb = syn.frame_rgb(base, 0, seed=seed).numpy()
    return f0, f1, rgb


def lk_points(h, w, f0, rng):
    p = cv2.goodFeaturesToTrack(f0, 150, 0.01, 7, blockSize=5)
    p = np.zeros((0, 1, 2), np.float32) if p is None else p
    grid = np.stack(np.meshgrid(np.arange(5, w, 37), np.arange(5, h, 29)), -1).reshape(-1, 1, 2).astype(np.float32)
    border = np.float32([[0, 0], [w - 1, h - 1], [0.5, h - 1.25], [w - 0.75, 0.25], [w / 2, 0], [3.3, h / 2],
                         [-4.5, 10], [w + 6.0, h - 3.0], [w - 1, 0], [0, h - 1]]).reshape(-1, 1, 2)
    sub = (rng.random((60, 1, 2)) * [w - 1, h - 1]).astype(np.float32)
    return np.concatenate([p.astype(np.float32), grid, border, sub], 0)


def make_kat(name, h, w, seed, kind):
    rng = np.random.default_rng(seed)
    f0, f1, rgb = scene(h, w, seed, kind)
    out = dict(f0=f0, f1=f1, rgb=rgb)
    out["gray"] = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)
    out["gray4"] = cv2.cvtColor(np.dstack([rgb, rgb[..., 0]]), cv2.COLOR_BGRA2GRAY)
    out["pyrdown"] = cv2.pyrDown(f0)
    for si, win in enumerate([(21, 21), (35, 35)]):
        ml, pyr = cv2.buildOpticalFlowPyramid(f0, win, 4, withDerivatives=True)
        out["pyr%d_maxlevel" % si] = np.int32(ml)
        for l in range(ml + 1):
            out["pyr%d_L%d" % (si, l)] = np.ascontiguousarray(pyr[2 * l])
            out["pyr%d_D%d" % (si, l)] = np.ascontiguousarray(pyr[2 * l + 1])
    for bs in (3, 10):
        out["mineig_bs%d" % bs] = cv2.cornerMinEigenVal(f0, bs, ksize=3)
    mask = np.zeros((h, w), np.uint8)
    mask[h // 5: h - h // 6, w // 7: w - w // 4] = 255
    mask[h // 2: h // 2 + 9, :] = 0
    out["mask"] = mask
    for gi, gp in enumerate(GFTT_SETS):
        for mi, m in enumerate([None, mask]):
            p = cv2.goodFeaturesToTrack(f0, mask=m, **gp)
            out["gftt%d_m%d" % (gi, mi)] = np.zeros((0, 1, 2), np.float32) if p is None else p
    pts = lk_points(h, w, f0, rng)
    out["lk_pts"] = pts
    for li, lp in enumerate(LK_SETS):
        p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        p0r, st0, err0 = cv2.calcOpticalFlowPyrLK(f1, f0, p1, None, **lp)
        out["lk%d_p1" % li], out["lk%d_st" % li] = p1, st
        out["lk%d_err" % li] = np.where(st == 1, err, 0).astype(np.float32)        # err at st==0 is uninitialised in cv2
        out["lk%d_p0r" % li], out["lk%d_st0" % li] = p0r, st0
        out["lk%d_err0" % li] = np.where(st0 == 1, err0, 0).astype(np.float32)
    # USE_INITIAL_FLOW and GET_MIN_EIGENVALS variants on set 0
    guess = (pts + np.float32([2.0, -1.0])).astype(np.float32)
    p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, guess.copy(), flags=cv2.OPTFLOW_USE_INITIAL_FLOW, **LK_SETS[0])
    out["lkinit_guess"], out["lkinit_p1"], out["lkinit_st"] = guess, p1, st
    p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS, **LK_SETS[0])
    out["lkeig_p1"], out["lkeig_st"], out["lkeig_err"] = p1, st, err
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith(("gftt0", "lk0"))})


def make_edge():
    """Edge cases: flat image, empty mask, tiny images, odd sizes, pyramid level rule."""
    out = {}
    flat = np.full((40, 50), 77, np.uint8)
    out["flat_gftt_none"] = np.int32(cv2.goodFeaturesToTrack(flat, 100, 0.01, 5) is None)
    rng = np.random.default_rng(5)
    for (h, w) in [(33, 35), (5, 7), (3, 3), (64, 1), (1, 64), (2, 9), (101, 203)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out["rnd_%dx%d" % (h, w)] = a
        out["pyrdown_%dx%d" % (h, w)] = cv2.pyrDown(a)
        if h >= 3 and w >= 3:
            dx = cv2.Scharr(a, cv2.CV_16S, 1, 0)
            dy = cv2.Scharr(a, cv2.CV_16S, 0, 1)
            out["scharr_%dx%d" % (h, w)] = np.stack([dx, dy], -1)
    for (h, w, win) in [(100, 120, 35), (140, 140, 35), (142, 142, 35), (64, 48, 21), (45, 45, 21)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ml, pyr = cv2.buildOpticalFlowPyramid(a, (win, win), 4, withDerivatives=False)
        out["levels_%dx%d_w%d" % (h, w, win)] = np.int32([ml] + [x for p in pyr for x in p.shape[:2]])
    a = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    out["emptymask_img"] = a
    out["emptymask_none"] = np.int32(cv2.goodFeaturesToTrack(a, 100, 0.01, 5, mask=np.zeros_like(a)) is None)
    np.savez_compressed(os.path.join(HERE, "kat_edge.npz"), **out)
    print("kat_edge.npz", len(out))


def make_sequence():
    """Run the unmodified reference s0_1 LucasKanade.run() on five synthetic JPEGs."""
    seqdir = os.path.join(HERE, "seq")
    os.makedirs(seqdir, exist_ok=True)
    base = syn.base_texture(270, 480, 11, scene="texture")
    ice = syn.base_texture(270, 480, 12, scene="iceberg")
    names = []
    for t in range(5):
        g = syn.frame_gray(base, t, noise_sigma=1.0, seed=11).numpy().astype(np.float32)
        i = syn.frame_gray(ice, t, vx=1.25, vy=0.5, noise_sigma=1.0, seed=12).numpy().astype(np.float32)
        blend = g.copy()
        blend[:, 300:] = i[:, 300:]                       # right part: icebergs on flat water
        rgb = np.stack([blend, np.clip(blend * 0.9 + 10, 0, 255), np.clip(blend * 1.05, 0, 255)], -1).astype(np.uint8)
        name = "20190724-13%02d00.jpg" % t
        Image.fromarray(rgb).save(os.path.join(seqdir, name), quality=95)
        names.append(name)
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.path"):
        sys.modules[m] = MagicMock()
    sys.path.insert(0, "/root/reference")
    import s0_1_test_lucaskanade_tracking as s01
    rec = {"gftt": [], "lk": []}
    real_gftt, real_lk = cv2.goodFeaturesToTrack, cv2.calcOpticalFlowPyrLK

    class Spy:
        def __getattr__(self, k):
            return getattr(cv2, k)

        def goodFeaturesToTrack(self, *a, **k):
            r = real_gftt(*a, **k)
            rec["gftt"].append(r)
            return r

        def calcOpticalFlowPyrLK(self, *a, **k):
            r = real_lk(*a, **k)
            rec["lk"].append((a[2].copy(), r[0].copy(), r[1].copy()))
            return r

    s01.cv2 = Spy()
    out = {}
    for di in (2, 3):
        rec["gftt"].clear(); rec["lk"].clear()
        lk = s01.LucasKanade(Path(seqdir), di, 60)
        lk.run()
        try:
            os.rmdir(os.path.join(seqdir, "plots_%d" % (60 * di)))
        except OSError:
            pass
        out["d%d_final_tracks" % di] = np.float32(lk.tracks)                     # (K,1,2): seeds of the last detection
        out["d%d_n_gftt" % di] = np.int32(len(rec["gftt"]))
        for i, p in enumerate(rec["gftt"]):
            out["d%d_gftt%d" % (di, i)] = p
        for i, (p0, p1, st) in enumerate(rec["lk"]):
            out["d%d_lk%d_p0" % (di, i)], out["d%d_lk%d_p1" % (di, i)], out["d%d_lk%d_st" % (di, i)] = p0, p1, st
        out["d%d_n_lk" % di] = np.int32(len(rec["lk"]))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "seq_expected.npz"), **out)
    print("seq_expected.npz", {k: v.shape for k, v in out.items() if "final" in k or "n_" in k})


def make_utm():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.path", "shapefile"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, "/root/reference")
    from imports import camtools as ct
    cam = object.__new__(ct.Camera)
    # the values create_calibration_file.py:8-30 writes; fields as Camera.__init__ fills them (camtools.py:126-147)
    width, height, chip = 6000, 4000, 22.3
    cam.pic = {"width": width, "height": height, "cropleft": np.int64(250), "croptop": np.int64(400),
               "cropright": np.int64(0), "cropbottom": np.int64(0)}
    cam.cam = {"chipsize": chip, "E": 377280.39, "N": 6525846.97, "H": 261.3 - 1.6 - 0.37,
               "theta": np.radians(300.0), "phi": np.radians(5.0), "psi": np.radians(-1.0),
               "sigma": (width / chip) * 18.0}
    rng = np.random.default_rng(9)
    xy = (rng.random((4000, 2)) * [5500, 1900] + [0, 1500]).astype(np.float32)   # below the horizon only
    EN = np.empty((len(xy), 2), np.float64)
    for i, (x, y) in enumerate(xy.tolist()):                # python floats, like np.load(...)['tracks'].tolist() at s2:233
        xu, yu = cam.photocords_cropped_to_uncropped(x, y)
        EN[i] = cam.photo_to_utm(xu, yu)
    camvec = np.float64([250, 400, width, height, cam.cam["sigma"], cam.cam["H"], cam.cam["theta"], cam.cam["phi"],
                         cam.cam["psi"], cam.cam["E"], cam.cam["N"], 0.0])
    np.savez_compressed(os.path.join(HERE, "utm_expected.npz"), xy=xy, EN=EN, cam=camvec)
    print("utm_expected.npz", EN[:2])


LK_EXTREME = [
    dict(winSize=(3, 3), maxLevel=0, criteria=(3, 30, 0.01)),
    dict(winSize=(63, 63), maxLevel=7, criteria=(3, 30, 0.01)),
    dict(winSize=(5, 61), maxLevel=2, criteria=(3, 200, 0.0)),        # maxCount clamps to 100, eps 0
    dict(winSize=(33, 32), maxLevel=3, criteria=(3, 0, 20.0)),        # maxCount 0: no iteration; eps clamps to 10
    dict(winSize=(21, 21), maxLevel=3, criteria=(0, 5, 0.5)),         # neither bit set: defaults 30 / 0.01
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


def make_extreme():
    """Parameter extremes and awkward shapes (width not a multiple of 4 / 16, tiny images)."""
    rng = np.random.default_rng(77)
    out = {}
    h, w = 150, 211
    base = syn.base_texture(h, w, 8, scene="texture")
    f0 = syn.frame_gray(base, 0, noise_sigma=1.0, seed=8).numpy()
    f1 = syn.frame_gray(base, 1, vx=1.5, vy=0.75, noise_sigma=1.0, seed=8).numpy()
    out["f0"], out["f1"] = f0, f1
    pts = lk_points(h, w, f0, rng)
    out["lk_pts"] = pts
    for li, lp in enumerate(LK_EXTREME):
        p1, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **lp)
        out["lk%d_p1" % li], out["lk%d_st" % li] = p1, st
        out["lk%d_err" % li] = np.where(st == 1, err, 0).astype(np.float32)
    mask = np.zeros((h, w), np.uint8)
    mask[h // 2, w // 2] = 255
    mask[10:40, 15:90] = 1                                  # any non-zero value allows
    out["mask"] = mask
    for gi, gp in enumerate(GFTT_EXTREME):
        for mi, m in enumerate([None, mask]):
            p = cv2.goodFeaturesToTrack(f0, mask=m, **gp)
            out["gftt%d_m%d" % (gi, mi)] = np.zeros((0, 1, 2), np.float32) if p is None else p
    for (hh, ww) in [(24, 24), (9, 130), (130, 9), (40, 37)]:
        a = rng.integers(0, 256, (hh, ww), dtype=np.uint8)
        b = np.roll(a, 1, axis=1)
        out["tiny_%dx%d_a" % (hh, ww)], out["tiny_%dx%d_b" % (hh, ww)] = a, b
        tp = (rng.random((25, 1, 2)) * [ww - 1, hh - 1]).astype(np.float32)
        out["tiny_%dx%d_pts" % (hh, ww)] = tp
        p1, st, err = cv2.calcOpticalFlowPyrLK(a, b, tp, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
        out["tiny_%dx%d_p1" % (hh, ww)], out["tiny_%dx%d_st" % (hh, ww)] = p1, st
        p = cv2.goodFeaturesToTrack(a, 0, 0.05, 3, blockSize=3)
        out["tiny_%dx%d_gftt" % (hh, ww)] = np.zeros((0, 1, 2), np.float32) if p is None else p
        rgb = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        out["tiny_%dx%d_rgb" % (hh, ww)] = rgb
        out["tiny_%dx%d_gray" % (hh, ww)] = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)
        ml, pyr = cv2.buildOpticalFlowPyramid(a, (5, 5), 3, withDerivatives=True)
        out["tiny_%dx%d_ml" % (hh, ww)] = np.int32(ml)
        for l in range(ml + 1):
            out["tiny_%dx%d_L%d" % (hh, ww, l)] = np.ascontiguousarray(pyr[2 * l])
            out["tiny_%dx%d_D%d" % (hh, ww, l)] = np.ascontiguousarray(pyr[2 * l + 1])
    np.savez_compressed(os.path.join(HERE, "kat_extreme.npz"), **out)
    print("kat_extreme.npz", len(out))


def make_s2():
    """Run the UNMODIFIED reference s2_cam_to_utm.cam_to_utm on synthetic track files (pd.read_excel is patched to
    return the parameter table because openpyxl is not installed; everything else is the reference's own code)."""
    import datetime as dt
    import tempfile
    import pandas as pd
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.path", "shapefile"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, "/root/reference")
    import s2_cam_to_utm as s2
    W, H = 6000, 4000
    row = dict(camera="cam1", start_day=20190701, end_day=20190801, image_width=W, image_height=H, sensor_width=22.3,
               easting=377280.39, northing=6525846.97, elevation=261.3, antenna_height=1.6, theta=300.0, phi=5.0,
               psi=-1.0, sigma=18.0, crop_left=250, crop_right=0, crop_top=400, crop_bottom=0, tracking_interval=60,
               mask="mask.shp")
    params = pd.DataFrame([row])
    real_read_excel = pd.read_excel
    pd.read_excel = lambda *a, **k: params
    rng = np.random.default_rng(21)
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "output", "cam1", "oblique", "20190724")
    tgt = os.path.join(tmp, "output", "cam1", "utm")
    os.makedirs(src); os.makedirs(tgt); os.makedirs(os.path.join(tmp, "data"))
    minutes = pd.date_range("2019-07-24 12:00", "2019-07-24 14:00", freq="min")
    tides = pd.DataFrame({"date": minutes, "depth_tide_ellipsoid": 0.3 + 0.01 * np.arange(len(minutes))})
    tide_path = os.path.join(tmp, "data", "tide_2019.pickle")
    tides.to_pickle(tide_path)
    out = {}
    stamps = ["20190724-125700", "20190724-125900", "20190724-130100", "20190724-130300"]
    for si, stamp in enumerate(stamps):
        M = 120
        p0 = (rng.random((M, 2)) * [5200, 1700] + [100, 1500]).astype(np.float32)       # water, below the horizon
        step = rng.normal(0, 1.0, (M, 1, 2)) + rng.normal(0, 0.25, (M, 2, 2))            # ~0.2 m/px * px/60 s
        step[::7] *= 12.0                                                                 # too fast (criterion 1 / 2)
        step[3::11, 1] *= -1.0                                                            # direction flips (criterion 3)
        step[5::13] = 0.0                                                                 # no motion at all
        tr = np.concatenate([p0[:, None, :], p0[:, None, :] + np.cumsum(step, 1)], 1).astype(np.float32)
        q = np.abs(rng.normal(0, 0.2, (M, 2))).astype(np.float32)
        np.savez(os.path.join(src, "%s_120sec_at_60sec_tracks.npz" % stamp), tracks=tr, trackquality=q)
        out["in%d_tracks" % si] = tr
    np.savez(os.path.join(src, "20190724-130500_120sec_at_60sec_tracks.npz"), tracks=[], trackquality=[])   # empty group
    stamps.append("20190724-130500")
    s2.cam_to_utm((src, tgt, "cam1", 1.7, 0.0, 2.5, 60, 0.1, os.path.join(tmp, "parameter_file.xlsx"), tide_path))
    pd.read_excel = real_read_excel
    files = sorted(os.listdir(tgt))
    out["stamps"] = np.array(stamps)
    out["out_files"] = np.array(files)
    for fi, f in enumerate(files):
        z = np.load(os.path.join(tgt, f))
        for k in ("x", "y", "u", "v", "speed", "time"):
            out["out%d_%s" % (fi, k)] = z[k]
    out["params_keys"] = np.array(list(row.keys()))
    out["params_vals"] = np.array([str(v) for v in row.values()])
    out["tide_minutes"] = np.array([str(m) for m in minutes])
    out["tide_values"] = tides["depth_tide_ellipsoid"].to_numpy()
    np.savez_compressed(os.path.join(HERE, "s2_expected.npz"), **out)
    print("s2_expected.npz", files, {k: v.shape for k, v in out.items() if k.startswith("out") and k.endswith("_x")})


if __name__ == "__main__":
    if "--only-s2" in sys.argv:
        make_s2()
    elif "--only-extreme" in sys.argv:
        make_extreme()
    else:
        torch.manual_seed(0)
        cv2.setNumThreads(1)
        make_kat("kat_texture.npz", 240, 320, 3, "texture")
        make_kat("kat_iceberg.npz", 201, 333, 4, "iceberg")
        make_edge()
        make_sequence()
        make_utm()
        make_extreme()
        make_s2()
