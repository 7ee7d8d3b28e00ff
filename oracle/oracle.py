"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.

ctypes front end of oracle/ibt_oracle.c, the plain-C restatement of the OpenCV-internal
arithmetic that the reference's tracking loop resolves to
(/root/reference/s1_lucaskanade_tracking.py:296-450,
/root/reference/s0_1_test_lucaskanade_tracking.py:57-181; OpenCV pinned 4.9.0 / 4.10.0
in environment.yml:254 / s0_1.yml:199, wheel here 4.13.0.92).  Function names and
signatures follow cv2's so parity tests read like calls of the reference.

Pinning: the reference ships no tests or golden vectors for this path; the oracle is
pinned against outputs of the cv2 wheel recorded in tests/golden/ (make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libibt_oracle.so")

COLOR_BGR2GRAY = 6
TERM_CRITERIA_COUNT = 1
TERM_CRITERIA_EPS = 2
OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_LK_GET_MIN_EIGENVALS = 8


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, "ibt_oracle.c"), os.path.join(_HERE, "jpeg_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_pyramid_levels.restype = C.c_int
        _lib.orc_gftt_select.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def cvtColor(src, code=COLOR_BGR2GRAY, coeffset=0):
    """A.1; cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) at s1:311."""
    assert code == COLOR_BGR2GRAY
    src = np.ascontiguousarray(src)
    assert src.dtype == np.uint8 and src.ndim == 3 and src.shape[2] in (3, 4)
    dst = np.empty(src.shape[:2], np.uint8)
    lib().orc_gray_u8(_p(src), C.c_int64(dst.size), C.c_int(src.shape[2]), C.c_int(coeffset), _p(dst))
    return dst


def pyrDown(src):
    """A.2."""
    src = np.ascontiguousarray(src)
    assert src.dtype == np.uint8 and src.ndim == 2
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown_u8(_p(src), C.c_int(h), C.c_int(w), _p(dst))
    return dst


def scharr_deriv(src):
    """A.4; (h,w,2) int16, channel 0 = dx, 1 = dy."""
    src = np.ascontiguousarray(src)
    h, w = src.shape
    dst = np.empty((h, w, 2), np.int16)
    lib().orc_scharr_i16(_p(src), C.c_int(h), C.c_int(w), _p(dst))
    return dst


def pyramid_sizes(h, w, winSize, maxLevel):
    sizes = (C.c_int * (2 * (maxLevel + 1)))()
    ml = lib().orc_pyramid_levels(C.c_int(h), C.c_int(w), C.c_int(winSize[0]), C.c_int(winSize[1]),
                                  C.c_int(maxLevel), sizes)
    return ml, [(sizes[2 * i], sizes[2 * i + 1]) for i in range(ml + 1)]


def buildOpticalFlowPyramid(img, winSize, maxLevel, withDerivatives=True):
    """A.3; returns (maxLevelOut, [L0, D0, L1, D1, ...]) like cv2 (unpadded arrays)."""
    img = np.ascontiguousarray(img)
    ml, sizes = pyramid_sizes(img.shape[0], img.shape[1], winSize, maxLevel)
    out = []
    lvl = img
    for l in range(ml + 1):
        if l > 0:
            lvl = pyrDown(lvl)
        out.append(lvl)
        if withDerivatives:
            out.append(scharr_deriv(lvl))
    return ml, out


def _criteria(criteria):
    typ, cnt, eps = criteria
    if not (typ & TERM_CRITERIA_COUNT):
        cnt = 30
    if not (typ & TERM_CRITERIA_EPS):
        eps = 0.01
    return int(cnt), float(eps)


def calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts=None, winSize=(21, 21), maxLevel=3,
                         criteria=(TERM_CRITERIA_COUNT | TERM_CRITERIA_EPS, 30, 0.01), flags=0,
                         minEigThreshold=1e-4, return_iters=False):
    """A.5; same call as s1:323 / s1:326.  With return_iters also returns the (N, L+1)
    int32 per-level iteration counts (column l = pyramid level l)."""
    prevPts = np.asarray(prevPts)
    assert prevPts.dtype == np.float32, "prevPts must be float32 (cv2 asserts the same)"
    shp = prevPts.shape
    pts = np.ascontiguousarray(prevPts.reshape(-1, 2))
    n = pts.shape[0]
    if n == 0:
        return (None, None, None) + ((None,) if return_iters else ())
    prevImg, nextImg = np.asarray(prevImg), np.asarray(nextImg)
    assert prevImg.shape == nextImg.shape and prevImg.ndim in (2, 3)
    if prevImg.ndim == 3 and prevImg.shape[2] == 1:
        prevImg, nextImg = prevImg[..., 0], nextImg[..., 0]
    # multi-channel images (cv2 accepts 3 channels; the reference always converts to gray first, s1:311): OpenCV builds the
    # pyramid and the Scharr planes per channel and sums every window quantity over the channels as well
    chI = [prevImg] if prevImg.ndim == 2 else [np.ascontiguousarray(prevImg[..., c]) for c in range(prevImg.shape[2])]
    chJ = [nextImg] if nextImg.ndim == 2 else [np.ascontiguousarray(nextImg[..., c]) for c in range(nextImg.shape[2])]
    cn = len(chI)
    lvI, dvI, lvJ = [], [], []
    for c in range(cn):
        ml, pI = buildOpticalFlowPyramid(chI[c], winSize, maxLevel, True)
        ml2, pJ = buildOpticalFlowPyramid(chJ[c], winSize, maxLevel, False)
        assert ml == ml2
        lvI += [pI[2 * l] for l in range(ml + 1)]
        dvI += [pI[2 * l + 1] for l in range(ml + 1)]
        lvJ += list(pJ)
    L = ml + 1
    arrI = (C.c_void_p * (L * cn))(*[a.ctypes.data for a in lvI])
    arrJ = (C.c_void_p * (L * cn))(*[a.ctypes.data for a in lvJ])
    arrD = (C.c_void_p * (L * cn))(*[a.ctypes.data for a in dvI])
    hs = (C.c_int * L)(*[a.shape[0] for a in lvI[:L]])
    ws = (C.c_int * L)(*[a.shape[1] for a in lvI[:L]])
    use_init = bool(flags & OPTFLOW_USE_INITIAL_FLOW)
    if use_init:
        nxt = np.ascontiguousarray(np.asarray(nextPts, np.float32).reshape(-1, 2)).copy()
    else:
        nxt = np.zeros((n, 2), np.float32)
    st = np.empty(n, np.uint8)
    err = np.empty(n, np.float32)
    iters = np.zeros((n, L), np.int32)
    cnt, eps = _criteria(criteria)
    lib().orc_lk_cn(arrI, arrJ, arrD, C.c_int(cn), hs, ws, C.c_int(ml), _p(pts), _p(nxt), C.c_int(n),
                 C.c_int(winSize[0]), C.c_int(winSize[1]), C.c_int(cnt), C.c_double(eps),
                 C.c_double(minEigThreshold), C.c_int(use_init),
                 C.c_int(bool(flags & OPTFLOW_LK_GET_MIN_EIGENVALS)), _p(st), _p(err), _p(iters))
    res = (nxt.reshape(shp), st.reshape(n, 1), err.reshape(n, 1))
    return res + ((iters,) if return_iters else ())


def cornerMinEigenVal(img, blockSize, ksize=3):
    """A.6 steps 1-3."""
    assert ksize == 3
    img = np.ascontiguousarray(img)
    h, w = img.shape
    eig = np.empty((h, w), np.float32)
    lib().orc_mineig_f32(_p(img), C.c_int(h), C.c_int(w), C.c_int(blockSize), _p(eig))
    return eig


def cornerHarris(img, blockSize, ksize=3, k=0.04):
    """cv2.cornerHarris(img, blockSize, 3, k): the response goodFeaturesToTrack(useHarrisDetector=True) ranks (A.6 step 3')."""
    assert ksize == 3
    img = np.ascontiguousarray(img)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    lib().orc_harris_f32(_p(img), C.c_int(h), C.c_int(w), C.c_int(blockSize), C.c_double(float(k)), _p(out))
    return out


def gftt_select(eig, mask, maxCorners, qualityLevel, minDistance):
    """A.6 steps 4-8 on a given eig map (not modified)."""
    eig = np.array(eig, np.float32, copy=True, order="C")
    h, w = eig.shape
    if mask is not None:
        mask = np.ascontiguousarray(mask)
        assert mask.shape == eig.shape and mask.dtype == np.uint8
    cap = max(1024, eig.size // 4)
    out = np.empty((cap, 2), np.float32)
    n = lib().orc_gftt_select(_p(eig), _p(mask) if mask is not None else None, C.c_int(h), C.c_int(w),
                              C.c_int(int(maxCorners)), C.c_double(qualityLevel), C.c_double(minDistance),
                              _p(out), C.c_int(cap))
    if n == 0:
        return None
    return out[:n].reshape(-1, 1, 2).copy()


def goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, mask=None, blockSize=3,
                        useHarrisDetector=False, k=0.04):
    """A.6; same call as s1:437.  useHarrisDetector=True ranks cornerHarris(image, blockSize, 3, k) instead of lambda_min (the
    reference never sets it; cv2's signature has it)."""
    resp = cornerHarris(image, blockSize, 3, k) if useHarrisDetector else cornerMinEigenVal(image, blockSize)
    return gftt_select(resp, mask, maxCorners, qualityLevel, minDistance)


def photo_to_utm(xy, cam):
    """A.9; Camera.photocords_cropped_to_uncropped + Camera.photo_to_utm
    (imports/camtools.py:414-421, 286-332).  cam = 12 float64, see ibt_oracle.c."""
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    cam = np.ascontiguousarray(cam, np.float64)
    out = np.empty_like(xy)
    lib().orc_photo_to_utm(_p(xy), C.c_int64(xy.shape[0]), _p(cam), _p(out))
    return out


def fb_check(p0, p0r, threshold=1.0):
    """A.7 (s1:329-333): FB distance and validity; status/err are not consulted."""
    diff = np.abs(np.asarray(p0, np.float32) - np.asarray(p0r, np.float32)).reshape(-1, 2)
    dist = np.hypot(diff[:, 0], diff[:, 1])
    return dist, dist < threshold


def track_velocities(tracks, cam, tracking_interval, min_speed, max_speed, max_speedfactor, max_angle, speed_threshold):
    """Per-track body of cam_to_utm, s2_cam_to_utm.py:243-343, restated with the reference's own Python loops and list
    max()/min() calls (small inputs only).  Returns EN (M,T+1,2), uv (M,T,2), speed (M,T) float64 and keep (M,) bool."""
    import warnings
    tracks = np.asarray(tracks, np.float32)
    M, T = tracks.shape[0], tracks.shape[1] - 1
    EN = photo_to_utm(tracks.reshape(-1, 2).astype(np.float64), cam).reshape(M, T + 1, 2)
    uv = np.zeros((M, T, 2)); sp = np.zeros((M, T)); keep = np.ones(M, bool)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in range(M):
            usub, vsub, ssub = [], [], []
            for i in range(1, T + 1):
                xm = (EN[m, i, 0] - EN[m, i - 1, 0]) / float(tracking_interval)
                ym = (EN[m, i, 1] - EN[m, i - 1, 1]) / float(tracking_interval)
                usub.append(xm); vsub.append(ym); ssub.append(np.hypot(xm, ym))
            uv[m, :, 0], uv[m, :, 1], sp[m] = usub, vsub, ssub
            if (np.mean(ssub) < min_speed) or (max(ssub) > max_speed):
                keep[m] = False
                continue
            if max(ssub) > speed_threshold and T >= 2:
                ang, rat = [], []
                for c1, c2 in zip(range(0, T - 1), range(1, T)):
                    dot = usub[c1] * usub[c2] + vsub[c1] * vsub[c2]
                    mag1, mag2 = np.hypot(usub[c1], vsub[c1]), np.hypot(usub[c2], vsub[c2])
                    ang.append(abs(np.degrees(np.arccos(dot / (mag1 * mag2)))))
                    rat.append(max([ssub[c1], ssub[c2]]) / min([ssub[c1], ssub[c2]]))
                if max(rat) > max_speedfactor or max(ang) > max_angle:
                    keep[m] = False
    return EN, uv, sp, keep


# ---- JPEG (oracle/jpeg_oracle.c): np.array(Image.open(image)) at s1:310 -------------------------------------------
class JpegUnsupported(ValueError):
    pass


def imread_jpeg(data):
    """bytes of a baseline JPEG -> what np.array(PIL.Image.open(...)) holds: (H,W,3) RGB u8, or (H,W) u8 for mode L."""
    buf = np.frombuffer(data, dtype=np.uint8)
    L = lib()
    w, h, nc = C.c_int(), C.c_int(), C.c_int()
    rc = L.orc_jpeg_info(_p(buf), C.c_long(buf.size), C.byref(w), C.byref(h), C.byref(nc))
    if rc == -5:
        raise JpegUnsupported("oracle: JPEG variant not handled")
    if rc != 0:
        raise ValueError("oracle: not a decodable JPEG (%d)" % rc)
    out = np.zeros((h.value, w.value, 3) if nc.value == 3 else (h.value, w.value), np.uint8)
    rc = L.orc_jpeg_decode(_p(buf), C.c_long(buf.size), _p(out))
    if rc != 0:
        raise ValueError("oracle: JPEG decode failed (%d)" % rc)
    return out


def jpeg_recompress(rgb, quality=75, subsampling="4:2:0"):
    """What np.array(Image.open(f)) holds after Image.fromarray(rgb).save(f, quality=, subsampling=): the save-and-reopen
    round trip of the reference's cropping pre-pass (imports/camtools.py:80,102,232 -> s1:310); Pillow defaults = 75, 4:2:0."""
    hs, vs = {"4:4:4": (1, 1), "4:2:2": (2, 1), "4:2:0": (2, 2)}[subsampling]
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    H, W = rgb.shape[:2]
    out = np.zeros((H, W, 3), np.uint8)
    rc = lib().orc_jpeg_recompress(_p(rgb), C.c_long(W * 3), W, H, int(quality), hs, vs, _p(out))
    if rc != 0:
        raise ValueError("oracle: recompress failed (%d)" % rc)
    return out


# ---- s3 gridding (numpy restatement of s3_utm_to_gridded_utm.py:391-421 + imports/tracking_misc.py:15-58) -----------------
def contains_points(poly, points):
    """matplotlib.path.Path(poly).contains_points(points), radius 0: the even-odd "crossings" rule of matplotlib's
    src/_path.h point_in_path_impl (polygon implicitly closed), elementwise fp64."""
    v = np.asarray(poly, np.float64).reshape(-1, 2)
    p = np.asarray(points, np.float64).reshape(-1, 2)
    tx, ty = p[:, 0], p[:, 1]
    inside = np.zeros(len(p), bool)
    for e in range(len(v)):
        v0, v1 = v[e], v[(e + 1) % len(v)]
        yflag0, yflag1 = v0[1] >= ty, v1[1] >= ty
        cross = ((v1[1] - ty) * (v0[0] - v1[0]) >= (v1[0] - tx) * (v0[1] - v1[1])) == yflag1
        inside ^= (yflag0 != yflag1) & cross
    return inside


def grid_cells(fjord_x, fjord_y, spacing):
    """tracking_misc.py:25-58: (polygons, centerpoints, indices, rows, cols) of the cells whose centre is in the fjord."""
    import math
    topleft = [min(fjord_x), max(fjord_y)]
    cols = int(math.ceil((max(fjord_x) - min(fjord_x)) / spacing))
    rows = int(math.ceil((max(fjord_y) - min(fjord_y)) / spacing))
    fj = np.vstack((fjord_x, fjord_y)).T
    polys, cents, idx = [], [], []
    for i in range(cols):
        for j in range(rows):
            x, y = topleft[0] + i * spacing, topleft[1] - j * spacing
            point = [x + 0.5 * spacing, y - 0.5 * spacing]
            if contains_points(fj, [point])[0]:
                polys.append([(x, y), (x + spacing, y), (x + spacing, y - spacing), (x, y - spacing)])
                cents.append(point)
                idx.append([i, j])
    return polys, cents, idx, rows, cols


def grid_bin(polys, x, y, u, v):
    """s3:391-411 per cell polygon: (count, np.sum(u_selected), np.sum(v_selected))."""
    points = np.vstack((x, y)).T
    disp = np.vstack((u, v)).T
    out = []
    for poly in polys:
        sel = contains_points(poly, points)
        d = disp[sel == 1]
        out.append((len(d), np.sum(d[:, 0]), np.sum(d[:, 1])))
    return out
