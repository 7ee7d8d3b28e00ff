"""CPU: the JPEG oracle (oracle/jpeg_oracle.c) against Pillow's outputs (tests/golden/jpeg/, made by
make_jpeg_golden.py with the real `np.array(Image.open(f))` of s1_lucaskanade_tracking.py:310), and the host-only
marker parser of the product library (ibt_jpeg_parse; no GPU work)."""
import ctypes as C
import glob
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

JDIR = os.path.join(GOLDEN, "jpeg")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(JDIR, "*.jpg")))


def _read(name):
    with open(os.path.join(JDIR, name + ".jpg"), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def expected():
    return dict(np.load(os.path.join(JDIR, "expected.npz")))


def test_fixture_set():
    assert len(NAMES) >= 16


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_pillow_golden(oracle, expected, name):
    out = oracle.imread_jpeg(_read(name))
    assert out.shape == expected[name].shape and out.dtype == np.uint8
    assert np.array_equal(out, expected[name])


def test_oracle_matches_pillow_live(oracle):
    """Fresh encodes (quality x subsampling x optimize x odd sizes) decoded by the Pillow of this box."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(11)
    n = 0
    for (h, w) in [(64, 64), (37, 53), (9, 130), (131, 7), (16, 5)]:
        smooth = np.cumsum(np.cumsum(rng.normal(0, 3, (h, w, 3)), 0), 1)
        smooth = ((smooth - smooth.min()) / (np.ptp(smooth) + 1e-9) * 255).astype(np.uint8)
        for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), smooth):
            for sub in (0, 1, 2):
                for q in (25, 75, 98):
                    for opt in (False, True):
                        bio = io.BytesIO()
                        try:
                            Image.fromarray(img).save(bio, "JPEG", quality=q, subsampling=sub, optimize=opt)
                        except OSError:
                            continue                 # Pillow's encoder buffer is too small for some tiny optimised files
                        ref = np.array(Image.open(io.BytesIO(bio.getvalue())))
                        assert np.array_equal(oracle.imread_jpeg(bio.getvalue()), ref), (h, w, sub, q, opt)
                        n += 1
    assert n > 150


def test_oracle_rejects_progressive(oracle):
    Image = pytest.importorskip("PIL.Image")
    bio = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(bio, "JPEG", progressive=True)
    with pytest.raises(oracle.JpegUnsupported):
        oracle.imread_jpeg(bio.getvalue())


# ---- host-only entry point of the product library -------------------------------------------------------------
def _parse(data):
    from iceberg_tracking_code_b200 import _native as N
    buf = np.frombuffer(data, np.uint8)
    info = N.ibt_jpeg_info_t()
    rc = N.lib().ibt_jpeg_parse(C.c_void_p(buf.ctypes.data), buf.size, C.byref(info))
    return rc, info


@pytest.mark.parametrize("name", NAMES)
def test_parse_header(expected, name):
    from iceberg_tracking_code_b200 import _native as N
    data = _read(name)
    rc, info = _parse(data)
    assert rc == 0
    assert (info.restart_interval > 0) == name.startswith("dri_")
    exp = expected[name]
    assert (info.height, info.width) == exp.shape[:2]
    assert info.ncomp == (3 if exp.ndim == 3 else 1)
    # the entropy-coded segment ends right before the EOI marker Pillow writes at the end of the file
    assert data[info.scan_offset + info.scan_bytes:info.scan_offset + info.scan_bytes + 2] == b"\xff\xd9"
    sub = {"420": (2, 2), "422": (2, 1), "444": (1, 1)}
    for key, hv in sub.items():
        if "_" + key in name:
            assert (info.hsamp[0], info.vsamp[0]) == hv
    assert N.lib().ibt_jpeg_workspace_bytes(C.byref(info)) > 0
    q = np.ctypeslib.as_array(info.quant)
    assert q[info.qsel[0]].min() >= 1


def test_parse_rejects():
    from iceberg_tracking_code_b200 import _native as N
    assert _parse(b"\x00" * 64)[0] == N.IBT_E_INVALID
    assert _parse(b"\xff\xd8\xff\xd9")[0] == N.IBT_E_INVALID
    data = _read("tex_420_default")
    assert _parse(data[:200])[0] == N.IBT_E_INVALID                        # truncated inside the headers
    Image = pytest.importorskip("PIL.Image")
    bio = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(bio, "JPEG", progressive=True)
    assert _parse(bio.getvalue())[0] == N.IBT_E_UNSUPPORTED


# ---- the save-and-reopen round trip of the cropping pre-pass (camtools.py:80 -> s1:310) -----------------------------------
SUBS = {0: "4:4:4", 1: "4:2:2", 2: "4:2:0"}


@pytest.fixture(scope="module")
def recompress_golden():
    z = dict(np.load(os.path.join(JDIR, "recompress.npz")))
    return {k[3:]: (z[k], z["out_" + k[3:]], z["kw_" + k[3:]]) for k in z if k.startswith("in_")}


def test_oracle_recompress_matches_pillow_golden(oracle, recompress_golden):
    """recompress.npz holds Pillow's own save + open of each array (tests/golden/make_recompress_golden.py)."""
    assert len(recompress_golden) >= 12
    for name, (src, exp, kw) in recompress_golden.items():
        out = oracle.jpeg_recompress(src, int(kw[0]), SUBS[int(kw[1])])
        assert np.array_equal(out, exp), name


def test_oracle_recompress_matches_pillow_live(oracle):
    """img.save(f) + np.array(Image.open(f)) by the Pillow of this box: sizes 1..70 (every edge-replication case of the 16x16
    MCU), random qualities, the three sampling modes, noise / ramps / blocks."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(23)

    def round_trip(a, **kw):
        b = io.BytesIO()
        Image.fromarray(a).save(b, "JPEG", **kw)
        return np.array(Image.open(io.BytesIO(b.getvalue())))

    n = 0
    for trial in range(90):
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        if trial % 3 == 0:
            a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        elif trial % 3 == 1:
            yy, xx = np.mgrid[0:H, 0:W]
            a = np.stack([(xx * 3 + yy * 2) % 256, (xx * yy) % 256, (255 - xx * 4) % 256], -1).astype(np.uint8)
        else:
            a = rng.integers(0, 256, (H // 4 + 1, W // 4 + 1, 3), dtype=np.uint8).repeat(4, 0).repeat(4, 1)[:H, :W].copy()
        q = int(rng.integers(1, 101))
        for kw in ({}, dict(quality=q), dict(quality=q, subsampling=0), dict(quality=q, subsampling=1), dict(quality=100)):
            out = oracle.jpeg_recompress(a, kw.get("quality", 75), SUBS[kw.get("subsampling", 2)])
            assert np.array_equal(out, round_trip(a, **kw)), (H, W, kw)
            n += 1
    assert n == 450
