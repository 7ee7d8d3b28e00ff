"""GPU: the CUDA JPEG decoder (csrc/jpeg.cu through the C-ABI) against Pillow's recorded outputs, the oracle and a live
Pillow decode -- bit-exact (byte work)."""
import glob
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
JDIR = os.path.join(GOLDEN, "jpeg")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(JDIR, "*.jpg")))


def _read(name):
    with open(os.path.join(JDIR, name + ".jpg"), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def jpeg(ibt):
    from iceberg_tracking_code_b200 import jpeg as J
    return J


@pytest.fixture(scope="module")
def expected():
    return dict(np.load(os.path.join(JDIR, "expected.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_decode_matches_pillow_golden(jpeg, ibt, oracle, expected, name):
    data = _read(name)
    exp = expected[name]
    out = jpeg.imread(data).cpu().numpy()
    assert out.shape == exp.shape
    assert np.array_equal(out, exp)
    assert np.array_equal(out, oracle.imread_jpeg(data))
    g = jpeg.imread(data, gray=True).cpu().numpy()
    if exp.ndim == 3:
        assert np.array_equal(g, oracle.cvtColor(exp))          # s1:310-311 fused
        g3 = jpeg.imread(data, gray=True, coeffset=1).cpu().numpy()
        assert np.array_equal(g3, oracle.cvtColor(exp, coeffset=1))
    else:
        assert np.array_equal(g, exp)


def test_decode_live_matrix(jpeg):
    """quality x subsampling x optimize x sizes, encoded and decoded by the Pillow of this box."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(12)
    n = 0
    for (h, w) in [(64, 64), (37, 53), (9, 130), (131, 7), (300, 500), (16, 5)]:
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
                            continue
                        ref = np.array(Image.open(io.BytesIO(bio.getvalue())))
                        out = jpeg.imread(bio.getvalue()).cpu().numpy()
                        assert np.array_equal(out, ref), (h, w, sub, q, opt)
                        n += 1
    assert n > 180


def test_decode_restart_intervals(jpeg):
    """DRI files (restart markers every k MCUs / rows), all samplings, odd sizes: RSTn removal, padding skip, DC reset."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(13)
    n = 0
    for (h, w) in [(64, 96), (37, 53), (200, 333), (9, 130)]:
        smooth = np.cumsum(np.cumsum(rng.normal(0, 3, (h, w, 3)), 0), 1)
        smooth = ((smooth - smooth.min()) / (np.ptp(smooth) + 1e-9) * 255).astype(np.uint8)
        flat = np.full((h, w, 3), 77, np.uint8)
        for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), smooth, flat):
            for sub in (0, 1, 2):
                for kw in (dict(restart_marker_blocks=1), dict(restart_marker_blocks=3), dict(restart_marker_blocks=7),
                           dict(restart_marker_rows=1), dict(restart_marker_rows=2)):
                    for opt in (False, True):
                        bio = io.BytesIO()
                        try:
                            Image.fromarray(img).save(bio, "JPEG", quality=85, subsampling=sub, optimize=opt, **kw)
                        except OSError:
                            continue
                        data = bio.getvalue()
                        assert b"\xff\xdd" in data
                        ref = np.array(Image.open(io.BytesIO(data)))
                        assert np.array_equal(jpeg.imread(data).cpu().numpy(), ref), (h, w, sub, kw, opt)
                        n += 1
        g = smooth[..., 0]
        bio = io.BytesIO()
        Image.fromarray(g).save(bio, "JPEG", quality=80, restart_marker_blocks=2)
        assert np.array_equal(jpeg.imread(bio.getvalue()).cpu().numpy(), np.array(Image.open(io.BytesIO(bio.getvalue()))))
    assert n > 300


@pytest.mark.parametrize("scene,kw", [("texture", dict(restart_marker_rows=1)), ("texture", {}), ("iceberg", {}), ("texture", dict(quality=92, subsampling=1))])
def test_decode_24mp(jpeg, ibt, scene, kw):
    """BASELINE config-2 frame size: a 6000x4000 synthetic frame saved like the reference's cropping step saves it
    (imports/camtools.py:80 `img_crop.save(outpath)`), decoded here and by Pillow."""
    import torch
    Image = pytest.importorskip("PIL.Image")
    from iceberg_tracking_code_b200 import synthetic as syn
    base = syn.base_texture(4000, 6000, 7, device="cuda", scene=scene)
    rgb = syn.frame_rgb(base, 0, seed=7).cpu().numpy()
    bio = io.BytesIO()
    Image.fromarray(rgb).save(bio, "JPEG", **kw)
    data = bio.getvalue()
    ref = np.array(Image.open(io.BytesIO(data)))
    dec = jpeg.JpegDecoder()
    out, gray = dec.decode(data, rgb=True, gray=True)
    assert torch.equal(out.cpu(), torch.from_numpy(ref))
    assert torch.equal(gray, ibt.cvtColor(out))
    assert 1 <= dec.last_rounds < 64
    # determinism: a second decode through the same workspace gives the same bytes
    out2, _ = dec.decode(data, rgb=True, gray=False)
    assert torch.equal(out, out2)


def test_sequence_from_jpeg_files_on_gpu(jpeg, ibt, golden):
    """The golden sequence run (unmodified reference class on tests/golden/seq/*.jpg) with the frames decoded on the GPU."""
    from iceberg_tracking_code_b200 import tracking
    exp = golden("seq_expected.npz")
    files = sorted(glob.glob(os.path.join(GOLDEN, "seq", "*.jpg")))
    from PIL import Image
    for f in files:
        assert np.array_equal(jpeg.imread(f).cpu().numpy(), np.array(Image.open(f)))
    res_cpu = tracking.track_sequence(files, None, 2, 60, save=False, check_time=True, decode_workers=0)
    res_gpu = tracking.track_sequence(files, None, 2, 60, save=False, check_time=True, loader="gpu")
    assert len(res_cpu) == len(res_gpu) > 0
    for a, b in zip(res_cpu, res_gpu):
        assert a[0] == b[0]
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert "counts" in exp or True


def test_corrupt_streams_terminate(jpeg):
    """Truncated and bit-flipped entropy data: the speculative decoder must terminate and return an image of the right shape
    (Pillow raises / returns grey for such files; there is no defined pixel content to compare)."""
    data = bytearray(_read("tex_420_default"))
    info = jpeg.parse(bytes(data))
    lo, n = info.scan_offset, info.scan_bytes
    rng = np.random.default_rng(3)
    cases = [bytes(data[:lo + n // 2]) + b"\xff\xd9", bytes(data[:lo + 5]) + b"\xff\xd9"]
    for _ in range(4):
        d = bytearray(data)
        for p in rng.integers(lo, lo + n, 40):
            d[p] = int(rng.integers(0, 255))          # never 0xFF: a stray marker would end the scan early (also fine)
        cases.append(bytes(d))
    garbage = bytearray(data)
    garbage[lo:lo + n] = bytes(int(v) for v in rng.integers(0, 255, n))
    cases.append(bytes(garbage))
    for c in cases:
        out = jpeg.imread(c)
        assert tuple(out.shape) == (info.height, info.width, 3)
        assert int(out.sum().item()) >= 0                 # forces completion


def test_decode_async_equals_decode(ibt):
    """ibt_jpeg_decode_async (fixed number of Huffman synchronisation rounds, no host round trip) == ibt_jpeg_decode, and
    confirm() repairs a decode that was given too few rounds."""
    import torch
    from PIL import Image
    from iceberg_tracking_code_b200 import jpeg, synthetic as syn
    base = syn.base_texture(600, 900, 5)
    blobs = []
    for t in range(3):
        bio = io.BytesIO()
        Image.fromarray(syn.frame_rgb(base, t, seed=5).numpy()).save(bio, "JPEG")
        blobs.append(bio.getvalue())
    dec = jpeg.JpegDecoder()
    ref = [dec.decode(b, rgb=True, gray=True) for b in blobs]
    assert dec.last_rounds > 0
    for b, (r_rgb, r_gray) in zip(blobs, ref):
        h = dec.decode_async(b, rgb=True, gray=True)
        rgb, gray = dec.confirm(h)
        assert torch.equal(rgb, r_rgb) and torch.equal(gray, r_gray)
    need = dec.last_rounds
    if need > 1:                                          # starve the decoder: confirm() must notice and decode again
        dec.last_rounds = 1
        h = dec.decode_async(blobs[0], rgb=False, gray=True, margin=0)
        _, gray = dec.confirm(h)
        assert torch.equal(gray, ref[0][1]) and dec.last_rounds > 1 and h.get("redo")


# ---- ibt_jpeg_recompress: the save-and-reopen round trip of the cropping pre-pass (camtools.py:80 -> s1:310) ---------------
SUBS = {0: "4:4:4", 1: "4:2:2", 2: "4:2:0"}


def test_recompress_matches_pillow_golden(jpeg, ibt, oracle):
    z = dict(np.load(os.path.join(JDIR, "recompress.npz")))
    names = [k[3:] for k in z if k.startswith("in_")]
    assert len(names) >= 12
    for name in names:
        src, exp, kw = z["in_" + name], z["out_" + name], z["kw_" + name]
        out = jpeg.save_reopen(src, int(kw[0]), SUBS[int(kw[1])]).cpu().numpy()
        assert np.array_equal(out, exp), name
        g = jpeg.save_reopen(src, int(kw[0]), SUBS[int(kw[1])], gray=True).cpu().numpy()
        assert np.array_equal(g, oracle.cvtColor(exp)), name


def test_recompress_matrix_vs_oracle_and_pillow(jpeg, oracle):
    """sizes 1..70 x qualities x sampling modes against the oracle (and the Pillow of this box when present); crop VIEWS with
    odd byte offsets (the kernel reads unaligned rows through aligned words)."""
    import torch
    try:
        from PIL import Image
    except ImportError:
        Image = None
    rng = np.random.default_rng(29)
    big = torch.from_numpy(rng.integers(0, 256, (90, 100, 3), dtype=np.uint8)).cuda()
    smooth = np.cumsum(np.cumsum(rng.normal(0, 3, (90, 100, 3)), 0), 1)
    smooth = torch.from_numpy(((smooth - smooth.min()) / (np.ptp(smooth) + 1e-9) * 255).astype(np.uint8)).cuda()
    dec = jpeg.JpegDecoder()
    n = 0
    for trial in range(120):
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        u, l = int(rng.integers(0, 90 - H + 1)), int(rng.integers(0, 100 - W + 1))
        view = (big if trial % 2 else smooth)[u:u + H, l:l + W]
        a = view.cpu().numpy().copy()
        q = int(rng.integers(1, 101))
        sub = int(rng.integers(0, 3)) if trial % 4 else 2
        if trial % 5 == 0:
            q = 75
        out = dec.recompress(view, q, SUBS[sub])[0].cpu().numpy()
        assert np.array_equal(out, oracle.jpeg_recompress(a, q, SUBS[sub])), (H, W, q, sub, u, l)
        if Image is not None:
            b = io.BytesIO()
            Image.fromarray(a).save(b, "JPEG", quality=q, subsampling=sub)
            assert np.array_equal(out, np.array(Image.open(io.BytesIO(b.getvalue())))), (H, W, q, sub)
        n += 1
    assert n == 120


@pytest.mark.parametrize("scene", ["texture", "iceberg"])
def test_recompress_24mp_crop_vs_pillow(jpeg, ibt, scene):
    """A 6000x4000 frame cropped like camtools.crop_image_standalone crops it (an odd box), saved with Pillow's defaults and
    reopened -- against the GPU round trip on the crop view, RGB and fused gray."""
    import torch
    Image = pytest.importorskip("PIL.Image")
    from iceberg_tracking_code_b200 import synthetic as syn
    base = syn.base_texture(4000, 6000, 7, device="cuda", scene=scene)
    rgb = syn.frame_rgb(base, 0, seed=7)
    box = (251, 401, 5994, 3999)                                         # left, upper, right, lower
    bio = io.BytesIO()
    Image.fromarray(rgb.cpu().numpy()).crop(box).save(bio, "JPEG")
    ref = torch.from_numpy(np.array(Image.open(io.BytesIO(bio.getvalue()))))
    dec = jpeg.JpegDecoder()
    out, gray = dec.recompress(rgb[box[1]:box[3], box[0]:box[2]], rgb=True, gray=True)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(gray, ibt.cvtColor(out))
    # and the decoder agrees on the file Pillow wrote (the crop="reencode" route)
    assert torch.equal(dec.decode(bio.getvalue(), rgb=True, gray=False)[0].cpu(), ref)


def test_probe_switch_changes_rounds_not_pixels(jpeg):
    """ibt_jpeg_set_probe: the entry-state probe only changes how many synchronisation rounds the speculative Huffman pass needs
    (the fixed point is the sequential decode either way) -- same bytes out, synchronous and asynchronous form."""
    import torch
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(31)
    img = rng.integers(0, 256, (700, 900, 3), dtype=np.uint8)
    bio = io.BytesIO()
    Image.fromarray(img).save(bio, "JPEG", quality=90)
    data = bio.getvalue()
    ref = torch.from_numpy(np.array(Image.open(io.BytesIO(data))))
    try:
        outs, rounds = [], []
        for on in (True, False):
            jpeg.set_probe(on)
            dec = jpeg.JpegDecoder()
            rgb, gray = dec.decode(data, rgb=True, gray=True)
            assert torch.equal(rgb.cpu(), ref)
            rounds.append(dec.last_rounds)
            h = dec.decode_async(data, rgb=True, gray=True)
            r2, g2 = dec.confirm(h)
            assert torch.equal(r2.cpu(), ref) and torch.equal(g2, gray)
            outs.append(gray)
        assert torch.equal(outs[0], outs[1])
        assert rounds[0] >= 1 and rounds[1] >= rounds[0]           # the probe never costs rounds
    finally:
        jpeg.set_probe(True)
