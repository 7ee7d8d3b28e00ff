"""Host-side logic of the path (file naming, time-gap rule, sharding, the gather) -- CPU only, incl. a 2-rank gloo run."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_npz_name_and_s2_parsing():
    from iceberg_tracking_code_b200.tracking import npz_name
    name = npz_name("/data/out/cam1/oblique/20190724/20190724-130000.jpg", 2, 60)
    assert name == "/data/out/cam1/oblique/20190724/20190724-130000_120sec_at_60sec_tracks.npz"
    base = os.path.basename(name)
    # what s2_cam_to_utm.py:177,197 recovers from the name
    assert int(base.split('_')[-2].split('sec')[0]) == 60
    assert base.split('_')[0] == "20190724-130000"
    # the reference's split('.')[0] quirk (s1:394) is kept: a dotted directory truncates the name
    assert npz_name("/d.x/20190724-130000.jpg", 2, 60).startswith("/d_")


def test_group_time_ok():
    from iceberg_tracking_code_b200.tracking import group_time_ok
    f = ["20190724-130000.jpg", "20190724-130100.jpg", "20190724-130201.jpg"]
    assert group_time_ok(f, 60)
    assert group_time_ok(["a/20190724-130000.jpg", "a/20190724-130058.jpg"], 60)
    assert not group_time_ok(["20190724-130000.jpg", "20190724-130057.jpg"], 60)
    assert not group_time_ok(["20190724-130000.jpg", "20190724-130100.jpg", "20190724-130300.jpg"], 60)   # missed photo


def test_shard_groups_partition():
    from iceberg_tracking_code_b200 import sharding as sh
    assert sh.n_groups(1440, 2) == 719 and sh.n_groups(5, 2) == 2 and sh.n_groups(5, 3) == 1 and sh.n_groups(2, 2) == 0
    for total in (0, 1, 7, 719, 720):
        for world in (1, 2, 3, 8):
            blocks = [sh.shard_groups(total, r, world) for r in range(world)]
            assert sum(n for _, n in blocks) == total
            pos = 0
            for g0, n in blocks:
                assert g0 == pos
                pos += n
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    # 1440 frames, T=2, 8 ranks: ~90 groups = 181 frames (one halo frame) per rank
    g0, n = sh.shard_groups(719, 3, 8)
    first, last = sh.frame_range(g0, n, 2)
    assert n == 90 and last - first + 1 == 181
    assert sh.frame_range(*sh.shard_groups(719, 2, 8), 2)[1] == first       # halo frame is the next block's seed frame


def test_pack_unpack_roundtrip():
    from iceberg_tracking_code_b200 import sharding as sh
    rng = np.random.default_rng(0)
    res = [(0, None, rng.random((5, 3, 2), np.float32), rng.random((5, 2), np.float32)),
           (2, None, np.zeros((0,), np.float64), np.zeros((0,), np.float64)),
           (4, None, rng.random((1, 3, 2), np.float32), rng.random((1, 2), np.float32))]
    out = sh.unpack_results(*sh.pack_results(res, 2))
    assert [o[0] for o in out] == [0, 2, 4]
    for (s, _p, t, q), (s2, t2, q2) in zip(res, out):
        assert t.shape == t2.shape and np.array_equal(t, t2) and np.array_equal(q, q2) and t.dtype == t2.dtype


_WORKER = r'''
import os, sys
import numpy as np
import pytest
import torch.distributed as dist
sys.path.insert(0, %r)
from iceberg_tracking_code_b200 import sharding as sh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
T, total = 2, 7
g0, n = sh.shard_groups(total, rank, world)
res = []
for g in range(g0, g0 + n):
    rng = np.random.default_rng(g)
    m = int(rng.integers(0, 6)) if g != 3 else 0
    if m:
        res.append((g * T, None, rng.random((m, T + 1, 2)).astype(np.float32), rng.random((m, T)).astype(np.float32)))
    else:
        res.append((g * T, None, np.zeros((0,), np.float64), np.zeros((0,), np.float64)))
out = sh.gather_results(res, T)
assert [o[0] for o in out] == [g * T for g in range(total)], out
for g, (seed, t, q) in enumerate(out):
    rng = np.random.default_rng(g)
    m = int(rng.integers(0, 6)) if g != 3 else 0
    if m:
        assert np.array_equal(t, rng.random((m, T + 1, 2)).astype(np.float32)) and np.array_equal(q, rng.random((m, T)).astype(np.float32))
    else:
        assert t.shape == (0,)
# payloads left where the collective put them (to_host=False): same groups, tensor views
out2 = sh.gather_results(res, T, to_host=False)
assert [o[0] for o in out2] == [o[0] for o in out]
for (s1, t1, q1), (s2, t2, q2) in zip(out, out2):
    assert np.array_equal(np.asarray(t1), np.asarray(t2)) and np.array_equal(np.asarray(q1), np.asarray(q2))
dist.destroy_process_group()
sys.stdout.write("rank " + str(rank) + " ok\n"); sys.stdout.flush()
'''


def test_gather_two_ranks_gloo(tmp_path):
    import socket
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    with socket.socket() as sk:                      # a free port: the driver may run several suites on one host
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_camera_fields_and_utm_params():
    """Camera mirrors camtools.Camera's dictionaries (camtools.py:126-147); checked against the golden camera vector."""
    from iceberg_tracking_code_b200.camera import Camera
    g = np.load(os.path.join(ROOT, "tests", "golden", "utm_expected.npz"))
    params = dict(image_width=6000, image_height=4000, sensor_width=22.3, easting=377280.39, northing=6525846.97,
                  elevation=261.3, antenna_height=1.6, theta=300.0, phi=5.0, psi=-1.0, sigma=18.0, crop_left=250,
                  crop_right=0, crop_top=400, crop_bottom=0, tracking_interval=60)
    cam = Camera(camname="cam1", parameters=params, tide_elevation=0.37)
    assert np.allclose(cam.utm_params(), g["cam"], rtol=0, atol=1e-9)
    assert cam.crop_box() == (250, 400, 6000, 4000)


def test_bench_reference_arm_contract(tmp_path):
    """--impl reference on a non-zero rank exits 0 without work (contract for torchrun N>1)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_crop_image_parallel_matches_pillow(tmp_path):
    """camtools.py:237-258: pool of processes, Pillow defaults, same bytes as the serial reference recipe."""
    import shutil
    from PIL import Image
    from iceberg_tracking_code_b200.camera import Camera
    seq = os.path.join(ROOT, "tests", "golden", "seq")
    names = sorted(os.listdir(seq))[:3]
    src, tgt = tmp_path / "src", tmp_path / "tgt"
    src.mkdir(); tgt.mkdir()
    for n in names:
        shutil.copy(os.path.join(seq, n), src / n)
    w, h = Image.open(src / names[0]).size
    cam = Camera(camname="cam1", parameters=dict(image_width=w, image_height=h, sensor_width=22.3, easting=0.0, northing=0.0,
                 elevation=10.0, antenna_height=0.0, theta=0.0, phi=0.0, psi=0.0, sigma=18.0, crop_left=16, crop_right=8,
                 crop_top=10, crop_bottom=20))
    cam.crop_image_parallel([str(src / n) for n in names], str(tgt), 2)
    for n in names:
        ref = tmp_path / ("ref_" + n)
        Image.open(src / n).crop((16, 10, w - 8, h - 20)).save(ref)          # crop_image_standalone, camtools.py:76-80
        assert (tgt / n).read_bytes() == ref.read_bytes()


def test_view_loader_fallbacks(tmp_path, capsys, caplog):
    """crop="view" host logic: a sound baseline JPEG goes to the GPU decoder; a truncated scan, a crop box that leaves the frame
    and a progressive file are read with Pillow the way camtools.crop_image_standalone reads them (log entry + TRUNCATED)."""
    import io
    from PIL import Image
    from iceberg_tracking_code_b200.tracking import ViewLoader

    class StubGpu:                                       # stands in for GpuJpegLoader (no GPU in this test)
        def decode(self, data):
            return "gpu"
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, (64, 96, 3), dtype=np.uint8))
    tgt = tmp_path / "cam1" / "oblique" / "20190724"
    tgt.mkdir(parents=True)
    good = tmp_path / "good.jpg"; img.save(good)
    bio = io.BytesIO(); img.save(bio, "JPEG")
    trunc = tmp_path / "trunc.jpg"; trunc.write_bytes(bio.getvalue()[: len(bio.getvalue()) * 2 // 3])
    prog = tmp_path / "prog.jpg"; img.save(prog, progressive=True)
    names = {str(tgt / "good.jpg"): str(good), str(tgt / "trunc.jpg"): str(trunc), str(tgt / "prog.jpg"): str(prog)}
    vl = ViewLoader(StubGpu(), (8, 4, 88, 60), names)
    assert vl(str(tgt / "good.jpg")) == "gpu" and vl.fallbacks == 0
    import logging
    with caplog.at_level(logging.INFO):
        a = vl(str(tgt / "trunc.jpg"))
    assert isinstance(a, np.ndarray) and a.shape == (56, 80, 3) and vl.fallbacks == 1
    assert "TRUNCATED" in capsys.readouterr().out
    # the reference's log entry (camtools.py:84-92; logging.basicConfig is a no-op under pytest's own handlers, the records are not)
    assert any(str(trunc) in r.getMessage() for r in caplog.records)
    p = vl(str(tgt / "prog.jpg"))
    assert isinstance(p, np.ndarray) and p.shape == (56, 80, 3)
    assert np.array_equal(p, np.array(Image.open(prog).crop((8, 4, 88, 60))))
    big = ViewLoader(StubGpu(), (8, 4, 120, 60), names)(str(tgt / "good.jpg"))         # box wider than the 96 px frame
    assert isinstance(big, np.ndarray) and big.shape == (56, 112, 3) and not big[:, 96 - 8:].any()   # PIL pads with zeros
    # crop="emulate": the fallback does the reference's save + reopen in memory (camtools.py:80 -> s1:310)
    class StubGpuRe(StubGpu):
        reencode = (75, "4:2:0")
    e = ViewLoader(StubGpuRe(), (8, 4, 88, 60), names)(str(tgt / "prog.jpg"))
    ref = tmp_path / "ref.jpg"
    Image.open(prog).crop((8, 4, 88, 60)).save(ref)
    assert np.array_equal(e, np.array(Image.open(ref)))


def test_cv_functions_take_cv2_positional_order():
    """The drop-in functions take their arguments in the positional order cv2's Python bindings document (first line of the
    wheel's own __doc__), so a call site written for cv2 (s1:311,323,326,437) binds the same values whether it passes them by
    position or by keyword.  buildOpticalFlowPyramid is the stated exception (its 4th positional is withDerivatives; cv2 has the
    output placeholder `pyramid` there)."""
    import inspect
    import re
    cv2 = pytest.importorskip("cv2")
    from iceberg_tracking_code_b200 import cv
    for name in ("cvtColor", "goodFeaturesToTrack", "calcOpticalFlowPyrLK", "cornerMinEigenVal", "cornerHarris", "pyrDown"):
        first = getattr(cv2, name).__doc__.splitlines()[0]
        theirs = [p for p in re.sub(r"[\[\]]", "", first[first.index("(") + 1:first.index(")")]).replace(" ", "").split(",") if p]
        ours = [p.name for p in inspect.signature(getattr(cv, name)).parameters.values()
                if p.kind == inspect.Parameter.POSITIONAL_OR_KEYWORD]
        n = min(len(ours), len(theirs))
        assert n >= 1 and ours[:n] == theirs[:n], (name, ours, theirs)
    assert cv.COLOR_BGR2GRAY == cv2.COLOR_BGR2GRAY and cv.COLOR_RGB2GRAY == cv2.COLOR_RGB2GRAY
    assert cv.COLOR_BGRA2GRAY == cv2.COLOR_BGRA2GRAY and cv.COLOR_RGBA2GRAY == cv2.COLOR_RGBA2GRAY
    assert cv.TERM_CRITERIA_EPS == cv2.TERM_CRITERIA_EPS and cv.TERM_CRITERIA_COUNT == cv2.TERM_CRITERIA_COUNT
    assert cv.OPTFLOW_USE_INITIAL_FLOW == cv2.OPTFLOW_USE_INITIAL_FLOW
    assert cv.OPTFLOW_LK_GET_MIN_EIGENVALS == cv2.OPTFLOW_LK_GET_MIN_EIGENVALS
