"""End to end from JPEG files on disk (what s1_lucaskanade_tracking.py actually reads, s1:307-311): 24 MP frames saved
by Pillow like the reference's cropping step, tracked with re-seeding every track_len frames.
  gpu     : loader="gpu"  -- file bytes -> csrc/jpeg.cu -> gray -> pyramids -> LK fwd/bwd/FB (+ GFTT, compaction, D2H)
  pillow  : loader=load_image in a thread pool (the previous host path: Pillow decode, 72 MB upload per frame)
  cpu ref : the reference's own loop body (Pillow decode + cv2.cvtColor + 2 x cv2.calcOpticalFlowPyrLK) on a few pairs."""
import io, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from iceberg_tracking_code_b200 import build, synthetic as syn
from iceberg_tracking_code_b200.tracking import track_sequence, SequenceTracker, load_image

build.build()
H, W, NF, T = 4000, 6000, int(os.environ.get("NF", 37)), 2
maxc = int(os.environ.get("MAXC", 20000))
tmp = tempfile.mkdtemp(prefix="ibt_seq_")
base = syn.base_texture(H, W, 7, device="cuda", scene=os.environ.get("SCENE", "texture"))
files = []
for t in range(NF):
    rgb = syn.frame_rgb(base, t % 12, seed=7 + t).cpu().numpy()          # (the scene wraps every 12 frames: stays inside the margin)
    f = os.path.join(tmp, "20190724-13%02d00.jpg" % t)
    Image.fromarray(rgb).save(f)
    files.append(f)
del base
mb = sum(os.path.getsize(f) for f in files) / NF / 1e6
gp = dict(maxCorners=maxc, qualityLevel=0.007, minDistance=10, blockSize=10)
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
out = {"frames": NF, "jpeg_mb_per_frame": mb}
ref_res = None
for name, kw in (("gpu", dict(loader="gpu")),
                 ("pillow", dict(loader=load_image, decode_workers=os.cpu_count()))):
    trk = SequenceTracker(gp, lp)
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = track_sequence(files, None, T, 60, tracker=trk, save=False, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if ref_res is None:
        ref_res = res
    else:
        assert all(np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) for a, b in zip(ref_res, res)), name
    out[name + "_ms_per_frame"] = best / (NF - 1) * 1e3
    out[name + "_frames_per_s"] = (NF - 1) / best
out["tracks_per_group"] = int(ref_res[0][2].shape[0])
try:
    import cv2
    cv2.setNumThreads(os.cpu_count())
    p0 = ref_res[0][2][:, 0, :].reshape(-1, 1, 2).astype(np.float32)
    t0 = time.perf_counter()
    prev = cv2.cvtColor(np.array(Image.open(files[0])), cv2.COLOR_BGR2GRAY)
    n = 3
    for i in range(1, n + 1):
        cur = cv2.cvtColor(np.array(Image.open(files[i])), cv2.COLOR_BGR2GRAY)
        p1, st, err = cv2.calcOpticalFlowPyrLK(prev, cur, p0, None, **lp)
        p0r, st, err = cv2.calcOpticalFlowPyrLK(cur, prev, p1, None, **lp)
        prev = cur
    out["cpu_ref_ms_per_frame"] = (time.perf_counter() - t0) / (n + 1) * 1e3
    out["cpu_threads"] = os.cpu_count()
except ImportError:
    pass
print(json.dumps(out))
