import io, json, os, sys, tempfile, time
sys.path.insert(0, ".")
import numpy as np, torch
from PIL import Image
from iceberg_tracking_code_b200 import build, synthetic as syn
from iceberg_tracking_code_b200.tracking import track_sequence, SequenceTracker
build.build()
H, W, NF, T = 4000, 6000, 13, 2
tmp = tempfile.mkdtemp(prefix="ibt_seq_")
base = syn.base_texture(H, W, 7, device="cuda", scene=os.environ.get("SCENE", "iceberg"))
files = []
for t in range(NF):
    f = os.path.join(tmp, "20190724-13%02d00.jpg" % t); Image.fromarray(syn.frame_rgb(base, t, seed=7).cpu().numpy()).save(f); files.append(f)
gp = dict(maxCorners=20000, qualityLevel=0.007, minDistance=10, blockSize=10)
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
trk = SequenceTracker(gp, lp)
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = track_sequence(files, None, T, 60, tracker=trk, save=False, loader="gpu")
    torch.cuda.synchronize(); print("rep", rep, "%.2f ms/frame" % ((time.perf_counter() - t0) / (NF - 1) * 1e3), flush=True)
from iceberg_tracking_code_b200.tracking import GpuJpegLoader, read_file
gpu = GpuJpegLoader(trk.device)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = track_sequence(files, None, T, 60, tracker=trk, save=False, loader=lambda p: gpu.decode(read_file(p)), decode_workers=0)
    torch.cuda.synchronize(); print("bytes path rep", rep, "%.2f ms/frame" % ((time.perf_counter() - t0) / (NF - 1) * 1e3), flush=True)
for sb in ("10", "9"):
    os.environ["IBT_JPEG_SBITS"] = sb
import subprocess
