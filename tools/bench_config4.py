"""BASELINE config 4 at full size: 200 000 grid-seeded points on a 24 MP pair, winSize 31, maxLevel 5, criteria (3, 30, 0.01),
forward + backward + FB in one launch.  (Parity of this configuration at full size against the CPU oracle:
tests/test_gpu_sequence.py::test_config4_full_size_sample.)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from iceberg_tracking_code_b200 import build, cv, synthetic as syn

build.build()
H, W, N = 4000, 6000, 200000
lp = dict(winSize=(31, 31), maxLevel=5, criteria=(3, 30, 0.01))
base = syn.base_texture(H, W, 7, device="cuda")
g0, g1 = syn.frame_gray(base, 0), syn.frame_gray(base, 1)
del base
pts = syn.grid_points(H, W, step=11, start=10, limit=N).reshape(-1, 2).cuda()      # np.mgrid[10:4000:11, 10:6000:11]: 197 835 points
N = pts.shape[0]
pa, pb = cv.FramePyramid(g0, lp["winSize"], lp["maxLevel"], True), cv.FramePyramid(g1, lp["winSize"], lp["maxLevel"], True)
p1 = torch.empty((N, 2), dtype=torch.float32, device="cuda")
fbd = torch.empty((N,), dtype=torch.float32, device="cuda")
it = torch.zeros((1,), dtype=torch.int64, device="cuda")
for _ in range(3):
    cv.lk_fb_into(pa, pb, pts, lp, p1, fbd, None, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    cv.lk_fb_into(pa, pb, pts, lp, p1, fbd, None, it)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"points": N, "config": "config4: 6000x4000 pair, 200k-class grid (step 11), win 31, maxLevel 5, (3,30,0.01), fwd+bwd+FB",
                  "levels": pa.maxLevel + 1, "ms": ms, "points_per_s": N / (ms * 1e-3),
                  "iterations_per_point_pair": it.item() / reps / N, "iterations_per_s": it.item() / reps / (ms * 1e-3),
                  "fb_valid_fraction": float((fbd < 1).float().mean())}))
