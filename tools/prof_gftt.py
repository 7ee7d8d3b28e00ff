"""Profiling helper: one 24 MP goodFeaturesToTrack call (config-2 parameters) after a warm-up call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iceberg_tracking_code_b200 import cv, synthetic as syn

H, W = 4000, 6000
base = syn.base_texture(H, W, 7, device="cuda")
g = syn.frame_gray(base, 0)
del base
maxc = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = cv.goodFeaturesToTrack(g, maxCorners=maxc, qualityLevel=0.007, minDistance=10, blockSize=10)
    torch.cuda.synchronize(); print("gftt ms", (time.perf_counter() - t0) * 1e3, None if p is None else tuple(p.shape))

# phase stamps of the select kernel (GfttCounters.tstamp: after 6+2+64+4096 uint32)
import numpy as np
ws = cv._gftt_workspace(g.device, H, W)[0]
off = (8 + 64 + 4096) * 4
ts = ws[off:off + 16 * 8].cpu().numpy().view(np.uint64).astype(np.int64)
names = {0: "start", 1: "hist+zero", 2: "bins", 3: "select", 4: "pass4", 5: "pass5", 6: "pass6", 7: "pass7", 8: "pos+cells", 9: "cull", 10: "write"}
prev = ts[0]
for i in range(1, 11):
    if ts[i] >= ts[0] and ts[i] > 0:
        print("  %-10s +%6.1f us (at %6.1f)" % (names[i], (ts[i] - prev) / 1e3, (ts[i] - ts[0]) / 1e3)); prev = ts[i]


