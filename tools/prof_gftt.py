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
