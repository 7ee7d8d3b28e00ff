"""LK kernel time vs number of points (24 MP pair, win 31, L4): shows launch / tail overheads vs steady throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iceberg_tracking_code_b200 import cv, synthetic as syn

H, W = 4000, 6000
base = syn.base_texture(H, W, 7, device="cuda")
f0, f1 = syn.frame_gray(base, 0), syn.frame_gray(base, 1)
del base
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
pts_all = cv.goodFeaturesToTrack(f0, maxCorners=0, qualityLevel=0.007, minDistance=10, blockSize=10).reshape(-1, 2).contiguous()
pa, pb = cv.FramePyramid(f0, (31, 31), 4), cv.FramePyramid(f1, (31, 31), 4)
for n in (1000, 5000, 20000, 40000, 80000, pts_all.shape[0]):
    pts = pts_all[:n].contiguous()
    p1 = torch.empty_like(pts); fbd = torch.empty((n,), device="cuda")
    it = torch.zeros((1,), dtype=torch.int64, device="cuda")
    for _ in range(3): cv.lk_fb_into(pa, pb, pts, lp, p1, fbd, None, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): cv.lk_fb_into(pa, pb, pts, lp, p1, fbd, None, it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("n %6d  %.3f ms  %.1f M pts/s  %.0f M iter/s  (%.1f it/pt)" % (n, ms, n / ms / 1e3, it.item() / 10 / ms / 1e3, it.item() / 10 / n))
