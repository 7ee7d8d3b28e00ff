"""One fused LK launch (24 MP pair, 20 k Shi-Tomasi points, win 31, L4) for ncu: `ncu -k regex:lk_kernel -c 1 --set full ...`.
WIN / NPTS / STERR=1 (status + err buffers) select variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iceberg_tracking_code_b200 import cv, synthetic as syn

H, W = 4000, 6000
win, n = int(os.environ.get("WIN", 31)), int(os.environ.get("NPTS", 20000))
base = syn.base_texture(H, W, 7, device="cuda")
f0, f1 = syn.frame_gray(base, 0), syn.frame_gray(base, 1)
del base
lp = dict(winSize=(win, win), maxLevel=4, criteria=(3, 30, 0.01))
pts = cv.goodFeaturesToTrack(f0, maxCorners=n, qualityLevel=0.007, minDistance=10, blockSize=10).reshape(-1, 2).contiguous()
pa, pb = cv.FramePyramid(f0, (win, win), 4), cv.FramePyramid(f1, (win, win), 4)
if os.environ.get("STERR") == "1":
    for _ in range(3):
        r = cv.calcOpticalFlowPyrLK_FB(pa, pb, pts, **lp)
else:
    p1 = torch.empty_like(pts); fbd = torch.empty((pts.shape[0],), device="cuda")
    for _ in range(3):
        cv.lk_fb_into(pa, pb, pts, lp, p1, fbd, None, None)
torch.cuda.synchronize()
print("ok", pts.shape[0])
