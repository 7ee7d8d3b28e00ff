"""The per-frame prepare (cvtColor + pyramid/Scharr of a 24 MP frame) a few times, for ncu:
    ncu --set full -k regex:"pyr_level|gray_c3" -s 12 -c 6 ... python tools/prof_prepare.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iceberg_tracking_code_b200 import synthetic as syn
from iceberg_tracking_code_b200.tracking import SequenceTracker

H, W = 4000, 6000
base = syn.base_texture(H, W, 7, device="cuda")
frames = [syn.frame_rgb(base, t, seed=7) for t in range(3)]
del base
trk = SequenceTracker(lk_params=dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01)))
pyr = trk.prepare(frames[0])
for i in range(4):
    trk.prepare(frames[(i + 1) % 3], reuse=pyr)
torch.cuda.synchronize()
print("ok")
