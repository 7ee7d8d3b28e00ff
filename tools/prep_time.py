import os, sys, torch
sys.path.insert(0, "/root/repo")
from iceberg_tracking_code_b200 import synthetic as syn
from iceberg_tracking_code_b200.tracking import SequenceTracker
H, W = 4000, 6000
base = syn.base_texture(H, W, 7, device="cuda")
frames = [syn.frame_rgb(base, t, seed=7) for t in range(3)]
trk = SequenceTracker(lk_params=dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01)))
pyr = trk.prepare(frames[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(30):
    flush.fill_(i & 255)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); trk.prepare(frames[(i + 1) % 3], reuse=pyr); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort(); print("prepare us: min %.1f median %.1f" % (ts[0], ts[len(ts) // 2]), os.environ.get("IBT_PYR_BIG_MIN"))
