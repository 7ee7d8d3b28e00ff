"""Config-3-style sequence throughput on one GPU: consecutive-pair tracking with re-seeding every track_len frames
(track_sequence on device-resident synthetic 24 MP frames; GFTT, compaction and the D2H of every group's tracks included)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iceberg_tracking_code_b200 import synthetic as syn
from iceberg_tracking_code_b200.tracking import track_sequence, SequenceTracker

H, W, NF, T = 4000, 6000, int(os.environ.get("NF", 13)), 2
maxc = int(os.environ.get("MAXC", 20000))
base = syn.base_texture(H, W, 7, device="cuda")
frames = [syn.frame_rgb(base, t, seed=7) for t in range(NF)]
del base
gp = dict(maxCorners=maxc, qualityLevel=0.007, minDistance=10, blockSize=10)
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
trk = SequenceTracker(gp, lp)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = track_sequence(frames, None, T, 60, loader=None, tracker=trk, save=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"frames": NF, "groups": len(res), "tracks_per_group": int(res[0][2].shape[0]), "s": dt,
                      "frames_per_s": (NF - 1) / dt, "ms_per_frame": dt / (NF - 1) * 1e3}))
