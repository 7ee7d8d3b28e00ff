"""BASELINE config 3 / 5: a full-day sequence (default 1440 synthetic 24 MP frames, one per minute), consecutive-pair
tracking with re-seeding every `track_len` frames, sharded by contiguous time blocks over the ranks, final
NCCL gather of all tracks; with UTM=1 every track vertex is also projected to map coordinates (config 5, K4).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/day_sequence.py          (NF=1440 MAXC=20000 T=2 UTM=0)

Each rank synthesises the frames of its own block ON DEVICE before the timed region (180 frames + 1 halo = 13 GB of
RGB per rank at N=8); nothing is shipped from the host.  Timed: gray, pyramids, GFTT, LK fwd/bwd/FB, compaction, D2H of
every group's tracks, the UTM projection of this rank's tracks (UTM=1) and the gather (rank 0 copies the gathered arrays to host
memory unless GATHER_HOST=0).  Time = max over ranks (barrier + synchronize on both sides)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from iceberg_tracking_code_b200 import build, sharding as sh, synthetic as syn, tracking as trk

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
build.build()
H, W = int(os.environ.get("H", 4000)), int(os.environ.get("W", 6000))
NF, T, MAXC = int(os.environ.get("NF", 1440)), int(os.environ.get("T", 2)), int(os.environ.get("MAXC", 20000))
gp = dict(maxCorners=MAXC, qualityLevel=0.007, minDistance=10, blockSize=10)
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
total = sh.n_groups(NF, T)
g0, n = sh.shard_groups(total, rank, world)
first, last = sh.frame_range(g0, n, T)
base = syn.base_texture(H, W, 100, device="cuda")
# the scene drifts slowly and wraps every 12 frames so that a day-long sequence stays inside the texture margin
frames = {t: syn.frame_rgb(base, t % 12, seed=100 + t) for t in range(first, last + 1)}
del base
torch.cuda.synchronize()
imagelist = list(range(NF))
tracker = trk.SequenceTracker(gp, lp)
# warm-up: the whole job once, untimed (allocator pools, pinned staging buffers of the gather, NCCL channels) -- a production run
# processes day after day with these in place
if os.environ.get("WARM", "1") == "1":
    dev_w = {}
    res_w = trk.track_sequence(imagelist, None, T, 60, loader=lambda t: frames[t], tracker=tracker, save=False, check_time=False,
                               first_group=g0, n_groups=n, decode_workers=0, device_results=dev_w)
    sh.gather_results(res_w, T, to_host="rank0" if os.environ.get('GATHER_HOST', '1') == '1' else False, device_results=dev_w)
    del res_w, dev_w
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
dev_res = {}
res = trk.track_sequence(imagelist, None, T, 60, loader=lambda t: frames[t], tracker=tracker, save=False, check_time=False,
                         first_group=g0, n_groups=n, decode_workers=0, device_results=dev_res)
torch.cuda.synchronize()
t_track = time.perf_counter() - t0
utm_vertices = 0
if os.environ.get("UTM", "0") == "1":
    from iceberg_tracking_code_b200 import camera
    # the synthetic camera of SURVEY 8d config 5 (create_calibration_file.py:8-30 values, image size = the frame)
    cam = camera.Camera(camname="cam1", parameters=dict(image_width=W, image_height=H, sensor_width=22.3, easting=377280.39,
                        northing=6525846.97, elevation=261.3, antenna_height=0.0, theta=300.0, phi=5.0, psi=-1.0, sigma=18.0))
    for _seed, (t_d, _q) in dev_res.items():          # every track vertex of this rank's block -> UTM (fp64), on the device arrays
        en = cam.tracks_to_utm(t_d)
        utm_vertices += en.shape[0] * en.shape[1]
    torch.cuda.synchronize()
t_utm = time.perf_counter() - t0 - t_track
tg0 = time.perf_counter()
allres = sh.gather_results(res, T, to_host="rank0" if os.environ.get('GATHER_HOST', '1') == '1' else False, device_results=dev_res)
torch.cuda.synchronize()
t_gather = time.perf_counter() - tg0
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
tt = torch.tensor([dt, t_track], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    ntracks = sum(len(a[1]) for a in allres)
    pairs = total * T
    print(json.dumps({"config": "day sequence %dx%d, %d frames, track_len %d, top-%d corners, win 31, L4" % (W, H, NF, T, MAXC),
                      "world": world, "groups": len(allres), "tracks_gathered": ntracks, "frame_pairs": pairs,
                      "seconds": float(tt[0]), "track_seconds_max_rank": float(tt[1]),
                      "frames_per_s": pairs / float(tt[0]), "tracked_points_per_s": ntracks * T / float(tt[0]),
                      "utm_vertices_rank0": utm_vertices, "utm_seconds_rank0": t_utm, "gather_seconds_rank0": t_gather}))
if world > 1:
    dist.destroy_process_group()
