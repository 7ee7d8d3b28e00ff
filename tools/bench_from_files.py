"""bench.py's from-files loop (JPEG bytes in host memory -> decode on the GPU -> pyramid -> fused LK, results read back) alone,
for several numbers of decoders in flight:  python tools/bench_from_files.py 2 3 4"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ.setdefault("IBT_BENCH_BREAKDOWN", "1")
import bench  # noqa: E402
from iceberg_tracking_code_b200 import build  # noqa: E402

build.build()
from iceberg_tracking_code_b200 import cv  # noqa: E402
from iceberg_tracking_code_b200.tracking import SequenceTracker  # noqa: E402

dev = torch.device("cuda", 0)
frames = bench.make_frames(dev)
pts = []
for f in frames:
    p = cv.goodFeaturesToTrack(cv.cvtColor(f, cv.COLOR_BGR2GRAY), **bench.GFTT)
    pts.append(p.reshape(bench.NPTS, 2).contiguous())
trk = SequenceTracker(bench.GFTT, bench.LK)
host_frames = [f.cpu().pin_memory() for f in frames]
for nd in [int(a) for a in sys.argv[1:]] or [2, 3]:
    pipe = bench.PairPipeline(trk, dev, frames[0])
    r = bench.run_from_files(trk, pipe, pts, host_frames, 120, cv, dev, ndec=nd)
    r.pop("api", None)
    print(json.dumps(r), flush=True)
    del pipe
    torch.cuda.empty_cache()
