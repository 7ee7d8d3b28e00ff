"""JPEG decode of a 24 MP frame: csrc/jpeg.cu vs Pillow on this box (np.array(Image.open(f)), s1:310)."""
import io
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, ".")
from iceberg_tracking_code_b200 import jpeg, synthetic as syn, build  # noqa: E402

build.build()
H, W = 4000, 6000
for scene in ("texture", "iceberg"):
    base = syn.base_texture(H, W, 7, device="cuda", scene=scene)
    rgb = syn.frame_rgb(base, 0, seed=7).cpu().numpy()
    bio = io.BytesIO()
    Image.fromarray(rgb).save(bio, "JPEG")                 # Pillow defaults, like the reference's crop step
    data = bio.getvalue()
    t0 = time.perf_counter()
    for _ in range(3):
        ref = np.array(Image.open(io.BytesIO(data)))
    t_pil = (time.perf_counter() - t0) / 3
    dec = jpeg.JpegDecoder()
    out, gray = dec.decode(data, rgb=True, gray=True)
    ok = bool(torch.equal(out.cpu(), torch.from_numpy(ref)))
    for mode, kw in (("gray only", dict(rgb=False, gray=True)), ("rgb+gray", dict(rgb=True, gray=True))):
        for _ in range(3):
            dec.decode(data, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            dec.decode(data, **kw)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
        print("%s %s: file %.2f MB, GPU %.3f ms/frame (events) %.3f ms wall, rounds %d, Pillow %.1f ms, bit-exact %s" %
              (scene, mode, len(data) / 1e6, e0.elapsed_time(e1) / n, wall, dec.last_rounds, t_pil * 1e3, ok), flush=True)
