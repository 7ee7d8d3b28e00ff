"""The save-and-reopen round trip of the cropping pre-pass (camtools.py:80 -> s1:310) on a 24 MP frame: ibt_jpeg_recompress
(csrc/jpeg.cu) vs Pillow's crop + save + open on this box."""
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
box = (250, 400, 6000, 4000)                     # the crop of the reference's cam3 (left, upper, right, lower)
for scene in ("texture", "iceberg"):
    base = syn.base_texture(H, W, 7, device="cuda", scene=scene)
    rgb = syn.frame_rgb(base, 0, seed=7)
    img = Image.fromarray(rgb.cpu().numpy())
    t0 = time.perf_counter()
    bio = io.BytesIO()
    img.crop(box).save(bio, "JPEG")
    ref = np.array(Image.open(io.BytesIO(bio.getvalue())))
    t_pil = time.perf_counter() - t0
    dec = jpeg.JpegDecoder()
    view = rgb[box[1]:box[3], box[0]:box[2]]
    out, gray = dec.recompress(view, rgb=True, gray=True)
    ok = bool(torch.equal(out.cpu(), torch.from_numpy(ref)))
    for mode, kw in (("gray only", dict(rgb=False, gray=True)), ("rgb+gray", dict(rgb=True, gray=True))):
        for _ in range(3):
            dec.recompress(view, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            dec.recompress(view, **kw)
        e1.record()
        torch.cuda.synchronize()
        print("%s %s: crop %dx%d, GPU %.3f ms/frame, Pillow crop+save+open %.1f ms (1 core), bit-exact %s" %
              (scene, mode, box[2] - box[0], box[3] - box[1], e0.elapsed_time(e1) / n, t_pil * 1e3, ok), flush=True)
