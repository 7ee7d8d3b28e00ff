"""One 24 MP decode for ncu (launch list / details of the JPEG kernels)."""
import io
import sys

import torch
from PIL import Image

sys.path.insert(0, ".")
from iceberg_tracking_code_b200 import jpeg, synthetic as syn, build  # noqa: E402

build.build()
base = syn.base_texture(4000, 6000, 7, device="cuda", scene="texture")
rgb = syn.frame_rgb(base, 0, seed=7).cpu().numpy()
bio = io.BytesIO()
Image.fromarray(rgb).save(bio, "JPEG")
dec = jpeg.JpegDecoder()
for _ in range(3):
    dec.decode(bio.getvalue(), rgb=False, gray=True)
torch.cuda.synchronize()
print("rounds", dec.last_rounds)
