"""Generates tests/golden/jpeg/recompress.npz: RGB arrays and what the reference's cropping pre-pass + frame read make of them,
i.e. `img_crop.save(outpath)` (imports/camtools.py:80,102,232; Pillow defaults unless stated) followed by
`np.array(Image.open(image))` (s1_lucaskanade_tracking.py:310), produced by the REAL Pillow of the build container.
oracle.jpeg_recompress (oracle/jpeg_oracle.c) and the CUDA kernels (ibt_jpeg_recompress, csrc/jpeg.cu) are checked against
these arrays.

Run:  python tests/golden/make_recompress_golden.py       (needs Pillow; not needed to RUN the tests)
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from PIL import Image  # noqa: E402

from make_jpeg_golden import rgb_scene  # noqa: E402


def cases():
    rng = np.random.default_rng(17)
    tex = rgb_scene(203, 317, 3, "texture")
    ice = rgb_scene(256, 384, 4, "iceberg")
    noise = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    sat = (rng.integers(0, 2, (40, 56, 3)) * 255).astype(np.uint8)             # saturated corners of the RGB cube
    yield "tex_default", tex[10:131, 20:177], {}                               # 121 x 157: odd in both directions
    yield "ice_default", ice[:96, :160], {}                                    # whole MCUs
    yield "ice_even_partial", ice[7:57, 3:101], {}                             # 50 x 98: even, not a multiple of 16
    yield "noise_default", noise, {}
    yield "noise_q100", noise, dict(quality=100)
    yield "noise_q5", noise, dict(quality=5)
    yield "sat_default", sat, {}
    yield "tex_422_q90", tex[:33, :50], dict(quality=90, subsampling=1)
    yield "tex_444_q60", tex[:47, :29], dict(quality=60, subsampling=0)
    yield "tiny_1x1", tex[:1, :1], {}
    yield "thin_2x17", tex[:2, :17], {}
    yield "thin_19x1", tex[:19, :1], {}


def main():
    out = {}
    for name, img, kw in cases():
        img = np.ascontiguousarray(img)
        bio = io.BytesIO()
        Image.fromarray(img).save(bio, "JPEG", **kw)
        back = np.array(Image.open(io.BytesIO(bio.getvalue())))
        out["in_" + name] = img
        out["out_" + name] = back
        out["kw_" + name] = np.array([kw.get("quality", 75), kw.get("subsampling", 2)], np.int32)
        print(name, img.shape, "max |diff|", int(np.abs(img.astype(int) - back).max()))
    np.savez_compressed(os.path.join(HERE, "jpeg", "recompress.npz"), **out)


if __name__ == "__main__":
    main()
