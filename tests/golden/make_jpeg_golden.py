"""Generates tests/golden/jpeg/: small JPEG files written by Pillow (the encoder the reference's cropping step uses,
imports/camtools.py:63-104 `img_crop.save(outpath)`) and what `np.array(Image.open(f))` -- the call at
s1_lucaskanade_tracking.py:310 -- returns for each of them in the BUILD container (Pillow 12.2, libjpeg-turbo).
The oracle (oracle/jpeg_oracle.c) and the CUDA decoder (csrc/jpeg.cu) are both checked against these arrays.

Run:  python tests/golden/make_jpeg_golden.py       (needs Pillow; not needed to RUN the tests)
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from PIL import Image  # noqa: E402

from iceberg_tracking_code_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(HERE, "jpeg")


def rgb_scene(h, w, seed, kind):
    base = syn.base_texture(h, w, seed, scene=kind)
    return syn.frame_rgb(base, 0, seed=seed).numpy()


def cases():
    rng = np.random.default_rng(5)
    tex = rgb_scene(203, 317, 3, "texture")
    ice = rgb_scene(256, 384, 4, "iceberg")
    noise = rng.integers(0, 256, (96, 131, 3), dtype=np.uint8)
    flat = np.full((120, 200, 3), (40, 90, 160), np.uint8)
    yy, xx = np.mgrid[0:90, 0:150]
    ramp = np.stack([(xx * 255 // 149), (yy * 255 // 89), ((xx + yy) * 255 // 238)], -1).astype(np.uint8)
    yield "tex_420_default", tex, {}                                   # img.save(path): quality 75, 4:2:0
    yield "tex_422_q90", tex, dict(quality=90, subsampling=1)
    yield "tex_444_q95_opt", tex, dict(quality=95, subsampling=0, optimize=True)
    yield "ice_420_default", ice, {}
    yield "ice_420_q30_opt", ice, dict(quality=30, optimize=True)
    yield "noise_420_q100", noise, dict(quality=100)
    yield "noise_444_q50", noise, dict(quality=50, subsampling=0)
    yield "flat_420", flat, {}
    yield "ramp_422_q85", ramp, dict(quality=85, subsampling=1)
    yield "gray_q80", tex[..., 1], dict(quality=80)
    yield "tiny_1x1", tex[:1, :1], {}
    yield "tiny_8x8_444", tex[:8, :8], dict(subsampling=0)
    yield "thin_17x1", tex[:17, :1], {}
    yield "thin_3x40_422", tex[:3, :40], dict(subsampling=1)
    yield "odd_37x53", tex[:37, :53], {}
    yield "dri_420", ice[:64, :96], dict(restart_marker_blocks=3)      # restart markers: IBT_E_UNSUPPORTED on the GPU


def main():
    os.makedirs(OUT, exist_ok=True)
    expected = {}
    for name, img, kw in cases():
        bio = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(img)).save(bio, "JPEG", **kw)
        data = bio.getvalue()
        with open(os.path.join(OUT, name + ".jpg"), "wb") as f:
            f.write(data)
        expected[name] = np.array(Image.open(io.BytesIO(data)))
        print(name, expected[name].shape, len(data), "bytes")
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)


if __name__ == "__main__":
    main()
