"""Seeded synthetic time-lapse frames.

The reference ships no images (README.md:9 asks the user to bring a folder of JPEGs), so the
tests and bench.py use these scenes: a blurred-noise texture translated by a constant sub-pixel
velocity (multiples of 1/32 px), optional per-frame sensor noise, optional RGB, and an
"iceberg" scene with flat water, bright blobs and a textureless band that exercises
status==0 / min-eigenvalue rejects.  Written once in torch so the same code runs on the CPU
(tests; bit-reproducible from the seed) and on the GPU (bench.py builds 24 MP frames on
device, never shipping them from the host inside a timed region).
"""
import math

import torch

VX = 2.3125      # px / frame, multiples of 1/32 (exactly representable)
VY = -1.40625
MARGIN = 32


def _gauss_blur(img, sigma):
    r = max(1, int(math.ceil(4 * sigma)))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=img.device)
    k = torch.exp(-(x * x) / (2 * sigma * sigma))
    k = (k / k.sum())
    t = img[None, None]
    t = torch.nn.functional.pad(t, (r, r, r, r), mode="reflect")
    t = torch.nn.functional.conv2d(t, k.view(1, 1, 1, -1))
    t = torch.nn.functional.conv2d(t, k.view(1, 1, -1, 1))
    return t[0, 0]


def base_texture(h, w, seed, device="cpu", scene="texture"):
    """(h+2*MARGIN, w+2*MARGIN) float32 in [0,255]."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    H, W = h + 2 * MARGIN, w + 2 * MARGIN
    if scene == "texture":
        n = torch.randint(0, 256, (H, W), generator=g, dtype=torch.int32).to(torch.float32).to(device)
        b = _gauss_blur(n, 2.0)
        b = (b - b.min()) / (b.max() - b.min()) * 255.0
        return b
    if scene == "iceberg":
        img = torch.full((H, W), 30.0, device=device)
        nblob = max(8, (h * w) // 12000)
        cy = torch.randint(0, H, (nblob,), generator=g)
        cx = torch.randint(0, int(W * 0.6), (nblob,), generator=g)     # right 40 % stays flat water
        amp = torch.randint(120, 226, (nblob,), generator=g).to(torch.float32)
        imp = torch.zeros((H, W))
        imp[cy, cx] = amp
        imp = imp.to(device)
        blobs = _gauss_blur(imp, 3.0) * (2 * math.pi * 9.0)
        fine = torch.randint(0, 256, (H, W), generator=g, dtype=torch.int32).to(torch.float32).to(device)
        fine = _gauss_blur(fine, 1.0) - 127.5
        wgt = (blobs / 60.0).clamp(0, 1)
        return (img + blobs + fine * wgt).clamp(0, 255)
    raise ValueError(scene)


def frame_gray(base, t, vx=VX, vy=VY, noise_sigma=0.0, seed=0):
    """Frame t as (h,w) uint8: base resampled bilinearly at offset (MARGIN+t*vx, MARGIN+t*vy)."""
    H, W = base.shape
    h, w = H - 2 * MARGIN, W - 2 * MARGIN
    ox, oy = MARGIN + t * vx, MARGIN + t * vy
    ix, iy = int(math.floor(ox)), int(math.floor(oy))
    ax, ay = ox - ix, oy - iy
    assert 0 <= ix and ix + w + 1 <= W and 0 <= iy and iy + h + 1 <= H, "shift leaves the margin"
    p = (base[iy:iy + h, ix:ix + w] * ((1 - ax) * (1 - ay)) + base[iy:iy + h, ix + 1:ix + w + 1] * (ax * (1 - ay)) +
         base[iy + 1:iy + h + 1, ix:ix + w] * ((1 - ax) * ay) + base[iy + 1:iy + h + 1, ix + 1:ix + w + 1] * (ax * ay))
    if noise_sigma > 0:
        g = torch.Generator(device=base.device).manual_seed(int(seed) * 100003 + int(t) + 1)
        p = p + torch.randn(p.shape, generator=g, device=base.device) * noise_sigma
    return p.round().clamp(0, 255).to(torch.uint8)


def frame_rgb(base, t, vx=VX, vy=VY, noise_sigma=2.0, seed=0):
    """(h,w,3) uint8: three independently noised copies (exercises the gray kernel)."""
    chans = [frame_gray(base, t, vx, vy, noise_sigma, seed * 3 + c + 17) for c in range(3)]
    return torch.stack(chans, dim=-1).contiguous()


def grid_points(h, w, step=11, start=10, limit=None):
    """Dense-seeded stress points (config 4): np.mgrid[start:h:step, start:w:step] as (N,1,2) f32 (x,y)."""
    ys = torch.arange(start, h, step, dtype=torch.float32)
    xs = torch.arange(start, w, step, dtype=torch.float32)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    p = torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=-1)
    if limit is not None:
        p = p[:limit]
    return p.reshape(-1, 1, 2).contiguous()
