"""Host-side mirror of the parts of imports/camtools.py:Camera that sit on either side of the tracking hot path:
crop box (camtools.py:214-258), water-mask polygon -> H x W mask (camtools.py:184-211, used at s1:285-294), and the
photo -> UTM projection of track vertices (camtools.py:286-332, 414-421, used at s2_cam_to_utm.py:243-254).

File parsing (Excel parameter sheet, mask shapefile, tide pickle) is out of the kernel scope; it is kept minimal
here: the parameter file may be .xlsx (pandas + openpyxl, as in the reference) or .csv with the same column names,
the mask polygon a .shp (needs pyshp, as in the reference) or a .npy / .csv list of (x, y) vertices.
"""
import ctypes as C
import os.path as osp

import numpy as np
import torch

from . import _native as N
from . import cv
from ._crop import crop_image_standalone

PARAM_COLUMNS = ["camera", "start_day", "end_day", "image_width", "image_height", "sensor_width", "easting", "northing",
                 "elevation", "antenna_height", "theta", "phi", "psi", "sigma", "crop_left", "crop_right", "crop_top",
                 "crop_bottom", "tracking_interval", "mask"]


def read_paramfile(paramfile_path):
    import pandas as pd
    p = str(paramfile_path)
    if p.endswith(".csv"):
        return pd.read_csv(p)
    return pd.read_excel(p)


def _load_polygon(path):
    path = str(path)
    if path.endswith(".npy"):
        pts = np.load(path)
    elif path.endswith(".csv"):
        pts = np.loadtxt(path, delimiter=",")
    else:
        import shapefile                                   # pyshp, like camtools.py:32
        shapes = shapefile.Reader(path).shapes()
        pts = [(int(pt[0]), int(pt[1] * -1)) for pt in shapes[0].points]       # camtools.py:58-59 (tuples=1)
    return [(float(a), float(b)) for a, b in np.asarray(pts).reshape(-1, 2)]


class Camera(object):
    """Same constructor arguments and the same `cam` / `pic` dictionaries as camtools.Camera (camtools.py:111-147)."""

    def __init__(self, camname=None, date=None, paramfile_path=None, mask=0, tide_corr=0, tide_file='', datetime='',
                 parameters=None, maskpoly=None, tide_elevation=None):
        if parameters is None:
            paramfile = read_paramfile(paramfile_path)
            rows = paramfile.loc[(paramfile['camera'] == camname) & (paramfile['start_day'] <= int(date)) &
                                 (paramfile['end_day'] >= int(date))]
            if len(rows) == 0:
                raise ValueError('No calibration parameters found for this day')
            parameters = {k: rows[k].iloc[0] for k in rows.columns}
        g = parameters
        self.camname = camname
        self.cam, self.pic = {}, {}
        self.pic['width'] = g['image_width']
        self.pic['height'] = g['image_height']
        self.cam['chipsize'] = g['sensor_width']
        self.cam['E'] = g['easting']
        self.cam['N'] = g['northing']
        self.cam['H'] = g['elevation'] - g['antenna_height']
        self.cam['theta'] = np.radians(g['theta'])
        self.cam['phi'] = np.radians(g['phi'])
        self.cam['psi'] = np.radians(g['psi'])
        self.cam['sigma'] = (self.pic['width'] / self.cam['chipsize']) * g['sigma']
        for k_src, k_dst in (('crop_left', 'cropleft'), ('crop_right', 'cropright'), ('crop_top', 'croptop'),
                             ('crop_bottom', 'cropbottom')):
            self.pic[k_dst] = g.get(k_src, 0)
        self.tracking_interval = g.get('tracking_interval')
        if maskpoly is not None:
            self.maskpoly = [(float(a), float(b)) for a, b in maskpoly]
        elif mask == 1:
            self.maskpoly = _load_polygon(osp.join(osp.dirname(str(paramfile_path)), camname, g['mask']))   # camtools.py:154
        if tide_elevation is not None:                     # camtools.py:158-182 (tide lookup itself is file I/O)
            self.cam['H'] = self.cam['H'] - float(tide_elevation)

    # -- crop (host I/O, as in the reference) -----------------------------------------------------------
    def crop_box(self):
        return (self.pic['cropleft'], self.pic['croptop'], self.pic['width'] - self.pic['cropright'],
                self.pic['height'] - self.pic['cropbottom'])

    def crop_image_parallel(self, imagelist, targetworkspace, n_cpus=10):
        """camtools.py:237-258: crop every source frame and re-save it with Pillow's defaults, in a pool of n_cpus
        processes like the reference (the lossy re-encode is host work and part of the pixels the reference tracks;
        `lucaskanade_tracking(crop="view")` is the way around it)."""
        box = tuple(int(v) for v in self.crop_box())
        args = [(str(img), osp.join(str(targetworkspace), osp.basename(str(img))), box) for img in imagelist]
        if n_cpus and n_cpus > 1 and len(args) > 1:
            import multiprocessing
            with multiprocessing.get_context("spawn").Pool(processes=min(int(n_cpus), len(args))) as pool:
                pool.map(crop_image_standalone, args)
        else:
            for a in args:
                crop_image_standalone(a)

    # -- mask ---------------------------------------------------------------------------------------------
    def mask_image(self, h, w, device=None):
        """mask[mask_meshgrid(x, y, origin='upper left') == 1] = 255 (s1:285-291) as one kernel: (h,w) u8 on the GPU."""
        dev = cv._device() if device is None else torch.device(device)
        poly = np.float64([(a - self.pic['cropleft'], b - self.pic['croptop']) for a, b in self.maskpoly])   # camtools.py:194
        if len(poly) >= 2 and np.array_equal(poly[0], poly[-1]):
            poly = poly[:-1]                               # explicit closing vertex: the kernel closes implicitly
        pd_ = torch.from_numpy(np.ascontiguousarray(poly)).to(dev)
        out = torch.empty((h, w), dtype=torch.uint8, device=dev)
        N.check(N.lib().ibt_polygon_mask(cv._ptr(pd_), int(poly.shape[0]), h, w, cv._ptr(out), w, 255, cv._stream()),
                "ibt_polygon_mask")
        return out

    # -- projection ---------------------------------------------------------------------------------------
    def utm_params(self):
        """The 12 doubles ibt_photo_to_utm takes (include/ibt.h)."""
        return np.float64([self.pic['cropleft'], self.pic['croptop'], self.pic['width'], self.pic['height'],
                           self.cam['sigma'], self.cam['H'], self.cam['theta'], self.cam['phi'], self.cam['psi'],
                           self.cam['E'], self.cam['N'], 0.0])

    def tracks_to_utm(self, tracks):
        """All vertices of a (M, T+1, 2) float32 track array -> (M, T+1, 2) float64 UTM (E, N): the triple loop of
        s2_cam_to_utm.py:243-254 as one launch."""
        as_np = isinstance(tracks, np.ndarray)
        t = cv._to_dev(tracks, np.float32, "tracks")
        out = cv.photo_to_utm(t.reshape(-1, 2), self.utm_params()).reshape(tuple(t.shape))
        return out.cpu().numpy() if as_np else out
