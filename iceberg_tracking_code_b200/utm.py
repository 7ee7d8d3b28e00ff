"""The consumer right after the tracking hot path: s2_cam_to_utm.py `cam_to_utm(arguments)` with the same 10-tuple
argument and the same hourly `<YYYYMMDD>_<HH>00_<dt>s_utm.npz` files (x, y, u, v, speed, time), SURVEY.md 8f-2.

Per track file the reference runs a triple Python loop (track -> vertex -> projection, s2:243-254) and a per-track
filter cascade (s2:279-343); here the whole file is one kernel launch (`ibt_track_velocities`: fp64 projection, segment
velocities, the three plausibility criteria) followed by array slicing.  The hour roll-over state machine (s2:195-230,
360-363), the file-name parsing and the tide lookup (camtools.py:158-182) are host logic kept as in the reference,
except that the parameter sheet and the tide table are read once per call instead of once per file (Appendix C "N").
"""
import ctypes as C
import datetime as dt
import glob
import os.path as osp

import numpy as np
import torch

from . import _native as N
from . import cv
from .camera import Camera, read_paramfile


def datetime_to_epoch(stamp):
    """imports/tracking_misc.py:237-239"""
    return int((stamp - dt.datetime(1970, 1, 1)).total_seconds())


def track_velocities(tracks, cam_params, tracking_interval, min_speed, max_speed, max_speedfactor, max_angle,
                     speed_threshold):
    """tracks (M, T+1, 2) float32 (numpy or CUDA tensor) -> dict of CUDA tensors: EN (M,T+1,2) f64, uv (M,T,2) f64,
    speed (M,T) f64, keep (M,) bool -- the per-track body of s2:243-343 for one file in one launch."""
    t = cv._to_dev(tracks, np.float32, "tracks")
    if t.ndim != 3 or t.shape[2] != 2 or t.shape[1] < 2:
        raise cv.error("track_velocities: tracks must be (M, T+1, 2) float32 with T >= 1")
    M, T = t.shape[0], t.shape[1] - 1
    dev = t.device
    EN = torch.empty((M, T + 1, 2), dtype=torch.float64, device=dev)
    uv = torch.empty((M, T, 2), dtype=torch.float64, device=dev)
    sp = torch.empty((M, T), dtype=torch.float64, device=dev)
    keep = torch.empty((M,), dtype=torch.uint8, device=dev)
    camv = (C.c_double * 12)(*[float(v) for v in cam_params])
    N.check(N.lib().ibt_track_velocities(cv._ptr(t), M, T, camv, float(tracking_interval), float(min_speed), float(max_speed),
                                         float(max_speedfactor), float(max_angle), float(speed_threshold), cv._ptr(EN),
                                         cv._ptr(uv), cv._ptr(sp), cv._ptr(keep), cv._stream()), "ibt_track_velocities")
    return dict(EN=EN, uv=uv, speed=sp, keep=keep.bool())


class _HourLists:
    """One hour's worth of vectors (the six parallel lists of s2:180-193), kept as lists of arrays."""

    def __init__(self):
        self.parts = {k: [] for k in ("x", "y", "u", "v", "speed", "time")}

    def extend(self, **arrs):
        for k, a in arrs.items():
            self.parts[k].append(a)

    def arrays(self):
        out = {}
        for k, lst in self.parts.items():
            if lst:
                out[k] = np.concatenate(lst)
            else:
                out[k] = np.zeros((0,), np.float64)          # np.savez of an empty list
        return out


def cam_to_utm(arguments):
    """Same call as s2_cam_to_utm.py:163: arguments = (source_workspace, target_workspace, camname, max_speed, min_speed,
    max_speedfactor, max_angle, speed_threshold, paramfile_path, tide_file_path)."""
    (source_workspace, target_workspace, camname, max_speed, min_speed,
     max_speedfactor, max_angle, speed_threshold, paramfile_path, tide_file_path) = arguments
    source_workspace, target_workspace = str(source_workspace), str(target_workspace)
    datestring = osp.basename(source_workspace)
    npzs = sorted(glob.glob(source_workspace + '/*.npz'))
    if len(npzs) == 0:
        print('folder {}: no files'.format(datestring))
        return
    tracking_interval = int(osp.basename(npzs[0]).split('_')[-2].split('sec')[0])          # s2:177

    import pandas as pd
    paramfile = read_paramfile(paramfile_path)
    rows = paramfile.loc[(paramfile['camera'] == camname) & (paramfile['start_day'] <= int(datestring)) &
                         (paramfile['end_day'] >= int(datestring))]
    if len(rows) == 0:
        raise ValueError('No calibration parameters found for this day')
    parameters = {k: rows[k].iloc[0] for k in rows.columns}
    tides = pd.read_pickle(str(tide_file_path))
    tide_by_minute = {ts.to_pydatetime(): float(v) for ts, v in zip(pd.to_datetime(tides['date']), tides['depth_tide_ellipsoid'])}

    cur, nxt = _HourLists(), _HourLists()
    next_hour = None
    npz_time_dt = None
    for c, npz in enumerate(npzs):
        npz_time = osp.basename(npz).split('_')[0]                                          # s2:197
        npz_time_dt = dt.datetime.strptime(npz_time, '%Y%m%d-%H%M%S')
        current_hour = npz_time_dt.hour
        if c == 0:
            next_hour = (npz_time_dt + dt.timedelta(hours=1)).hour
        if current_hour == next_hour:                                                       # s2:207-230
            label = npz_time_dt - dt.timedelta(hours=1)
            np.savez(osp.join(target_workspace, '{}_{}00_{}s_utm.npz'.format(label.strftime('%Y%m%d'), label.strftime('%H'),
                                                                             tracking_interval)), **cur.arrays())
            cur, nxt = nxt, _HourLists()
            next_hour = (npz_time_dt + dt.timedelta(hours=1)).hour

        tracks = np.load(npz)['tracks']
        if tracks.ndim != 3 or tracks.shape[0] == 0:
            continue                                                                        # an empty group (s1 wrote [])
        # tide-corrected camera of this file (camtools.py:158-182): one value per minute
        minute = npz_time_dt.replace(second=0, microsecond=0)
        cam = Camera(camname=camname, parameters=parameters, tide_elevation=float(tide_by_minute[minute]))
        r = track_velocities(np.ascontiguousarray(tracks, np.float32), cam.utm_params(), tracking_interval, min_speed,
                             max_speed, max_speedfactor, max_angle, speed_threshold)
        keep = r["keep"].cpu().numpy()
        EN = r["EN"].cpu().numpy()[keep]
        uv = r["uv"].cpu().numpy()[keep]
        sp = r["speed"].cpu().numpy()[keep]
        T = uv.shape[1]
        times = [npz_time_dt + dt.timedelta(seconds=i * tracking_interval) for i in range(T)]          # s2:281
        in_cur = np.array([t.hour == current_hour for t in times])
        epochs = np.array([datetime_to_epoch(t) for t in times], np.int64)
        for lists, sel in ((cur, in_cur), (nxt, ~in_cur)):
            if sel.any() and len(EN):
                lists.extend(x=EN[:, :-1, 0][:, sel].ravel(), y=EN[:, :-1, 1][:, sel].ravel(),
                             u=uv[:, sel, 0].ravel(), v=uv[:, sel, 1].ravel(), speed=sp[:, sel].ravel(),
                             time=np.tile(epochs[sel], len(EN)))
    np.savez(osp.join(target_workspace, '{}_{}00_{}s_utm.npz'.format(npz_time_dt.strftime('%Y%m%d'), npz_time_dt.strftime('%H'),
                                                                     tracking_interval)), **cur.arrays())   # s2:361-363
    print('folder {} done: {} files'.format(datestring, len(npzs)))
