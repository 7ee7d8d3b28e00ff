"""The consumer after s2: s3_utm_to_gridded_utm.py `utm_to_gridded_utm(arguments)` with the same 14-tuple argument and the
same `<start>-<end>_<min>min_<grid>m.npz` / `..._full_day_<grid>m.npz` files.

The reference merges the hourly velocity files of all cameras for a time window (clock-drift corrected), lays a grid of
square cells over the fjord outline and, per cell, runs matplotlib.path.Path(poly).contains_points over ALL velocities
(s3:391-421) before averaging.  Here the membership test, the ordering and the sums are one C-ABI call (`ibt_grid_bin`,
csrc/grid.cu: 3x3 candidate cells per point, stable radix sort, numpy's pairwise summation per cell), and the
cell-centre-in-fjord test of imports/tracking_misc.py:52 is `ibt_points_in_polygon`.  Host logic (time windows, drift
table, file names, thresholds) follows the reference line by line; plotting (plot_switch) is out of scope and ignored.
"""
import ctypes as C
import datetime as dt
import glob
import math
import os.path as osp

import numpy as np
import torch

from . import _native as N
from . import cv
from .camera import read_paramfile


# ---- imports/tracking_misc.py helpers (host) ---------------------------------------------------------------------------
def round_time(time, round_to=60):
    """tracking_misc.py:205-219"""
    seconds = (time.replace(tzinfo=None) - time.min).seconds
    rounding = (seconds + round_to / 2) // round_to * round_to
    return time + dt.timedelta(0, rounding - seconds, -time.microsecond)


def datetime_to_epoch(stamp):
    """tracking_misc.py:237-239"""
    return int((stamp - dt.datetime(1970, 1, 1)).total_seconds())


def epoch_to_datetime(epoch):
    """tracking_misc.py:241-243"""
    return dt.timedelta(seconds=float(epoch)) + dt.datetime(1970, 1, 1)


def correct_time_drift(camnr, date, time_drift_file):
    """tracking_misc.py:332-349: seconds to add to the photo time stamps of `camnr` on `date` ('%Y%m%d')."""
    sel = time_drift_file[(time_drift_file['cam'] == camnr) & (time_drift_file['start_date'] < int(date)) &
                          (time_drift_file['end_date'] >= int(date))]
    date_dt = dt.datetime.strptime(date, '%Y%m%d')
    date_dt_start = dt.datetime.strptime(str(sel['start_date'].iloc[0]), '%Y%m%d')
    difference = date_dt - date_dt_start
    correction_seconds = sel['drift_start_sec'].iloc[0] + difference.days * sel['drift_pday_sec'].iloc[0]
    return round(correction_seconds, 1)


def return_velocities_by_time(workspace, start_time, end_time):
    """tracking_misc.py:245-293: the velocities of the hourly files of `workspace` with start <= time < end."""
    start_epoch, end_epoch = datetime_to_epoch(start_time), datetime_to_epoch(end_time)
    hour = start_time.replace(minute=0, second=0)
    last = end_time.replace(minute=0, second=0)
    sel = {k: [np.array([])] for k in ("x", "y", "u", "v", "speed", "time")}
    while hour <= last:                                    # pd.date_range(start_trunc, end_trunc, freq='H')
        files = glob.glob(osp.join(str(workspace), hour.strftime('%Y%m%d_%H00') + '*.npz'))
        if files:
            try:
                npz = np.load(files[0])
                time = npz['time']
                mask = (time >= start_epoch) & (time < end_epoch)
                part = {k: npz[k][mask] for k in sel}
                for k in sel:
                    sel[k].append(part[k])
            except Exception:                              # noqa: BLE001  (the reference passes on any error, :289-291)
                pass
        hour += dt.timedelta(hours=1)
    return [np.concatenate(sel[k]) for k in ("x", "y", "u", "v", "speed", "time")]


# ---- the grid -------------------------------------------------------------------------------------------------------
def points_in_polygon(poly_xy, pts_xy):
    """matplotlib.path.Path(poly_xy).contains_points(pts_xy) on the GPU -> bool array."""
    poly = np.ascontiguousarray(poly_xy, np.float64).reshape(-1, 2)
    pts = np.ascontiguousarray(pts_xy, np.float64).reshape(-1, 2)
    dev = cv._device()
    dp, dq = torch.from_numpy(poly).to(dev), torch.from_numpy(pts).to(dev)
    out = torch.empty((pts.shape[0],), dtype=torch.uint8, device=dev)
    N.check(N.lib().ibt_points_in_polygon(cv._ptr(dp), poly.shape[0], cv._ptr(dq), pts.shape[0], cv._ptr(out), cv._stream()),
            "ibt_points_in_polygon")
    return out.cpu().numpy().astype(bool)


def create_grid_across_fjord(fjord, spacing):
    """tracking_misc.py:25-58: [polygons, centerpoints, indices, topleft_px_center, rows, cols] of the cells whose centre
    lies inside the fjord outline, column by column (i outer, j inner), same floating-point expressions."""
    fx, fy = np.asarray(fjord['x']), np.asarray(fjord['y'])
    topleft = [min(fx), max(fy)]
    topleft_px_center = [min(fx) + 0.5 * spacing, max(fy) - 0.5 * spacing]
    cols = int(math.ceil((max(fx) - min(fx)) / spacing))
    rows = int(math.ceil((max(fy) - min(fy)) / spacing))
    ii, jj = np.meshgrid(np.arange(cols), np.arange(rows), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    x = topleft[0] + ii * spacing                          # origin of create_squares (tracking_misc.py:44)
    y = topleft[1] - jj * spacing
    xr, yb = x + spacing, y - spacing
    polys = np.stack([np.stack([x, y], -1), np.stack([xr, y], -1), np.stack([xr, yb], -1), np.stack([x, yb], -1)], 1)
    centres = np.stack([x + 0.5 * spacing, y - 0.5 * spacing], -1)
    inside = points_in_polygon(np.vstack((fx, fy)).T, centres) if len(centres) else np.zeros((0,), bool)
    polygons = [[tuple(v) for v in p] for p in polys[inside].tolist()]
    centerpoints = centres[inside].tolist()
    indices = [[int(a), int(b)] for a, b in zip(ii[inside], jj[inside])]
    return [polygons, centerpoints, indices, topleft_px_center, rows, cols]


def grid_bin(x, y, u, v, topleft, spacing, cols, rows):
    """Per cell of the full cols x rows grid: (count int32, sum_u f64, sum_v f64), flat index i*rows + j (ibt_grid_bin)."""
    dev = cv._device()
    arrs = [torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev) for a in (x, y, u, v)]
    n = arrs[0].numel()
    ncell = cols * rows
    need = N.lib().ibt_grid_bin_workspace_bytes(n, cols, rows)
    if need <= 0 and n > 0:
        raise cv.error("grid_bin: grid or point set too large")
    ws = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=dev)
    count = torch.empty((ncell,), dtype=torch.int32, device=dev)
    su = torch.empty((ncell,), dtype=torch.float64, device=dev)
    sv = torch.empty((ncell,), dtype=torch.float64, device=dev)
    N.check(N.lib().ibt_grid_bin(cv._ptr(arrs[0]), cv._ptr(arrs[1]), cv._ptr(arrs[2]), cv._ptr(arrs[3]), n, float(topleft[0]),
                                 float(topleft[1]), float(spacing), cols, rows, cv._ptr(ws), ws.numel(), cv._ptr(count),
                                 cv._ptr(su), cv._ptr(sv), cv._stream()), "ibt_grid_bin")
    return count.cpu().numpy(), su.cpu().numpy(), sv.cpu().numpy()


# ---- the worker -------------------------------------------------------------------------------------------------------
def utm_to_gridded_utm(arguments):
    """Same call as s3_utm_to_gridded_utm.py:222: arguments = (camnames, source_path_head, source_path_tail, target_path,
    source_path_photos, paramfile_path, clockdrift_path, fjord_outline_path, day, time_window, grid_size,
    speedthreshold_cbar, observation_threshold, plot_switch).  Writes the same .npz files; plot_switch is ignored."""
    (camnames, source_path_head, source_path_tail, target_path, source_path_photos, paramfile_path, clockdrift_path,
     fjord_outline_path, day, time_window, grid_size, speedthreshold_cbar, observation_threshold, plot_switch) = arguments
    day_str = day.strftime('%Y%m%d')
    print('working on ' + day_str)
    paramfile = read_paramfile(paramfile_path)
    startlist, endlist, camnames_filtered = [], [], []
    for camname in camnames:                                                   # s3:245-262
        parameters = paramfile.loc[(paramfile['camera'] == camname) & (paramfile['start_day'] <= int(day_str)) &
                                   (paramfile['end_day'] >= int(day_str))]
        if len(parameters.index) == 1:
            timeobj = dt.datetime.strptime(parameters['start_time'].iloc[0], '%H:%M').time()
            start = timeobj.hour + timeobj.minute / 60.0
            end = start + parameters['tracking_duration'].iloc[0]
            startlist.append(start)
            endlist.append(end)
            camnames_filtered.append(camname)
    if len(camnames_filtered) == 0:
        print(day_str + ' done...')
        return
    edges = list(np.arange(min(startlist), max(endlist) + 0.001, time_window))  # s3:268-269
    start_hours, end_hours = edges[0:-1], edges[1:]
    if time_window == 24.0:
        start_hours, end_hours = [min(startlist)], [max(endlist)]
    time_drift_file = read_paramfile(clockdrift_path)
    fjord = np.load(str(fjord_outline_path))
    grid = None                                                                # the same for every window: built once
    for start_hour, end_hour in zip(start_hours, end_hours):
        cam_with_tracks = []
        start_datetime = day + dt.timedelta(hours=float(start_hour))
        end_datetime = day + dt.timedelta(hours=float(end_hour))
        time_diff = int((end_datetime - start_datetime).total_seconds() / 60.0)
        parts = {k: [] for k in ("x", "y", "u", "v")}
        mintimelist, maxtimelist = [], []
        for camname in camnames_filtered:                                      # s3:303-359
            try:
                time_correction = correct_time_drift(camname, day_str, time_drift_file)
            except Exception:                                                  # noqa: BLE001
                print(camname + ': no time drift correction available')
                time_correction = 0
            start_corr = start_datetime - dt.timedelta(seconds=float(time_correction))
            end_corr = end_datetime - dt.timedelta(seconds=float(time_correction))
            workspace = osp.join(str(source_path_head), camname, str(source_path_tail))
            if len(glob.glob(osp.join(workspace, day_str + '*utm.npz'))) == 0:
                continue
            x_sel, y_sel, u_sel, v_sel, _speed_sel, time_sel = return_velocities_by_time(workspace, start_corr, end_corr)
            if len(u_sel) > 0:
                cam_with_tracks.append(camname)
                mintimelist.append(epoch_to_datetime(min(time_sel)) + dt.timedelta(seconds=float(time_correction)))
                maxtimelist.append(epoch_to_datetime(max(time_sel)) + dt.timedelta(seconds=float(time_correction)))
                for k, a in zip(("x", "y", "u", "v"), (x_sel, y_sel, u_sel, v_sel)):
                    parts[k].append(a)
        if len(cam_with_tracks) == 0:
            continue
        x_all, y_all, u_all, v_all = (np.concatenate(parts[k]) for k in ("x", "y", "u", "v"))
        if len(x_all) == 0:
            continue
        if grid is None:
            grid = create_grid_across_fjord(fjord, grid_size)
        polygons_coarse, centerpoints_coarse, indices, topleft, rows, cols = grid
        topleft_corner = [min(fjord['x']), max(fjord['y'])]
        count_all, su_all, sv_all = grid_bin(x_all, y_all, u_all, v_all, topleft_corner, grid_size, cols, rows)
        out = {k: [] for k in ("grid_id", "i", "j", "x", "y", "u", "v", "speed", "count", "measured", "not_measured")}
        for counter, (poly, centerpoint, index) in enumerate(zip(polygons_coarse, centerpoints_coarse, indices)):   # s3:391-421
            c = index[0] * rows + index[1]
            nr_observations = int(count_all[c])
            if nr_observations > observation_threshold:
                mean_u = su_all[c] / nr_observations
                mean_v = sv_all[c] / nr_observations
                out["i"].append(index[0]); out["j"].append(index[1])
                out["x"].append(centerpoint[0]); out["y"].append(centerpoint[1])
                out["u"].append(mean_u); out["v"].append(mean_v)
                out["speed"].append(np.hypot(mean_u, mean_v))
                out["count"].append(nr_observations)
                out["grid_id"].append(counter)
                out["measured"].append(poly)
            else:
                out["not_measured"].append(poly)
        min_time = round_time(min(mintimelist), 30 * 60)
        max_time = round_time(max(maxtimelist), 30 * 60)
        if time_window == 24.0:                                                # s3:427-439
            npz_name = osp.join(str(target_path), '{}-{}_full_day_{}m.npz'.format(min_time.strftime('%Y%m%d_%H%M'),
                                                                                  max_time.strftime('%H%M'), grid_size))
        else:
            npz_name = osp.join(str(target_path), '{}-{}_{}min_{}m.npz'.format(start_datetime.strftime('%Y%m%d_%H%M'),
                                                                               end_datetime.strftime('%H%M'), time_diff, grid_size))
        np.savez(npz_name, grid_size=grid_size, topleft=topleft, rows=rows, cols=cols, grid_id=out["grid_id"], i=out["i"],
                 j=out["j"], x=out["x"], y=out["y"], u=out["u"], v=out["v"], speed=out["speed"], count=out["count"],
                 measured=out["measured"], not_measured=out["not_measured"])
    print(day_str + ' done...')
