"""Generates tests/golden/s3_expected.npz: the UNMODIFIED reference worker s3_utm_to_gridded_utm.utm_to_gridded_utm
(/root/reference/s3_utm_to_gridded_utm.py:222-467) run in the BUILD container on synthetic hourly velocity files of two
cameras, once with 30-minute windows and once with the full-day window; inputs and every array it wrote are stored.

Stand-ins needed because the dependencies are not installed here (everything else is the reference's own code):
  * matplotlib.path.Path  -> PathStandIn below: contains_point(s) by matplotlib's published even-odd "crossings" rule
                             (src/_path.h point_in_path_impl, radius 0), numpy fp64.  matplotlib itself is MagicMock.
  * pd.read_excel         -> returns the parameter / clock-drift tables (openpyxl is absent)
  * pd.date_range         -> 'H' -> 'h' (the frequency alias the reference uses was removed in pandas 3)

Run:  python tests/golden/make_s3_golden.py      (needs /root/reference; not needed to RUN the tests)
"""
import datetime as dt
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class PathStandIn:
    def __init__(self, vertices, *a, **k):
        self.v = np.asarray(vertices, np.float64).reshape(-1, 2)

    def contains_points(self, points, *a, **k):
        p = np.asarray(points, np.float64).reshape(-1, 2)
        tx, ty = p[:, 0], p[:, 1]
        inside = np.zeros(len(p), bool)
        v = self.v
        n = len(v)
        for e in range(n):
            v0, v1 = v[e], v[(e + 1) % n]
            yflag0, yflag1 = v0[1] >= ty, v1[1] >= ty
            cross = ((v1[1] - ty) * (v0[0] - v1[0]) >= (v1[0] - tx) * (v0[1] - v1[1])) == yflag1
            inside ^= (yflag0 != yflag1) & cross
        return inside

    def contains_point(self, point, *a, **k):
        return bool(self.contains_points([point])[0])


def main():
    import pandas as pd
    mpl_path = MagicMock()
    mpl_path.Path = PathStandIn
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.axes_grid1", "shapefile"):
        sys.modules.setdefault(m, MagicMock())
    sys.modules["matplotlib.path"] = mpl_path
    sys.modules["matplotlib"].path = mpl_path
    sys.path.insert(0, "/root/reference")
    import s3_utm_to_gridded_utm as s3
    import imports.tracking_misc as trm
    assert trm.mplPath.Path is PathStandIn and s3.mplPath.Path is PathStandIn

    params = pd.DataFrame([
        dict(camera="cam1", start_day=20190701, end_day=20190801, start_time="12:00", tracking_duration=2.0),
        dict(camera="cam2", start_day=20190701, end_day=20190801, start_time="12:30", tracking_duration=1.5),
        dict(camera="cam3", start_day=20190801, end_day=20190901, start_time="10:00", tracking_duration=4.0),   # other month
    ])
    drifts = pd.DataFrame([
        dict(cam="cam1", start_date=20190701, end_date=20190801, drift_start_sec=12.3, drift_pday_sec=0.7),
        # cam2 has no entry: "no time drift correction available" -> 0
    ])
    real_read_excel, real_date_range = pd.read_excel, pd.date_range
    pd.read_excel = lambda path, *a, **k: drifts if "drift" in str(path) else params
    pd.date_range = lambda *a, **k: real_date_range(*a, **{**k, "freq": "h" if k.get("freq") == "H" else k.get("freq")})

    rng = np.random.default_rng(33)
    tmp = tempfile.mkdtemp()
    head = os.path.join(tmp, "output")
    data = os.path.join(tmp, "data")
    os.makedirs(data)
    # fjord outline: an irregular polygon about 3.1 km x 2.3 km in UTM
    ang = np.linspace(0, 2 * np.pi, 37)[:-1]
    rad = 1.0 + 0.25 * np.sin(3 * ang) + 0.1 * np.cos(7 * ang)
    fx = 377000.0 + 1400.0 * rad * np.cos(ang) + 13.37
    fy = 6526000.0 + 1050.0 * rad * np.sin(ang) - 7.77
    np.savez(os.path.join(data, "fjord_outline.npz"), x=fx, y=fy, id=np.zeros(len(fx)))
    out = {"fjord_x": fx, "fjord_y": fy}
    day = dt.datetime(2019, 7, 24)
    inputs = []
    for cam, hours, n_per in (("cam1", (12, 13, 14), 9000), ("cam2", (12, 13), 6000)):
        ws = os.path.join(head, cam, "utm")
        os.makedirs(ws)
        for h in hours:
            n = n_per + 37 * h
            x = rng.uniform(fx.min() - 150, fx.max() + 150, n)
            y = rng.uniform(fy.min() - 150, fy.max() + 150, n)
            # some observations exactly on cell edges / corners of the 200 m grid, and a dense cluster (> 128 per cell)
            k = n // 50
            x[:k] = fx.min() + 200.0 * rng.integers(0, 14, k)
            y[k:2 * k] = fy.max() - 200.0 * rng.integers(0, 10, k)
            x[2 * k:3 * k] = fx.min() + 200.0 * 7 + rng.uniform(0, 200, k)
            y[2 * k:3 * k] = fy.max() - 200.0 * 5 - rng.uniform(0, 200, k)
            u = rng.normal(0.05, 0.2, n) * 10.0 ** rng.uniform(-2, 1, n)
            v = rng.normal(-0.02, 0.2, n) * 10.0 ** rng.uniform(-2, 1, n)
            t0 = int((day + dt.timedelta(hours=h) - dt.datetime(1970, 1, 1)).total_seconds())
            time = t0 + 60 * rng.integers(0, 60, n)
            f = "%s_%02d00_60s_utm.npz" % (day.strftime("%Y%m%d"), h)
            np.savez(os.path.join(ws, f), x=x, y=y, u=u, v=v, speed=np.hypot(u, v), time=time.astype(np.int64))
            inputs.append((cam, f))
            for key, arr in (("x", x), ("y", y), ("u", u), ("v", v), ("time", time.astype(np.int64))):
                out["in_%s_%s_%s" % (cam, f[:-4], key)] = arr
    out["in_files"] = np.array(["%s/%s" % cf for cf in inputs])

    runs = (("w30", 30 / 60.0, 200, 10), ("day", 24.0, 350, 25))
    for tag, time_window, grid_size, obs_thr in runs:
        tgt = os.path.join(tmp, "run_" + tag)
        os.makedirs(tgt)
        args = (["cam1", "cam2", "cam3"], head, "utm", tgt, data, os.path.join(data, "parameter_file.xlsx"),
                os.path.join(data, "camera_time_drifts.xlsx"), os.path.join(data, "fjord_outline.npz"), day, time_window, grid_size,
                0.5, obs_thr, 0)
        s3.utm_to_gridded_utm(args)
        files = sorted(os.listdir(tgt))
        out[tag + "_files"] = np.array(files)
        out[tag + "_args"] = np.array([time_window, grid_size, obs_thr], np.float64)
        for fi, f in enumerate(files):
            z = np.load(os.path.join(tgt, f))
            for k in z.files:
                out["%s_%d_%s" % (tag, fi, k)] = z[k]
        print(tag, files, [int(len(np.load(os.path.join(tgt, f))["count"])) for f in files])
    # helper functions of imports/tracking_misc.py evaluated by the reference itself
    stamps = [dt.datetime(2019, 7, 24, 12, 14, 59), dt.datetime(2019, 7, 24, 12, 15, 0), dt.datetime(2019, 7, 24, 23, 50, 1, 5000)]
    out["round_time_in"] = np.array([s.isoformat() for s in stamps])
    out["round_time_out"] = np.array([trm.round_time(s, 30 * 60).isoformat() for s in stamps])
    out["drift_cam1_20190724"] = np.float64(trm.correct_time_drift("cam1", "20190724", drifts))
    pd.read_excel, pd.date_range = real_read_excel, real_date_range
    np.savez_compressed(os.path.join(HERE, "s3_expected.npz"), **out)
    print("s3_expected.npz written,", os.path.getsize(os.path.join(HERE, "s3_expected.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
