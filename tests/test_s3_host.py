"""CPU: host logic of the s3 worker (iceberg_tracking_code_b200/gridding.py) against values the reference's own helper functions
produced (tests/golden/s3_expected.npz, imports/tracking_misc.py:205-349 evaluated by make_s3_golden.py)."""
import datetime as dt
import os

import numpy as np


def test_round_time_and_drift(golden):
    import pandas as pd
    from iceberg_tracking_code_b200 import gridding as grd
    g = golden("s3_expected.npz")
    for s, e in zip(g["round_time_in"], g["round_time_out"]):
        assert grd.round_time(dt.datetime.fromisoformat(str(s)), 30 * 60).isoformat() == str(e)
    drifts = pd.DataFrame([dict(cam="cam1", start_date=20190701, end_date=20190801, drift_start_sec=12.3, drift_pday_sec=0.7)])
    assert grd.correct_time_drift("cam1", "20190724", drifts) == float(g["drift_cam1_20190724"])
    assert grd.datetime_to_epoch(dt.datetime(2019, 7, 24, 12)) == 1563969600
    assert grd.epoch_to_datetime(1563969600) == dt.datetime(2019, 7, 24, 12)


def test_return_velocities_by_time(golden, tmp_path):
    """tracking_misc.py:245-293: hourly files, start <= time < end, hours without a file are skipped, float64 out."""
    from iceberg_tracking_code_b200 import gridding as grd
    g = golden("s3_expected.npz")
    ws = tmp_path / "cam1" / "utm"
    ws.mkdir(parents=True)
    parts = {}
    for cf in g["in_files"]:
        cam, f = str(cf).split("/")
        if cam != "cam1":
            continue
        a = {k: g["in_%s_%s_%s" % (cam, f[:-4], k)] for k in ("x", "y", "u", "v", "time")}
        np.savez(ws / f, speed=np.hypot(a["u"], a["v"]), **a)
        parts[f] = a
    start, end = dt.datetime(2019, 7, 24, 12, 29, 47, 600000), dt.datetime(2019, 7, 24, 15, 10)      # 15:00 has no file
    out = grd.return_velocities_by_time(str(ws), start, end)
    t0, t1 = grd.datetime_to_epoch(start), grd.datetime_to_epoch(end)
    exp = {k: [] for k in ("x", "y", "u", "v", "time")}
    for f in sorted(parts):
        m = (parts[f]["time"] >= t0) & (parts[f]["time"] < t1)
        for k in exp:
            exp[k].append(parts[f][k][m])
    for k, arr in zip(("x", "y", "u", "v"), out[:4]):
        assert arr.dtype == np.float64 and np.array_equal(arr, np.concatenate(exp[k]))
    assert out[5].dtype == np.float64 and np.array_equal(out[5], np.concatenate(exp["time"]).astype(np.float64))
    empty = grd.return_velocities_by_time(str(ws), dt.datetime(2019, 7, 25, 1), dt.datetime(2019, 7, 25, 2))
    assert all(len(a) == 0 and a.dtype == np.float64 for a in empty)
