"""GPU: the s3 consumer (gridding.utm_to_gridded_utm, csrc/grid.cu through the C-ABI) against the files the UNMODIFIED
reference worker wrote on the same inputs (tests/golden/s3_expected.npz, made by make_s3_golden.py) -- bit-exact, incl. the
floating-point cell sums (numpy's pairwise summation order is reproduced) -- and against the numpy oracle at larger sizes."""
import datetime as dt
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def grd(ibt):
    from iceberg_tracking_code_b200 import gridding
    return gridding


def _write_inputs(g, tmp):
    head, data = os.path.join(tmp, "output"), os.path.join(tmp, "data")
    os.makedirs(data)
    np.savez(os.path.join(data, "fjord_outline.npz"), x=g["fjord_x"], y=g["fjord_y"], id=np.zeros(len(g["fjord_x"])))
    for cf in g["in_files"]:
        cam, f = str(cf).split("/")
        ws = os.path.join(head, cam, "utm")
        os.makedirs(ws, exist_ok=True)
        a = {k: g["in_%s_%s_%s" % (cam, f[:-4], k)] for k in ("x", "y", "u", "v", "time")}
        np.savez(os.path.join(ws, f), speed=np.hypot(a["u"], a["v"]), **a)
    with open(os.path.join(data, "parameter_file.csv"), "w") as fh:
        fh.write("camera,start_day,end_day,start_time,tracking_duration\n"
                 "cam1,20190701,20190801,12:00,2.0\ncam2,20190701,20190801,12:30,1.5\ncam3,20190801,20190901,10:00,4.0\n")
    with open(os.path.join(data, "camera_time_drifts.csv"), "w") as fh:
        fh.write("cam,start_date,end_date,drift_start_sec,drift_pday_sec\ncam1,20190701,20190801,12.3,0.7\n")
    return head, data


@pytest.mark.parametrize("tag", ["w30", "day"])
def test_worker_matches_reference_files(grd, golden, tmp_path, tag):
    g = golden("s3_expected.npz")
    head, data = _write_inputs(g, str(tmp_path))
    tgt = os.path.join(str(tmp_path), "run")
    os.makedirs(tgt)
    time_window, grid_size, obs_thr = g[tag + "_args"]
    args = (["cam1", "cam2", "cam3"], head, "utm", tgt, data, os.path.join(data, "parameter_file.csv"),
            os.path.join(data, "camera_time_drifts.csv"), os.path.join(data, "fjord_outline.npz"), dt.datetime(2019, 7, 24),
            float(time_window), int(grid_size), 0.5, int(obs_thr), 0)
    grd.utm_to_gridded_utm(args)
    files = sorted(os.listdir(tgt))
    assert files == [str(f) for f in g[tag + "_files"]]
    for fi, f in enumerate(files):
        z = np.load(os.path.join(tgt, f))
        keys = [k[len("%s_%d_" % (tag, fi)):] for k in g if k.startswith("%s_%d_" % (tag, fi))]
        assert sorted(z.files) == sorted(keys)
        for k in keys:
            exp, got = g["%s_%d_%s" % (tag, fi, k)], z[k]
            assert got.shape == exp.shape and got.dtype == exp.dtype, (f, k, got.shape, exp.shape, got.dtype, exp.dtype)
            assert np.array_equal(got, exp), (f, k)          # bit-exact, the fp64 means included
        assert len(z["count"]) > 30 and z["count"].max() > 128   # pairwise blocks and the recursive split were exercised


def test_helpers_match_reference(grd, golden):
    import pandas as pd
    g = golden("s3_expected.npz")
    for s, e in zip(g["round_time_in"], g["round_time_out"]):
        assert grd.round_time(dt.datetime.fromisoformat(str(s)), 30 * 60).isoformat() == str(e)
    drifts = pd.DataFrame([dict(cam="cam1", start_date=20190701, end_date=20190801, drift_start_sec=12.3, drift_pday_sec=0.7)])
    assert grd.correct_time_drift("cam1", "20190724", drifts) == float(g["drift_cam1_20190724"])


@pytest.mark.parametrize("n,spacing", [(200000, 150), (50000, 97.5), (1000, 400), (0, 200)])
def test_grid_bin_vs_oracle(grd, oracle, n, spacing):
    """Random point clouds (plus points on cell seams), int and fractional spacing: counts and sums identical to numpy."""
    rng = np.random.default_rng(n + 1)
    ang = np.linspace(0, 2 * np.pi, 60)[:-1]
    fx = 5000.25 + 1300 * (1 + 0.3 * np.sin(5 * ang)) * np.cos(ang)
    fy = 9000.75 + 900 * (1 + 0.2 * np.cos(4 * ang)) * np.sin(ang)
    x = rng.uniform(fx.min() - 2 * spacing, fx.max() + 2 * spacing, n)
    y = rng.uniform(fy.min() - 2 * spacing, fy.max() + 2 * spacing, n)
    if n:
        k = n // 20
        x[:k] = fx.min() + spacing * rng.integers(-1, 20, k)
        y[k:2 * k] = fy.max() - spacing * rng.integers(-1, 15, k)
        x[2 * k:2 * k + 4] = [np.inf, -np.inf, 1e300, -1e300]      # far away: in no cell (a NaN coordinate is out of domain:
        y[2 * k + 4] = np.nan                                     # see csrc/grid.cu grid_assign_kernel)
    u = rng.normal(0, 1, n) * 10.0 ** rng.uniform(-3, 3, n)
    v = rng.normal(0, 1, n) * 10.0 ** rng.uniform(-3, 3, n)
    polys, cents, idx, rows, cols = oracle.grid_cells(fx, fy, spacing)
    grid = grd.create_grid_across_fjord({"x": fx, "y": fy}, spacing)
    assert grid[4] == rows and grid[5] == cols and grid[2] == idx
    assert np.array_equal(np.array(grid[0]), np.array(polys)) and np.array_equal(np.array(grid[1]), np.array(cents))
    cnt, su, sv = grd.grid_bin(x, y, u, v, [fx.min(), fy.max()], spacing, cols, rows)
    ref = oracle.grid_bin(polys, x, y, u, v)
    for (i, j), (c, a, b) in zip(idx, ref):
        q = i * rows + j
        assert cnt[q] == c and su[q] == a and sv[q] == b, (i, j, cnt[q], c, su[q], a)


def test_polygon_rule_gpu(ibt):
    """mask.cu (per pixel) and grid.cu (per point) against the literal transcription of matplotlib's point_in_path_impl on the
    degenerate cases of tests/test_polygon_rule.py: points on edges, vertices, edge extensions, collinear / repeated vertices,
    a concave vertex on the scan line.  (Parity against matplotlib itself is unpinned: it is not installable here.)"""
    import ctypes as C
    import torch
    from iceberg_tracking_code_b200 import _native as N, cv, gridding
    from test_polygon_rule import CASES, COLLINEAR, NOTCH, RECT, crossings_literal
    for poly, pt, want in CASES:
        assert bool(gridding.points_in_polygon(np.asarray(poly), np.asarray([pt]))[0]) == want, (poly, pt)
    for poly in (RECT, NOTCH, COLLINEAR, [(1.0, 1.0), (7.0, 2.0), (5.0, 6.0), (4.0, 3.0), (2.0, 6.0)]):
        xs, ys = np.meshgrid(np.arange(-2, 9, 0.5), np.arange(-2, 7, 0.5))
        pts = np.stack([xs.ravel(), ys.ravel()], 1)
        ref = np.array([crossings_literal(poly, x, y) for x, y in pts])
        assert np.array_equal(gridding.points_in_polygon(np.asarray(poly), pts).astype(bool), ref)
        # the mask kernel tests pixel (x, y) = integer coordinates against the polygon in pixel units
        h, w = 8, 10
        pd_ = torch.tensor(np.asarray(poly, np.float64), device="cuda")
        out = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        N.check(N.lib().ibt_polygon_mask(cv._ptr(pd_), len(poly), h, w, cv._ptr(out), w, 255, cv._stream()), "ibt_polygon_mask")
        refm = np.array([[255 if crossings_literal(poly, float(x), float(y)) else 0 for x in range(w)] for y in range(h)], np.uint8)
        assert np.array_equal(out.cpu().numpy(), refm), poly
