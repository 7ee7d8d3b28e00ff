"""The C-ABI library builds for sm_100a, loads without a GPU and exports exactly what include/ibt.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from iceberg_tracking_code_b200 import build
    return build.build()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ibt.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ibt_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(so_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", so_path], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = header_symbols()
    assert len(declared) >= 15
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/ibt.h but not exported: %s" % missing
    extra = sorted(s for s in exported if s.startswith("ibt_") and s not in declared)
    assert not extra, "exported but not declared in include/ibt.h: %s" % extra


def test_binding_covers_header(so_path):
    from iceberg_tracking_code_b200 import _native
    assert sorted(_native.SIGNATURES) == header_symbols()
    lib = _native.lib()                    # resolves every symbol
    assert lib.ibt_version() >= 100
    assert lib.ibt_error_string(0) == b"ok"
    assert b"invalid" in lib.ibt_error_string(-1)


def test_host_only_entry_points(so_path):
    """ibt_pyramid_levels is pure host code: the level rule of SURVEY A.3 (strict >, both dims)."""
    from iceberg_tracking_code_b200 import _native
    lib = _native.lib()
    sizes = (C.c_int * 16)()
    assert lib.ibt_pyramid_levels(4000, 6000, 31, 31, 4, sizes) == 4
    assert list(sizes)[:10] == [4000, 6000, 2000, 3000, 1000, 1500, 500, 750, 250, 375]
    assert lib.ibt_pyramid_levels(140, 140, 35, 35, 4, sizes) == 1
    assert lib.ibt_pyramid_levels(142, 142, 35, 35, 4, sizes) == 2
    assert lib.ibt_pyramid_levels(100, 120, 35, 35, 4, sizes) == 1
    assert lib.ibt_pyramid_levels(10, 10, 35, 35, 9, sizes) == _native.IBT_E_INVALID


def test_invalid_arguments_are_rejected_before_any_cuda_call(so_path):
    """Argument validation comes first in every entry point: these calls return IBT_E_INVALID (or the documented no-op) on a
    box without a GPU, i.e. without having touched the CUDA runtime -- no compute call is made here."""
    from iceberg_tracking_code_b200 import _native as N
    lib = N.lib()
    null = C.c_void_p(0)
    pyr = N.ibt_pyramid_t()                                   # zeroed: nlevels == 0
    PP = C.POINTER(N.ibt_pyramid_t)
    one = (PP * 1)(C.pointer(pyr))
    # version of the ABI with cv2's useHarrisDetector / k and the multi-channel LK entry point
    assert lib.ibt_version() >= 101
    # goodFeaturesToTrack: null image, image smaller than the 3x3 Sobel support, NaN k
    cnt = C.c_int(7)
    assert lib.ibt_gftt(null, 64, null, 0, 64, 64, 0, 0.01, 1.0, 3, 0, 0.04, null, 0, null, 0, C.byref(cnt), null) == N.IBT_E_INVALID
    assert cnt.value == 0                                     # the host count is cleared even on failure
    assert lib.ibt_gftt_async(null, 64, null, 0, 64, 64, 0, 0.01, 1.0, 3, 0, 0.04, null, 0, null, 0, null, null) == N.IBT_E_INVALID
    assert lib.ibt_gftt_workspace_bytes(0, 10) == 0 and lib.ibt_gftt_workspace_bytes(4000, 6000) > 0
    assert lib.ibt_corner_harris_f32(null, 8, 8, 8, 3, 0.04, null, 32, null) == N.IBT_E_INVALID
    assert lib.ibt_min_eigen_f32(null, 8, 8, 8, 3, null, 32, null) == N.IBT_E_INVALID
    # LK: window outside [3, IBT_MAX_WIN], empty pyramids, channel count outside 1..4; N == 0 is a no-op
    assert lib.ibt_lk(C.byref(pyr), C.byref(pyr), null, null, 5, 2, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_E_INVALID
    assert lib.ibt_lk(C.byref(pyr), C.byref(pyr), null, null, 5, 21, N.IBT_MAX_WIN + 1, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_E_INVALID
    assert lib.ibt_lk(C.byref(pyr), C.byref(pyr), null, null, 0, 21, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_OK
    assert lib.ibt_lk(C.byref(pyr), C.byref(pyr), null, null, 5, 21, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_E_INVALID
    for cn in (0, 5):
        assert lib.ibt_lk_multichannel(one, one, cn, null, null, 5, 21, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_E_INVALID
    assert lib.ibt_lk_multichannel(one, one, 1, null, null, 0, 21, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_OK
    assert lib.ibt_lk_multichannel(None, None, 3, null, null, 5, 21, 21, 30, 0.01, 1e-4, 0, null, null, null, null) == N.IBT_E_INVALID
    # pyramid build: no levels; gray: null pointers
    assert lib.ibt_pyramid_build(C.byref(pyr), 1, null) == N.IBT_E_INVALID
    assert lib.ibt_pyr_level_u8(null, 8, 8, 8, null, 0, null, 0, null) == N.IBT_E_INVALID
    assert lib.ibt_lk_set_max_ctas_per_sm(-1) == N.IBT_E_INVALID and lib.ibt_lk_set_max_ctas_per_sm(0) == N.IBT_OK


def test_struct_layout_matches_header():
    from iceberg_tracking_code_b200 import _native
    # int32 nlevels + 2*8 int32 + pad to 8 + 4 * 8 * 8 bytes
    assert C.sizeof(_native.ibt_pyramid_t) == 4 + 64 + 4 + 4 * 64
    # struct ibt_jpeg_info: 20 int32, 2 int64, quant 4x64 u16, dc 2 x (4x16) u8, ac (4x16) + (4x256) u8
    assert C.sizeof(_native.ibt_jpeg_info_t) == 20 * 4 + 2 * 8 + 4 * 64 * 2 + 2 * 64 + 64 + 1024
    assert _native.ibt_jpeg_info_t.scan_offset.offset == 80 and _native.ibt_jpeg_info_t.quant.offset == 96
    assert _native.ibt_pyramid_t.img.offset == 72


def test_sass_is_sm100(so_path):
    out = subprocess.run(["cuobjdump", "-lelf", so_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from iceberg_tracking_code_b200 import cv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cv.cvtColor(np.zeros((4, 4, 3), np.uint8), cv.COLOR_BGR2GRAY)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "iceberg_tracking_code_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "ibt_oracle" not in src, f
                assert "import cv2" not in src, f
    # tools/ are measurement scripts around the product path: they must not execute the oracle either
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
