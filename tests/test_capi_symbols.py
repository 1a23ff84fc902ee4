"""The C-ABI library loads and exports every symbol include/livevideo.h declares (no compute calls: no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "livevideo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(ln for ln in text.splitlines() if not ln.lstrip().startswith("#"))
    return sorted(set(re.findall(r"LV_API\s+[^;(]*?\b(\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    from livevideo_b200 import _build
    return _build.build()


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for name in ("YV12_to_RGB24", "FrameDiff_MB", "lv_create", "lv_destroy", "lv_convert_diff_batch",
                 "lv_convert_batch", "lv_diff_batch", "lv_step_host", "lv_set_prev_luma", "lv_get_prev_luma"):
        assert name in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/livevideo.h but not exported"


def test_binding_table_matches_header(built, lv):
    from livevideo_b200 import capi
    assert sorted(capi.SYMBOLS) == declared_symbols()
    L = lv.lib()
    assert L.lv_version() >= 100
    assert L.lv_strerror(0) == b"ok" and L.lv_strerror(-2) == b"unsupported frame geometry"


def test_no_cpu_fallback_without_device(lv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(lv.LiveVideoError) as ei:
        lv.Context(1920, 1080)
    assert ei.value.code == -3      # LV_E_NODEVICE: fails loudly, nothing is computed on the CPU


def test_argument_validation_needs_no_device(lv):
    import ctypes as C
    h = C.c_void_p()
    L = lv.lib()
    assert L.lv_create(0, 1920, 1080, 0, 20, 0, C.byref(h)) == -1      # max_streams < 1
    assert L.lv_create(0, 1920, 1080, 1, 256, 0, C.byref(h)) == -1     # tol out of range
    assert L.lv_create(0, 1921, 1080, 1, 20, 0, C.byref(h)) == -2      # width % 4
    assert L.lv_create(0, 1920, 1081, 1, 20, 0, C.byref(h)) == -2      # odd height
    assert L.lv_create(0, 1920, 1080, 1, 20, 0, None) == -1
    assert L.lv_set_params(None, 20, 0) == -1
    assert L.lv_frame_bytes_i420(None) == 0


def test_product_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "livevideo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "lvoracle" not in src and "lv_oracle" not in src and "liblv_oracle" not in src, f
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "liblivevideo.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_exchange_argument_checks_need_no_gpu(built, lv):
    """lv_exchange_*: sizes are pure arithmetic; creation without a device reports LV_E_NODEVICE (never a fallback)."""
    import ctypes as C
    L = lv.lib()
    assert L.lv_exchange_bytes(8, 8, 128) == 8 * (8 * 8 * 128) + 4 * 8 + 4 * 8 + 16     # u64 words + acks + gates + err, 16-byte multiple
    assert L.lv_exchange_bytes(0, 8, 128) == 0 and L.lv_exchange_bytes(8, 1, 128) == 0
    h = C.c_void_p()
    assert L.lv_exchange_create(0, 2, 2, 8, 16, C.byref(h)) == -1          # rank out of range: LV_E_ARG
    if L.lv_device_count() == 0:
        assert L.lv_exchange_create(0, 0, 2, 8, 16, C.byref(h)) == -3      # LV_E_NODEVICE
        assert not h.value
