"""ctypes binding of include/livevideo.h (the C ABI).  Loads the in-tree liblivevideo.so and nothing else:
there is no CPU fallback -- a missing library or a missing CUDA device raises."""
import ctypes as C
import os

from . import _build

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# name -> (restype, argtypes); one row per symbol declared in include/livevideo.h
SYMBOLS = {
    "lv_strerror": (C.c_char_p, [C.c_int]),
    "lv_last_error": (C.c_char_p, []),
    "lv_version": (C.c_int, []),
    "lv_device_count": (C.c_int, []),
    "YV12_to_RGB24": (None, [_vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "RGB24_to_YV12": (None, [_vp, _vp, C.c_int, C.c_int]),
    "YVU9_to_YV12": (None, [_vp, _vp, C.c_int, C.c_int]),
    "YUY2_to_YV12": (None, [_vp, _vp, C.c_int, C.c_int]),
    "YV12_to_YVU9": (None, [_vp, _vp, C.c_int, C.c_int]),
    "YV12_to_YUY2": (None, [_vp, _vp, C.c_int, C.c_int]),
    "FrameDiff_MB": (None, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "lv_convert_host_multi": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "lv_diff_host_multi": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "lv_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lv_destroy": (None, [_vp]),
    "lv_set_params": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lv_set_ref_mode": (C.c_int, [_vp, C.c_int]),
    "lv_frame_bytes_i420": (C.c_size_t, [_vp]),
    "lv_frame_bytes_bgr": (C.c_size_t, [_vp]),
    "lv_frame_mask_units": (C.c_size_t, [_vp]),
    "lv_frame_mbs": (C.c_size_t, [_vp]),
    "lv_set_prev_luma": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "lv_get_prev_luma": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "lv_reset_prev_luma": (C.c_int, [_vp, _vp]),
    "lv_convert_diff_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_convert_diff_batch_ex": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "lv_convert_diff_box_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "lv_convert_diff_box_batch_ex": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp,
                                               C.c_int]),
    "lv_object_windows": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "lv_object_update": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "lv_convert_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "lv_diff_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_write_bks": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "lv_object_box": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "lv_object_box_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "lv_narrow_pictures": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "lv_convert_diff_pictures": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp]),
    "lv_upload_async": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lv_upload_wait": (C.c_int, [_vp, C.c_int, _vp]),
    "lv_upload_release": (C.c_int, [_vp, C.c_int, _vp]),
    "lv_convert_format": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp]),
    "lv_format_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "lv_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "lv_host_free": (C.c_int, [_vp]),
    "lv_host_register": (C.c_int, [_vp, C.c_size_t]),
    "lv_host_unregister": (C.c_int, [_vp]),
    "lv_host_is_mapped": (C.c_int, [_vp, C.c_size_t]),
    "lv_host_step_plan": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, _vp, C.c_int]),
    "lv_step_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lv_step_host_pictures": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp,
                                        _vp, _vp, _vp]),
    "lv_shard_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lv_shard_destroy": (None, [_vp]),
    "lv_shard_device_count": (C.c_int, [_vp]),
    "lv_shard_range": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "lv_shard_ctx": (_vp, [_vp, C.c_int]),
    "lv_shard_stream": (_vp, [_vp, C.c_int]),
    "lv_shard_step": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "lv_gather_summaries": (C.c_int, [_vp, _vp]),
    "lv_shard_gathered": (_vp, [_vp, C.c_int]),
    "lv_shard_sync": (C.c_int, [_vp]),
    "lv_exchange_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "lv_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "lv_exchange_destroy": (None, [_vp]),
    "lv_exchange_buffer": (_vp, [_vp]),
    "lv_exchange_handle": (C.c_int, [_vp, _vp]),
    "lv_exchange_connect": (C.c_int, [_vp, _vp]),
    "lv_exchange_connect_ptrs": (C.c_int, [_vp, C.POINTER(_vp)]),
    "lv_exchange_publish": (C.c_int, [_vp, C.c_uint64, _vp, C.c_int, _vp]),
    "lv_exchange_collect": (C.c_int, [_vp, C.c_uint64, _vp, _vp]),
    "lv_exchange_error": (C.c_int, [_vp]),
    "lv_exchange_gate": (C.c_int, [_vp, _vp, C.c_int]),
    "lv_exchange_gate_open": (C.c_int, [_vp]),
    "lv_exchange_trace": (C.POINTER(C.c_uint64), [_vp]),
    "lv_exchange_wait_summary": (C.c_int, [_vp, C.c_uint64, _vp]),
    "lv_exchange_wait_table": (C.c_int, [_vp, C.c_uint64, _vp, C.POINTER(_vp)]),
    "lv_exchange_step": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   C.POINTER(C.c_uint64)]),
    "lv_exchange_flush": (C.c_int, [_vp]),
    "lv_exchange_result": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "lv_exchange_last_error": (C.c_char_p, []),
    "lv_set_kernel_variant": (C.c_int, [C.c_int]),
    "lv_get_kernel_variant": (C.c_int, []),
    "lv_set_zero_copy": (C.c_int, [C.c_int]),
    "lv_get_zero_copy": (C.c_int, []),
    "lv_kernel_launches": (C.c_uint64, [_vp]),
    "lv_device_sm_count": (C.c_int, [_vp]),
    "lv_gen_synthetic": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                   _vp, _vp]),
}

_lib = None


class LiveVideoError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().lv_strerror(code).decode()
        detail = lib().lv_last_error().decode()
        super().__init__(f"{where}: {msg} ({code})" + (f": {detail}" if detail else ""))


def lib_path():
    return os.environ.get("LV_LIB_OVERRIDE", _build.LIB_PATH)   # development only: A/B builds of the same ABI


def lib():
    """The loaded C-ABI library.  Raises if liblivevideo.so has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m livevideo_b200._build` (needs nvcc). "
                               "livevideo_b200 has no CPU fallback.")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, where):
    if code != 0:
        raise LiveVideoError(code, where)
