"""livevideo_b200 -- B200-native fused YUV420 -> BGR24 + frame-difference hot path of wonsang/LiveVideo.

Only the hot path lives here: `csrc/` (hand-written sm_100a kernels + the C ABI of include/livevideo.h)
and `api.py`, the host-side mirror of the reference interface.  Importing the package does not load
CUDA; the first call does, and fails loudly when liblivevideo.so or a device is missing.
"""
from .api import (DEFAULT_MB_THRESH, DEFAULT_TOL, ColorSpaceConversions, Context, FrameDiff_MB, FrameDiff_MB_multi,  # noqa: F401
                  gen_synthetic, get_kernel_variant, get_zero_copy, host_is_mapped, host_register, host_step_plan, host_unregister, pinned,
                  set_kernel_variant,
                  set_zero_copy)
from .capi import LiveVideoError, lib, lib_path  # noqa: F401

__version__ = "0.1.0"
