"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

* ``ColorSpaceConversions`` -- same class/method name and argument meaning as the reference's
  ``class ColorSpaceConversions`` (Convert.h:19-31): ``YV12_to_RGB24(src0, src1, src2, dst_ori, width, height)``
  fills ``dst_ori`` in place (bottom-up B,G,R); ``YV12_to_RGB24_multi`` converts several pictures per call the way
  CMainDlg::DisplayPic issues them (MainDlg.cpp:937-977).  The five converters the application never calls
  (Convert.h:25-30) are implemented too, bit-exact with the object code.
* ``FrameDiff_MB`` -- the difference primitive extracted from lib/JM/image.c:4791-4804 / 5143-5150.
* ``Context`` -- the batched device API (lv_create ... lv_destroy): fused conversion + difference over
  many streams with the per-stream previous-luma state of lib/JM/image.c:1553-1561.

PyTorch is used only as the owner of device memory and streams; every computation happens in
liblivevideo.so (hand-written sm_100a kernels).  There is no CPU path.
"""
import ctypes as C

import numpy as np

from . import capi

DEFAULT_TOL = 20        # BKD_STR_TOL, lib/JM/global.h:125
DEFAULT_MB_THRESH = 0   # "any changed pixel", lib/JM/image.c:5143-5150


def _host_ptr(a, dtype, nbytes, name, writable=False):
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{name}: expected a numpy array")
    if a.dtype != dtype or not a.flags.c_contiguous:
        raise ValueError(f"{name}: expected a C-contiguous {np.dtype(dtype).name} array")
    if a.nbytes < nbytes:
        raise ValueError(f"{name}: needs {nbytes} bytes, has {a.nbytes}")
    if writable and not a.flags.writeable:
        raise ValueError(f"{name}: not writable")
    return a.ctypes.data


class ColorSpaceConversions:
    """Drop-in mirror of the reference class (Convert.h:19-31)."""

    def __init__(self):
        capi.lib()   # fail now if the CUDA library is missing

    def YV12_to_RGB24(self, src0, src1, src2, dst_ori, width, height):
        px = width * height
        if width <= 0 or height <= 0:
            return
        if width % 4 or height % 2:
            raise ValueError("YV12_to_RGB24 needs width % 4 == 0 and height % 2 == 0 (the original overruns)")
        capi.lib().YV12_to_RGB24(_host_ptr(src0, np.uint8, px, "src0"), _host_ptr(src1, np.uint8, px // 4, "src1"),
                                 _host_ptr(src2, np.uint8, px // 4, "src2"),
                                 _host_ptr(dst_ori, np.uint8, 3 * px, "dst_ori", True), width, height)

    def YV12_to_RGB24_multi(self, src0, src1, src2, dst_ori, width, height):
        """lv_convert_host_multi: lists of per-picture arrays (the dialog's srcY[5], srcU[5], srcV[5], RGBbuf[5],
        MainDlg.cpp:946-974); uploads, kernels and downloads of the pictures overlap, one wait at the end."""
        n = len(src0)
        if not (len(src1) == len(src2) == len(dst_ori) == n):
            raise ValueError("YV12_to_RGB24_multi: the four lists must have the same length")
        px = width * height
        arr = lambda xs, nb, name, w=False: (C.c_void_p * n)(*[_host_ptr(a, np.uint8, nb, name, w) for a in xs])  # noqa: E731
        capi.check(capi.lib().lv_convert_host_multi(arr(src0, px, "src0"), arr(src1, px // 4, "src1"), arr(src2, px // 4, "src2"),
                                                    arr(dst_ori, 3 * px, "dst_ori", True), n, width, height),
                   "lv_convert_host_multi")

    def bind_multi(self, src0, src1, src2, dst_ori, width, height):
        """The same call with the pointer arrays built ONCE (the dialog's srcY[5] ... RGBbuf[5] are members that never
        change, MainDlg.h): returns a function that issues lv_convert_host_multi on them."""
        n = len(src0)
        px = width * height
        arrs = [(C.c_void_p * n)(*[_host_ptr(a, np.uint8, nb, name, wr) for a in xs])
                for xs, nb, name, wr in ((src0, px, "src0", False), (src1, px // 4, "src1", False),
                                         (src2, px // 4, "src2", False), (dst_ori, 3 * px, "dst_ori", True))]
        keep = (list(src0), list(src1), list(src2), list(dst_ori))      # the arrays must outlive the binding
        fn = capi.lib().lv_convert_host_multi

        def call(_keep=keep):
            capi.check(fn(arrs[0], arrs[1], arrs[2], arrs[3], n, width, height), "lv_convert_host_multi")
        return call

    # the five methods the application never calls (Convert.h:25-30): same (in, out, w, h) convention, in place
    def _fmt(self, name, kind, inp, out, w, h):
        if w <= 0 or h <= 0:
            return
        if w % 4 or h % 4:
            raise ValueError(f"{name} needs width % 4 == 0 and height % 4 == 0")
        L = capi.lib()
        getattr(L, name)(_host_ptr(inp, np.uint8, L.lv_format_bytes(kind, w, h, 0), "in"),
                         _host_ptr(out, np.uint8, L.lv_format_bytes(kind, w, h, 1), "out", True), w, h)

    def RGB24_to_YV12(self, inp, out, w, h):
        self._fmt("RGB24_to_YV12", 1, inp, out, w, h)

    def YVU9_to_YV12(self, inp, out, w, h):
        self._fmt("YVU9_to_YV12", 2, inp, out, w, h)

    def YUY2_to_YV12(self, inp, out, w, h):
        self._fmt("YUY2_to_YV12", 3, inp, out, w, h)

    def YV12_to_YVU9(self, inp, out, w, h):
        self._fmt("YV12_to_YVU9", 4, inp, out, w, h)

    def YV12_to_YUY2(self, inp, out, w, h):
        self._fmt("YV12_to_YUY2", 5, inp, out, w, h)


def FrameDiff_MB(cur_y, ref_y, width, height, tol=DEFAULT_TOL, mb_thresh=DEFAULT_MB_THRESH, mask=None,
                 mb_count=None, mb_map=None):
    """Host-pointer difference entry.  Output arrays are allocated when not given; returns (mask, mb_count, mb_map)."""
    px = width * height
    mbs = ((width + 15) // 16) * ((height + 15) // 16)
    if mask is None:
        mask = np.empty(px, dtype=np.uint8)
    if mb_count is None:
        mb_count = np.empty(mbs, dtype=np.uint16)
    if mb_map is None:
        mb_map = np.empty(mbs, dtype=np.uint8)
    if width <= 0 or height <= 0:
        return mask, mb_count, mb_map
    capi.lib().FrameDiff_MB(_host_ptr(cur_y, np.uint8, px, "cur_y"), _host_ptr(ref_y, np.uint8, px, "ref_y"), width,
                            height, int(tol), int(mb_thresh), _host_ptr(mask, np.uint8, px, "mask", True),
                            _host_ptr(mb_count, np.uint16, 2 * mbs, "mb_count", True),
                            _host_ptr(mb_map, np.uint8, mbs, "mb_map", True))
    return mask, mb_count, mb_map


def FrameDiff_MB_multi(cur_y, ref_y, width, height, tol=DEFAULT_TOL, mb_thresh=DEFAULT_MB_THRESH, mask=None, mb_count=None,
                       mb_map=None):
    """lv_diff_host_multi: FrameDiff_MB over lists of pictures in one pipelined call.  Output lists are allocated when
    not given; returns (mask, mb_count, mb_map) as lists of arrays."""
    n = len(cur_y)
    if len(ref_y) != n:
        raise ValueError("FrameDiff_MB_multi: cur_y and ref_y must have the same length")
    px = width * height
    mbs = ((width + 15) // 16) * ((height + 15) // 16)
    if mask is None:
        mask = [np.empty(px, dtype=np.uint8) for _ in range(n)]
    if mb_count is None:
        mb_count = [np.empty(mbs, dtype=np.uint16) for _ in range(n)]
    if mb_map is None:
        mb_map = [np.empty(mbs, dtype=np.uint8) for _ in range(n)]
    if n == 0 or width <= 0 or height <= 0:
        return mask, mb_count, mb_map

    def arr(xs, dt, nb, name, w=False):
        return (C.c_void_p * n)(*[_host_ptr(a, dt, nb, name, w) for a in xs])
    capi.check(capi.lib().lv_diff_host_multi(arr(cur_y, np.uint8, px, "cur_y"), arr(ref_y, np.uint8, px, "ref_y"), n, width, height,
                                             int(tol), int(mb_thresh), arr(mask, np.uint8, px, "mask", True),
                                             arr(mb_count, np.uint16, 2 * mbs, "mb_count", True),
                                             arr(mb_map, np.uint8, mbs, "mb_map", True)), "lv_diff_host_multi")
    return mask, mb_count, mb_map


def _dev_ptr(t, dtype, numel, name, device, align=1):
    import torch
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.device.index != device:
        raise ValueError(f"{name}: tensor is on cuda:{t.device.index}, context is on cuda:{device}")
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous {dtype} tensor")
    if t.numel() < numel:
        raise ValueError(f"{name}: needs {numel} elements, has {t.numel()}")
    if t.data_ptr() % align:
        raise ValueError(f"{name}: must be {align}-byte aligned (a slice at an odd offset?)")
    return t.data_ptr()


def _stream_handle(stream):
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


class Context:
    """lv_ctx: fused conversion + frame difference for `max_streams` camera streams of one geometry."""

    def __init__(self, width, height, max_streams=1, tol=DEFAULT_TOL, mb_thresh=DEFAULT_MB_THRESH, device=0):
        self._h = C.c_void_p()
        L = capi.lib()
        capi.check(L.lv_create(device, width, height, max_streams, tol, mb_thresh, C.byref(self._h)), "lv_create")
        self.device, self.width, self.height, self.max_streams = device, width, height, max_streams
        self.px = width * height
        self.frame_bytes = L.lv_frame_bytes_i420(self._h)
        self.bgr_bytes = L.lv_frame_bytes_bgr(self._h)
        self.mask_units = L.lv_frame_mask_units(self._h)
        self.mbs = L.lv_frame_mbs(self._h)
        self.mb_w, self.mb_h = (width + 15) // 16, (height + 15) // 16

    def close(self):
        if self._h:
            capi.lib().lv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, tol, mb_thresh):
        capi.check(capi.lib().lv_set_params(self._h, tol, mb_thresh), "lv_set_params")

    @property
    def kernel_launches(self):
        return int(capi.lib().lv_kernel_launches(self._h))

    @property
    def sm_count(self):
        return int(capi.lib().lv_device_sm_count(self._h))

    # ---- previous-luma state ---------------------------------------------------------------------
    def set_prev_luma(self, stream, y, cuda_stream=None):
        if isinstance(y, np.ndarray):
            capi.check(capi.lib().lv_set_prev_luma(self._h, stream, _host_ptr(y, np.uint8, self.px, "y"), 0,
                                                   _stream_handle(cuda_stream)), "lv_set_prev_luma")
        else:
            import torch
            capi.check(capi.lib().lv_set_prev_luma(self._h, stream, _dev_ptr(y, torch.uint8, self.px, "y", self.device),
                                                   1, _stream_handle(cuda_stream)), "lv_set_prev_luma")

    def set_ref_mode(self, static: bool):
        """static=True: every frame is differenced against the stream's stored plane (the static background of
        MainDlg.cpp:238-246, set with set_prev_luma), which is left untouched; False: against the previous frame."""
        capi.check(capi.lib().lv_set_ref_mode(self._h, 1 if static else 0), "lv_set_ref_mode")

    def get_prev_luma(self, stream, cuda_stream=None):
        out = np.empty(self.px, dtype=np.uint8)
        capi.check(capi.lib().lv_get_prev_luma(self._h, stream, out.ctypes.data, 0, _stream_handle(cuda_stream)),
                   "lv_get_prev_luma")
        return out

    def reset_prev_luma(self, cuda_stream=None):
        capi.check(capi.lib().lv_reset_prev_luma(self._h, _stream_handle(cuda_stream)), "lv_reset_prev_luma")

    # ---- device API --------------------------------------------------------------------------------
    def alloc_outputs(self, n, bgr=True, mask_bits=False, mask_bytes=False):
        """Device output tensors for n frames."""
        import torch
        dev = torch.device("cuda", self.device)
        out = {
            "bgr": torch.empty(n * self.bgr_bytes, dtype=torch.uint8, device=dev) if bgr else None,
            "mask_bits": torch.empty(n * self.mask_units, dtype=torch.int16, device=dev) if mask_bits else None,
            "mask_bytes": torch.empty(n * self.px, dtype=torch.uint8, device=dev) if mask_bytes else None,
            "mb_count": torch.empty(n * self.mbs, dtype=torch.int16, device=dev),
            "mb_map": torch.empty(n * self.mbs, dtype=torch.uint8, device=dev),
            "summary": torch.empty(n * 2, dtype=torch.int32, device=dev),
        }
        return out

    def convert_diff_batch(self, i420, n_streams, n_frames, out, first_stream=0, cuda_stream=None, summary_ptr=None):
        """One fused launch over n_streams*n_frames frames resident in device memory (asynchronous).
        summary_ptr: raw device address that receives the summary instead of out["summary"] (the peer exchange hands
        out the block of its own table, livevideo_b200/shard.py)."""
        import torch
        n = n_streams * n_frames
        d = self.device
        capi.check(capi.lib().lv_convert_diff_batch(
            self._h, _dev_ptr(i420, torch.uint8, n * self.frame_bytes, "i420", d, 16), first_stream, n_streams, n_frames,
            _dev_ptr(out.get("bgr"), torch.uint8, n * self.bgr_bytes, "bgr", d, 16),
            _dev_ptr(out.get("mask_bits"), torch.int16, n * self.mask_units, "mask_bits", d),
            _dev_ptr(out.get("mask_bytes"), torch.uint8, n * self.px, "mask_bytes", d, 16),
            _dev_ptr(out.get("mb_count"), torch.int16, n * self.mbs, "mb_count", d),
            _dev_ptr(out.get("mb_map"), torch.uint8, n * self.mbs, "mb_map", d),
            summary_ptr if summary_ptr is not None else _dev_ptr(out.get("summary"), torch.int32, n * 2, "summary", d),
            _stream_handle(cuda_stream)), "lv_convert_diff_batch")

    def convert_diff_box_batch(self, i420, n_streams, n_frames, out, windows, boxes=None, first_stream=0, cuda_stream=None,
                               windowed=False):
        """The fused step with the object box of lib/JM/image.c:4818-4838 accumulated by the same launch (no mask in
        memory).  windows: int32 CUDA tensor [n_streams*n_frames, n_objects, 4] = {x0, x1, y0, y1} inclusive; returns the
        int32 CUDA tensor [n_streams*n_frames, n_objects, 4] = {min_x, max_x, min_y, max_y}.  out["bgr"] may be None.
        windowed=True (LV_BOX_WINDOWED): plain fused step + a pass over the windows only; same results, faster when the
        windows are small."""
        import torch
        n, d = n_streams * n_frames, self.device
        if windows.dim() != 3 or windows.shape[0] != n or windows.shape[2] != 4:
            raise ValueError("windows: expected [n_streams*n_frames, n_objects, 4]")
        n_obj = windows.shape[1]
        if boxes is None:
            boxes = torch.empty((n, n_obj, 4), dtype=torch.int32, device=torch.device("cuda", d))
        capi.check(capi.lib().lv_convert_diff_box_batch_ex(
            self._h, _dev_ptr(i420, torch.uint8, n * self.frame_bytes, "i420", d, 16), first_stream, n_streams, n_frames,
            _dev_ptr(out.get("bgr"), torch.uint8, n * self.bgr_bytes, "bgr", d, 16),
            _dev_ptr(out.get("mb_count"), torch.int16, n * self.mbs, "mb_count", d),
            _dev_ptr(out.get("mb_map"), torch.uint8, n * self.mbs, "mb_map", d),
            _dev_ptr(out.get("summary"), torch.int32, n * 2, "summary", d),
            _dev_ptr(windows, torch.int32, n * n_obj * 4, "windows", d, 16), n_obj,
            _dev_ptr(boxes, torch.int32, n * n_obj * 4, "boxes", d, 16), _stream_handle(cuda_stream), 2 if windowed else 0),
            "lv_convert_diff_box_batch_ex")
        return boxes

    def object_windows(self, objects, hor_margin=8, ver_margin=8, cuda_stream=None):
        """Windows {x0, x1, y0, y1} of objects {PosX, PosY, Width, Height} (int32 CUDA tensor [..., 4];
        lib/JM/image.c:4786-4789, margins HOR/VER_DECODING_MARGIN = 8)."""
        import torch
        win = torch.empty_like(objects)
        n = objects.numel() // 4
        capi.check(capi.lib().lv_object_windows(self._h, _dev_ptr(objects, torch.int32, 4 * n, "objects", self.device, 16), n,
                                                hor_margin, ver_margin, win.data_ptr(), _stream_handle(cuda_stream)),
                   "lv_object_windows")
        return win

    def object_update(self, boxes, objects, cuda_stream=None):
        """objects <- boxes under the reference's guard min < max (lib/JM/image.c:4840-4848); in place."""
        import torch
        n = objects.numel() // 4
        capi.check(capi.lib().lv_object_update(self._h, _dev_ptr(boxes, torch.int32, 4 * n, "boxes", self.device, 16), n,
                                               _dev_ptr(objects, torch.int32, 4 * n, "objects", self.device, 16),
                                               _stream_handle(cuda_stream)), "lv_object_update")
        return objects

    def convert_batch(self, i420, n_frames, bgr, cuda_stream=None):
        import torch
        d = self.device
        capi.check(capi.lib().lv_convert_batch(
            self._h, _dev_ptr(i420, torch.uint8, n_frames * self.frame_bytes, "i420", d, 16), n_frames,
            _dev_ptr(bgr, torch.uint8, n_frames * self.bgr_bytes, "bgr", d, 16), _stream_handle(cuda_stream)),
            "lv_convert_batch")

    def diff_batch(self, cur_y, ref_y, n_frames, out, cuda_stream=None):
        import torch
        d = self.device
        capi.check(capi.lib().lv_diff_batch(
            self._h, _dev_ptr(cur_y, torch.uint8, n_frames * self.px, "cur_y", d, 16),
            _dev_ptr(ref_y, torch.uint8, n_frames * self.px, "ref_y", d, 16), n_frames,
            _dev_ptr(out.get("mask_bits"), torch.int16, n_frames * self.mask_units, "mask_bits", d),
            _dev_ptr(out.get("mask_bytes"), torch.uint8, n_frames * self.px, "mask_bytes", d, 16),
            _dev_ptr(out.get("mb_count"), torch.int16, n_frames * self.mbs, "mb_count", d),
            _dev_ptr(out.get("mb_map"), torch.uint8, n_frames * self.mbs, "mb_map", d),
            _dev_ptr(out.get("summary"), torch.int32, n_frames * 2, "summary", d), _stream_handle(cuda_stream)),
            "lv_diff_batch")

    def write_bks(self, cur_i420, bkd_i420, n_frames, out_i420, tol=25, obj_bits=None, cuda_stream=None):
        """The reference's WriteBks (MainDlg.cpp:771-835): background-subtracted I420 frames (device tensors)."""
        import torch
        d = self.device
        capi.check(capi.lib().lv_write_bks(
            self._h, _dev_ptr(cur_i420, torch.uint8, n_frames * self.frame_bytes, "cur_i420", d),
            _dev_ptr(bkd_i420, torch.uint8, self.frame_bytes, "bkd_i420", d), n_frames, tol,
            _dev_ptr(out_i420, torch.uint8, n_frames * self.frame_bytes, "out_i420", d),
            _dev_ptr(obj_bits, torch.int16, n_frames * self.mask_units, "obj_bits", d), _stream_handle(cuda_stream)),
            "lv_write_bks")

    def object_box(self, mask, x0, x1, y0, y1, mask_is_bits=False, cuda_stream=None):
        """(min_x, max_x, min_y, max_y) of the set mask pixels inside the window (lib/JM/image.c:4818-4838)."""
        import torch
        box = torch.empty(4, dtype=torch.int32, device=torch.device("cuda", self.device))
        need = self.mask_units if mask_is_bits else self.px
        capi.check(capi.lib().lv_object_box(
            self._h, _dev_ptr(mask, torch.int16 if mask_is_bits else torch.uint8, need, "mask", self.device),
            1 if mask_is_bits else 0, x0, x1, y0, y1, box.data_ptr(), _stream_handle(cuda_stream)), "lv_object_box")
        return tuple(int(v) for v in box.cpu())

    def narrow_pictures(self, pic16, n_frames, size_x, size_y, crop_left, crop_top, out_i420, cuda_stream=None):
        """Decoder pictures (int16 CUDA tensor: Y16, U16, V16 planes per picture) -> tight 8-bit I420 frames of this
        context's geometry: the reference's img2buf + write_out_picture (lib/JM/output.c:87-178, 362-453)."""
        import torch
        samples = size_x * size_y + 2 * ((size_x // 2) * (size_y // 2))
        capi.check(capi.lib().lv_narrow_pictures(
            self._h, _dev_ptr(pic16, torch.int16, n_frames * samples, "pic16", self.device), n_frames, size_x, size_y,
            crop_left, crop_top, _dev_ptr(out_i420, torch.uint8, n_frames * self.frame_bytes, "out_i420", self.device),
            _stream_handle(cuda_stream)), "lv_narrow_pictures")

    def convert_diff_pictures(self, pic16, size_x, size_y, crop_left, crop_top, n_streams, n_frames, out, first_stream=0,
                              cuda_stream=None):
        """convert_diff_batch fed with decoder pictures (int16 CUDA tensor, narrow_pictures layout, stream-major): ONE
        kernel narrows, crops, converts and differences (lv_convert_diff_pictures).  `out` as alloc_outputs gives it;
        out["bgr"] is mandatory."""
        import torch
        n = n_streams * n_frames
        d = self.device
        samples = size_x * size_y + 2 * ((size_x // 2) * (size_y // 2))
        capi.check(capi.lib().lv_convert_diff_pictures(
            self._h, _dev_ptr(pic16, torch.int16, n * samples, "pic16", d), size_x, size_y, crop_left, crop_top, first_stream,
            n_streams, n_frames, _dev_ptr(out.get("bgr"), torch.uint8, n * self.bgr_bytes, "bgr", d, 16),
            _dev_ptr(out.get("mask_bits"), torch.int16, n * self.mask_units, "mask_bits", d),
            _dev_ptr(out.get("mask_bytes"), torch.uint8, n * self.px, "mask_bytes", d, 16),
            _dev_ptr(out.get("mb_count"), torch.int16, n * self.mbs, "mb_count", d),
            _dev_ptr(out.get("mb_map"), torch.uint8, n * self.mbs, "mb_map", d),
            _dev_ptr(out.get("summary"), torch.int32, 2 * n, "summary", d), _stream_handle(cuda_stream)),
            "lv_convert_diff_pictures")

    def object_box_batch(self, masks, n_frames, x0, x1, y0, y1, mask_is_bits=False, cuda_stream=None):
        """One box per frame for n_frames masks laid out back to back: int32 CUDA tensor [n_frames, 4]."""
        import torch
        boxes = torch.empty((n_frames, 4), dtype=torch.int32, device=torch.device("cuda", self.device))
        need = n_frames * (self.mask_units if mask_is_bits else self.px)
        capi.check(capi.lib().lv_object_box_batch(
            self._h, _dev_ptr(masks, torch.int16 if mask_is_bits else torch.uint8, need, "masks", self.device),
            1 if mask_is_bits else 0, n_frames, x0, x1, y0, y1, boxes.data_ptr(), _stream_handle(cuda_stream)),
            "lv_object_box_batch")
        return boxes

    def upload_async(self, i420_host, n_frames_total, slot):
        """Pinned host -> device slot `slot` on the context's copy stream; returns a uint8 CUDA tensor view of the slot.
        Pair with upload_wait(slot) before the consumer and upload_release(slot) after it (double buffering)."""
        import torch
        p = C.c_void_p()
        nbytes = n_frames_total * self.frame_bytes
        capi.check(capi.lib().lv_upload_async(self._h, _host_ptr(i420_host, np.uint8, nbytes, "i420_host"), n_frames_total,
                                              slot, C.byref(p)), "lv_upload_async")
        return _DevView(p.value, nbytes, self.device).tensor()

    def upload_wait(self, slot, cuda_stream=None):
        capi.check(capi.lib().lv_upload_wait(self._h, slot, _stream_handle(cuda_stream)), "lv_upload_wait")

    def upload_release(self, slot, cuda_stream=None):
        capi.check(capi.lib().lv_upload_release(self._h, slot, _stream_handle(cuda_stream)), "lv_upload_release")

    def convert_format(self, kind, d_in, d_out, n_frames, cuda_stream=None):
        """Device form of the five other cscc.lib converters (kind 1..5, see include/livevideo.h)."""
        import torch
        L = capi.lib()
        capi.check(L.lv_convert_format(
            self._h, kind, _dev_ptr(d_in, torch.uint8, n_frames * L.lv_format_bytes(kind, self.width, self.height, 0), "d_in", self.device),
            _dev_ptr(d_out, torch.uint8, n_frames * L.lv_format_bytes(kind, self.width, self.height, 1), "d_out", self.device),
            n_frames, _stream_handle(cuda_stream)), "lv_convert_format")

    # ---- host API ------------------------------------------------------------------------------------
    def step_host(self, i420, n_streams, n_frames, bgr=None, mask_bits=None, mb_count=None, mb_map=None,
                  summary=None, first_stream=0):
        """Host buffers in, host buffers out (numpy arrays, ideally from `pinned()`); synchronous."""
        n = n_streams * n_frames

        def p(a, dt, nb, name, w=True):
            return None if a is None else _host_ptr(a, dt, nb, name, w)
        capi.check(capi.lib().lv_step_host(
            self._h, _host_ptr(i420, np.uint8, n * self.frame_bytes, "i420"), first_stream, n_streams, n_frames,
            p(bgr, np.uint8, n * self.bgr_bytes, "bgr"), p(mask_bits, np.uint16, 2 * n * self.mask_units, "mask_bits"),
            p(mb_count, np.uint16, 2 * n * self.mbs, "mb_count"), p(mb_map, np.uint8, n * self.mbs, "mb_map"),
            p(summary, np.uint32, 8 * n, "summary")), "lv_step_host")

    def step_host_pictures(self, pic16, size_x, size_y, crop_left, crop_top, n_streams, n_frames, bgr=None, mask_bits=None,
                           mb_count=None, mb_map=None, summary=None, first_stream=0):
        """step_host fed with the decoder's own pictures (numpy uint16: Y16, U16, V16 planes per picture): the narrowing
        and cropping of lib/JM/output.c:87-178, 362-453 runs on the device in front of the fused kernel."""
        n = n_streams * n_frames
        samples = size_x * size_y + 2 * ((size_x // 2) * (size_y // 2))

        def p(a, dt, nb, name, w=True):
            return None if a is None else _host_ptr(a, dt, nb, name, w)
        capi.check(capi.lib().lv_step_host_pictures(
            self._h, _host_ptr(pic16, np.uint16, 2 * n * samples, "pic16"), size_x, size_y, crop_left, crop_top,
            first_stream, n_streams, n_frames,
            p(bgr, np.uint8, n * self.bgr_bytes, "bgr"), p(mask_bits, np.uint16, 2 * n * self.mask_units, "mask_bits"),
            p(mb_count, np.uint16, 2 * n * self.mbs, "mb_count"), p(mb_map, np.uint8, n * self.mbs, "mb_map"),
            p(summary, np.uint32, 8 * n, "summary")), "lv_step_host_pictures")


class _DevView:
    """A torch uint8 tensor over device memory owned by the library (no copy), via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes, device):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        self._device = device

    def tensor(self):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", self._device))


class PinnedArray:
    """numpy view of pinned host memory from lv_host_alloc (freed with the object)."""

    def __init__(self, nbytes, dtype=np.uint8):
        self._p = C.c_void_p()
        capi.check(capi.lib().lv_host_alloc(C.byref(self._p), nbytes), "lv_host_alloc")
        buf = (C.c_uint8 * nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype)

    def __del__(self):
        try:
            if self._p:
                capi.lib().lv_host_free(self._p)
                self._p = C.c_void_p()
        except Exception:  # noqa: BLE001
            pass


def pinned(nbytes, dtype=np.uint8):
    return PinnedArray(nbytes, dtype)


def host_register(a):
    """Pin a numpy array the caller already owns (lv_host_register); keep it alive until host_unregister(a)."""
    capi.check(capi.lib().lv_host_register(a.ctypes.data, a.nbytes), "lv_host_register")


def host_is_mapped(a):
    """True when the numpy array lies in memory pinned and mapped through this library (the legacy entries' zero-copy path)."""
    return bool(capi.lib().lv_host_is_mapped(a.ctypes.data, a.nbytes))


def host_unregister(a):
    capi.check(capi.lib().lv_host_unregister(a.ctypes.data), "lv_host_unregister")


KERNEL_VARIANTS = {"tile": 0, "tma": 1, "tile_late": 2, "tile_early": 3}


def set_kernel_variant(name):
    """Select the kernel variant ("tile" default, "tma", "tile_late" / "tile_early" = tile with one order of its chroma
    loads forced); all are bit-identical."""
    capi.check(capi.lib().lv_set_kernel_variant(KERNEL_VARIANTS[name]), "lv_set_kernel_variant")


def get_kernel_variant():
    v = capi.lib().lv_get_kernel_variant()
    return {n: k for k, n in KERNEL_VARIANTS.items()}[v]


ZERO_COPY_MODES = {"off": 0, "plain": 1, "staged": 2}


def set_zero_copy(name):
    """Zero-copy mode of the host-pointer conversion entries on buffers pinned through this library: "off" (always stage
    through device buffers), "plain" (tile kernel's store pattern) or "staged" (default: 512 contiguous bytes per warp
    store).  Results are identical in every mode."""
    capi.check(capi.lib().lv_set_zero_copy(ZERO_COPY_MODES[name]), "lv_set_zero_copy")


def get_zero_copy():
    v = capi.lib().lv_get_zero_copy()
    return {n: k for k, n in ZERO_COPY_MODES.items()}[v]


def host_step_plan(width, height, n_streams, n_frames, in_frame_bytes=0):
    """lv_host_step_plan: the chunks (first stream, streams, first frame, frames) a host step would be cut into under the
    current environment (LV_HOST_CHUNK_MB, LV_HOST_SLOT_MB, LV_HOST_RAMP); no device is touched."""
    L = capi.lib()
    n = L.lv_host_step_plan(width, height, in_frame_bytes, n_streams, n_frames, None, 0)
    capi.check(min(n, 0), "lv_host_step_plan")
    buf = (C.c_int * (4 * max(n, 1)))()
    n2 = L.lv_host_step_plan(width, height, in_frame_bytes, n_streams, n_frames, C.cast(buf, C.c_void_p), n)
    assert n2 == n
    return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]


def gen_synthetic(kind, seed, first_stream, first_frame, n_streams, n_frames, width, height, out=None, device=0,
                  cuda_stream=None):
    """Synthetic I420 frames on the device (`uniform` or `camera` generator)."""
    import torch
    n = n_streams * n_frames
    fb = width * height * 3 // 2
    if out is None:
        out = torch.empty(n * fb, dtype=torch.uint8, device=torch.device("cuda", device))
    with torch.cuda.device(device):
        capi.check(capi.lib().lv_gen_synthetic({"uniform": 0, "camera": 1}[kind], seed, first_stream, first_frame,
                                               n_streams, n_frames, width, height,
                                               _dev_ptr(out, torch.uint8, n * fb, "out", device),
                                               _stream_handle(cuda_stream)), "lv_gen_synthetic")
    return out
