"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module (see oracle/lv_oracle.h).  The product package livevideo_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblv_oracle.so")
ORACLE32 = os.path.join(_HERE, "_ref", "oracle32")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)


def build(force=False):
    """Compile the C restatement (and oracle32 when /root/reference is present)."""
    src = os.path.join(_HERE, "lv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "lv_oracle.h"))):
        subprocess.run(["gcc", "-O3", "-fPIC", "-std=c11", "-pthread", "-shared", "-o", _LIB_PATH, src,
                        "-lpthread"], check=True, cwd=_HERE)
    if os.path.exists("/root/reference/cscc.lib") and (force or not os.path.exists(ORACLE32)):
        subprocess.run(["python3", os.path.join(_HERE, "ref32", "build_ref.py")], check=False)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.lvo_yv12_to_rgb24.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_int]
        L.lvo_yuv_to_bgr_px.argtypes = [C.c_int, C.c_int, C.c_int, _u8p]
        L.lvo_textbook_px.argtypes = [C.c_int, C.c_int, C.c_int, _u8p]
        L.lvo_is_nearly_zero.argtypes = [C.c_int, C.c_int]
        L.lvo_is_nearly_zero.restype = C.c_int
        L.lvo_frame_diff_mb.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u16p, _u8p]
        L.lvo_frame_diff_mb.restype = C.c_uint32
        L.lvo_pack_mask16.argtypes = [_u8p, C.c_int, C.c_int, _u16p]
        L.lvo_convert_diff_batch.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             _u8p, _u8p, _u16p, _u8p, _u32p, C.c_int]
        L.lvo_write_bks.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.lvo_object_box.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        for name in ("lvo_rgb24_to_yv12", "lvo_yvu9_to_yv12", "lvo_yuy2_to_yv12", "lvo_yv12_to_yvu9", "lvo_yv12_to_yuy2"):
            getattr(L, name).argtypes = [_u8p, _u8p, C.c_int, C.c_int]
        L.lvo_img2buf.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.lvo_picture16_to_i420.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.lvo_lowbias32.argtypes = [C.c_uint32]
        L.lvo_lowbias32.restype = C.c_uint32
        L.lvo_gen_uniform.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u8p]
        L.lvo_gen_camera.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u8p]
        L.lvo_crc32.argtypes = [_u8p, C.c_uint64]
        L.lvo_crc32.restype = C.c_uint32
        L.lvo_init_tables()
        _lib = L
    return _lib


def _p(a, t=_u8p):
    return None if a is None else a.ctypes.data_as(t)


def frame_bytes(w, h):
    return w * h * 3 // 2


def mb_dims(w, h):
    return (w + 15) // 16, (h + 15) // 16


def yv12_to_rgb24(y, u, v, w, h):
    y, u, v = (np.ascontiguousarray(a, dtype=np.uint8) for a in (y, u, v))
    out = np.empty(3 * w * h, dtype=np.uint8)
    lib().lvo_yv12_to_rgb24(_p(y), _p(u), _p(v), _p(out), w, h)
    return out


def convert_frame(i420, w, h):
    i420 = np.ascontiguousarray(i420, dtype=np.uint8).ravel()
    px = w * h
    return yv12_to_rgb24(i420[:px], i420[px:px + px // 4], i420[px + px // 4:px + px // 2], w, h)


def px(y, u, v):
    out = np.zeros(3, dtype=np.uint8)
    lib().lvo_yuv_to_bgr_px(int(y), int(u), int(v), _p(out))
    return tuple(int(x) for x in out)


def textbook_px(y, u, v):
    out = np.zeros(3, dtype=np.uint8)
    lib().lvo_textbook_px(int(y), int(u), int(v), _p(out))
    return tuple(int(x) for x in out)


def frame_diff_mb(cur, ref, w, h, tol, mb_thresh):
    cur = np.ascontiguousarray(cur, dtype=np.uint8).ravel()
    ref = np.ascontiguousarray(ref, dtype=np.uint8).ravel()
    mbw, mbh = mb_dims(w, h)
    mask = np.empty(w * h, dtype=np.uint8)
    cnt = np.empty(mbw * mbh, dtype=np.uint16)
    mp = np.empty(mbw * mbh, dtype=np.uint8)
    total = lib().lvo_frame_diff_mb(_p(cur), _p(ref), w, h, tol, mb_thresh, _p(mask), _p(cnt, _u16p), _p(mp))
    return mask, cnt, mp, int(total)


def pack_mask16(mask, w, h):
    mask = np.ascontiguousarray(mask, dtype=np.uint8).ravel()
    bits = np.empty((w // 16) * h, dtype=np.uint16)
    lib().lvo_pack_mask16(_p(mask), w, h, _p(bits, _u16p))
    return bits


def convert_diff_batch(i420, prev_luma, n_streams, n_frames, w, h, tol, mb_thresh,
                       want_bgr=True, want_mask=True, n_threads=1):
    """Returns dict(bgr, mask, mb_count, mb_map, summary); prev_luma is updated in place."""
    i420 = np.ascontiguousarray(i420, dtype=np.uint8).ravel()
    assert prev_luma.dtype == np.uint8 and prev_luma.flags.c_contiguous
    n = n_streams * n_frames
    mbw, mbh = mb_dims(w, h)
    bgr = np.empty(n * 3 * w * h, dtype=np.uint8) if want_bgr else None
    mask = np.empty(n * w * h, dtype=np.uint8) if want_mask else None
    cnt = np.empty(n * mbw * mbh, dtype=np.uint16)
    mp = np.empty(n * mbw * mbh, dtype=np.uint8)
    summ = np.empty(n * 2, dtype=np.uint32)
    lib().lvo_convert_diff_batch(_p(i420), _p(prev_luma), n_streams, n_frames, w, h, tol, mb_thresh,
                                 _p(bgr), _p(mask), _p(cnt, _u16p), _p(mp), _p(summ, _u32p), n_threads)
    return dict(bgr=bgr, mask=mask, mb_count=cnt, mb_map=mp, summary=summ.reshape(n, 2))


def write_bks(cur, bkd, w, h, tol):
    cur = np.ascontiguousarray(cur, dtype=np.uint8).ravel()
    bkd = np.ascontiguousarray(bkd, dtype=np.uint8).ravel()
    out = np.empty_like(cur)
    lib().lvo_write_bks(_p(cur), _p(bkd), w, h, tol, _p(out))
    return out


def object_box(mask, w, x0, x1, y0, y1):
    mask = np.ascontiguousarray(mask, dtype=np.uint8).ravel()
    box = (C.c_int * 4)()
    lib().lvo_object_box(_p(mask), w, x0, x1, y0, y1, box)
    return tuple(box)


def object_window(obj, w, h, hor_margin=8, ver_margin=8):
    """{min_x, max_x, min_y, max_y} of an object {PosX, PosY, Width, Height} (lib/JM/image.c:4786-4789)."""
    L = lib()
    L.lvo_object_window.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    o, win = (C.c_int * 4)(*[int(v) for v in obj]), (C.c_int * 4)()
    L.lvo_object_window(o, w, h, hor_margin, ver_margin, win)
    return tuple(win)


def object_update(box, obj):
    """Object {PosX, PosY, Width, Height} after the box (lib/JM/image.c:4840-4848)."""
    L = lib()
    L.lvo_object_update.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    b, o = (C.c_int * 4)(*[int(v) for v in box]), (C.c_int * 4)(*[int(v) for v in obj])
    L.lvo_object_update(b, o)
    return tuple(o)


# converter name -> (oracle32 mode, input bytes per 8 px, output bytes per 8 px)
FORMATS = {"rgb24_to_yv12": (1, 24, 12), "yvu9_to_yv12": (2, 9, 12), "yuy2_to_yv12": (3, 16, 12),
           "yv12_to_yvu9": (4, 12, 9), "yv12_to_yuy2": (5, 12, 16)}


def convert_format(name, data, w, h):
    """One of the five other cscc.lib converters (C restatement)."""
    mode, inb, outb = FORMATS[name]
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    assert data.size >= w * h // 8 * inb
    out = np.zeros(w * h // 8 * outb, dtype=np.uint8)
    getattr(lib(), "lvo_" + name)(_p(data), _p(out), w, h)
    return out


def img2buf(plane16, size_x, size_y, cl, cr, ct, cb):
    """lib/JM/output.c:87-178 (1-byte symbols from unsigned-short samples): low byte + crop."""
    plane16 = np.ascontiguousarray(plane16, dtype=np.uint16).ravel()
    out = np.empty((size_x - cl - cr) * (size_y - ct - cb), dtype=np.uint8)
    lib().lvo_img2buf(_p(plane16, _u16p), size_x, size_y, cl, cr, ct, cb, _p(out))
    return out


def picture16_to_i420(pic16, size_x, size_y, cl, cr, ct, cb):
    """write_out_picture's Y, U, V order and crops (output.c:362-453) over one decoder picture of 16-bit planes."""
    pic16 = np.ascontiguousarray(pic16, dtype=np.uint16).ravel()
    w, h = size_x - cl - cr, size_y - ct - cb
    out = np.empty(w * h * 3 // 2, dtype=np.uint8)
    lib().lvo_picture16_to_i420(_p(pic16, _u16p), size_x, size_y, cl, cr, ct, cb, _p(out))
    return out


def jm_img2buf(plane16, size_x, size_y, cl, cr, ct, cb):
    """The reference's OWN compiled img2buf (oracle/_ref/jm_ldecod --img2buf), or None when the binary is absent."""
    exe = os.path.join(_HERE, "_ref", "jm_ldecod")
    if not os.path.exists(exe):
        return None
    data = np.ascontiguousarray(plane16, dtype=np.uint16).tobytes()
    r = subprocess.run([exe, "--img2buf", str(size_x), str(size_y), str(cl), str(cr), str(ct), str(cb)], input=data,
                       capture_output=True, check=True)
    return np.frombuffer(r.stdout, dtype=np.uint8)


def lowbias32(x):
    return int(lib().lvo_lowbias32(x & 0xFFFFFFFF))


def gen(kind, seed, stream, frame, w, h):
    out = np.empty(frame_bytes(w, h), dtype=np.uint8)
    f = lib().lvo_gen_uniform if kind == "uniform" else lib().lvo_gen_camera
    f(seed, stream, frame, w, h, _p(out))
    return out


def gen_batch(kind, seed, n_streams, n_frames, w, h, first_frame=0, first_stream=0):
    out = np.empty((n_streams, n_frames, frame_bytes(w, h)), dtype=np.uint8)
    for s in range(n_streams):
        for t in range(n_frames):
            out[s, t] = gen(kind, seed, first_stream + s, first_frame + t, w, h)
    return out


def crc32(a):
    a = np.ascontiguousarray(a, dtype=np.uint8).ravel()
    return int(lib().lvo_crc32(_p(a), a.size))


def have_oracle32():
    """True when the reference's original i386 object code can execute on this host."""
    if not os.path.exists(ORACLE32):
        return False
    try:
        r = subprocess.run([ORACLE32, "0", "4", "2"], input=bytes([16] * 8 + [128] * 4), capture_output=True,
                           timeout=20)
        return r.returncode == 0 and r.stdout[:3] == b"\0\0\0"
    except (OSError, subprocess.SubprocessError):
        return False


def oracle32_convert(i420_frames, w, h, mode=0):
    """Run frames through the reference's ORIGINAL object code (cscc.lib Convert.obj)."""
    data = np.ascontiguousarray(i420_frames, dtype=np.uint8).tobytes()
    r = subprocess.run([ORACLE32, str(mode), str(w), str(h)], input=data, capture_output=True, check=True)
    return np.frombuffer(r.stdout, dtype=np.uint8)


def oracle32_time(i420_frame, w, h, reps):
    """(seconds, output) for `reps` in-process calls of the original YV12_to_RGB24."""
    data = np.ascontiguousarray(i420_frame, dtype=np.uint8).tobytes()
    r = subprocess.run([ORACLE32, "0", str(w), str(h), str(reps)], input=data, capture_output=True, check=True)
    txt = r.stderr.decode()
    s = int(txt.split("s=")[1].split()[0])
    ns = int(txt.split("ns=")[1].split()[0])
    return s + ns * 1e-9, np.frombuffer(r.stdout, dtype=np.uint8)
