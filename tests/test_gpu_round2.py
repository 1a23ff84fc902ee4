"""GPU parity tests of the round-2 entries, through the C ABI, against the oracle (run on the B200 box: -m gpu).

* lv_convert_diff_box_batch / lv_object_windows / lv_object_update -- motion_segment's I-frame step
  (lib/JM/image.c:4783-4848) with the box accumulated by the fused kernel
* lv_convert_host_multi / lv_diff_host_multi -- CMainDlg::DisplayPic's five conversions in one call (MainDlg.cpp:937-977)
* per-stream state: partial-stream steps touch nobody else's previous luma (lib/JM/image.c:1553-1561 per stream)
* alignment and pitch limits are error codes / fallbacks, never a crashed context
* the peer exchange across two PROCESSES on one GPU (CUDA IPC), its gate and its one-deadline give-up
"""
import ctypes as C
import mmap
import os
import socket
import sys
import time

import numpy as np
import pytest

from test_gpu_parity import assert_matches_oracle, run_fused

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


# ---- the fused object box ---------------------------------------------------------------------------------------------
def _oracle_boxes(orc, mask, w, h, wins):
    """wins: [n, k, 4] -> boxes [n, k, 4] by the reference's scan (lvo_object_box) on the oracle's mask."""
    n, k, _ = wins.shape
    px = w * h
    out = np.zeros((n, k, 4), np.int32)
    for f in range(n):
        for j in range(k):
            x0, x1, y0, y1 = (int(v) for v in wins[f, j])
            out[f, j] = orc.object_box(mask[f * px:(f + 1) * px], w, x0, x1, y0, y1)
    return out


@pytest.mark.parametrize("w,h,ns,nf,k", [(352, 288, 2, 3, 1), (1280, 720, 1, 2, 3), (1920, 1080, 1, 2, 16), (64, 48, 3, 2, 2),
                                         (720, 576, 1, 2, 2), (32752, 64, 1, 2, 2), (16, 4096, 1, 2, 3)])
def test_fused_box_matches_reference_scan(lv, orc, torch_cuda, w, h, ns, nf, k):
    """Boxes from the fused launch == lvo_object_box on the mask the same launch would have written; every other output
    of the step is unchanged; no mask buffer is allocated."""
    torch = torch_cuda
    n = ns * nf
    frames = orc.gen_batch("camera", 3, ns, nf, w, h)
    prev = np.zeros((ns, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
    rng = np.random.default_rng(w + k)
    wins = np.zeros((n, k, 4), np.int32)
    for f in range(n):
        for j in range(k):
            xa, xb = sorted(int(v) for v in rng.integers(0, w, 2))
            ya, yb = sorted(int(v) for v in rng.integers(0, h, 2))
            wins[f, j] = (xa, xb, ya, yb)
    wins[0, 0] = (0, w - 1, 0, h - 1)                    # the whole picture
    wins[-1, -1] = (5, 5, 7, 7)                          # a single pixel
    if n > 1:
        wins[1, 0] = (w - 1, w - 1, 0, h - 1)            # the last column
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    for windowed in (False, True):            # the fused box, and LV_BOX_WINDOWED: plain step + a pass over the windows
        ctx = lv.Context(w, h, ns, 20, 0)
        out = ctx.alloc_outputs(n)
        boxes = ctx.convert_diff_box_batch(d_in, ns, nf, out, torch.from_numpy(wins).cuda(), windowed=windowed)
        torch.cuda.synchronize()
        assert np.array_equal(boxes.cpu().numpy(), _oracle_boxes(orc, ref["mask"], w, h, wins)), windowed
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
        assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"])
        assert np.array_equal(out["mb_map"].cpu().numpy(), ref["mb_map"])
        assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"])
        for s in range(ns):
            assert np.array_equal(ctx.get_prev_luma(s), prev[s])
        ctx.close()


def test_fused_box_empty_windows_and_static_background_loop(lv, orc, torch_cuda):
    """The literal I-frame step: static background BkdBuf as the reference (MainDlg.cpp:238-246), windows from the objects'
    last position (image.c:4786-4789), box, then PosX/PosY/Width/Height under min < max (image.c:4840-4848) -- all on the
    device, no BGR, no mask; a window without a changed pixel returns the reference's initial values."""
    torch = torch_cuda
    w, h, nf = 352, 288, 4
    px = w * h
    bg = orc.gen("camera", 9, 0, 0, w, h)
    frames = np.stack([bg.copy() for _ in range(nf)])
    # a bright 40x24 object moving right, on an otherwise unchanged background; frame 3 has no object at all
    pos = [(100, 60), (110, 62), (120, 64)]
    for t, (x, y) in enumerate(pos):
        yy = frames[t, :px].reshape(h, w)
        yy[y:y + 24, x:x + 40] = 250
    objects = np.array([[[118, 70, 30, 20], [300, 250, 10, 10]]] * nf, np.int32)      # object 1 never sees a change
    ctx = lv.Context(w, h, 1, 20, 0)
    ctx.set_ref_mode(True)
    ctx.set_prev_luma(0, bg[:px].copy())
    d_obj = torch.from_numpy(objects.reshape(nf, 2, 4)).cuda()
    wins = ctx.object_windows(d_obj)
    out = ctx.alloc_outputs(nf, bgr=False)
    boxes = ctx.convert_diff_box_batch(torch.from_numpy(frames.reshape(-1)).cuda(), 1, nf, out, wins)
    # the same step with the boxes from the pass over the windows (static reference, no BGR): identical boxes
    boxes_w = ctx.convert_diff_box_batch(torch.from_numpy(frames.reshape(-1)).cuda(), 1, nf, ctx.alloc_outputs(nf, bgr=False), wins,
                                         windowed=True)
    assert torch.equal(boxes, boxes_w)
    ctx.object_update(boxes, d_obj)
    torch.cuda.synchronize()
    # oracle: the reference's loops
    want_w = np.zeros((nf, 2, 4), np.int32)
    want_b = np.zeros((nf, 2, 4), np.int32)
    want_o = objects.copy()
    for t in range(nf):
        mask = orc.frame_diff_mb(frames[t, :px], bg[:px], w, h, 20, 0)[0]
        for j in range(2):
            want_w[t, j] = orc.object_window(objects[t, j], w, h)
            x0, x1, y0, y1 = (int(v) for v in want_w[t, j])
            want_b[t, j] = orc.object_box(mask, w, x0, x1, y0, y1)
            want_o[t, j] = orc.object_update(want_b[t, j], objects[t, j])
    assert np.array_equal(wins.cpu().numpy(), want_w)
    assert np.array_equal(boxes.cpu().numpy(), want_b)
    assert np.array_equal(d_obj.cpu().numpy(), want_o)
    # sanity of the scenario itself: the object was found in frames 0..2 and object 1 / frame 3 stayed as they were
    assert want_b[0, 0].tolist() == [100, 139, 60, 83] and want_o[0, 0].tolist() == [119, 71, 39, 23]
    x0, x1, y0, y1 = want_w[3, 0]
    assert want_b[3, 0].tolist() == [x1 + 1, x0 - 1, y1 + 1, y0 - 1] and want_o[3, 0].tolist() == objects[3, 0].tolist()
    assert np.array_equal(ctx.get_prev_luma(0), bg[:px])         # the static background is left untouched
    ctx.close()


# ---- several pictures per call ----------------------------------------------------------------------------------------
def _own(nb):
    return np.frombuffer(mmap.mmap(-1, nb), dtype=np.uint8, count=nb)


@pytest.mark.parametrize("w,h,n,register", [(352, 288, 5, False), (1920, 1080, 5, True), (1280, 720, 7, False), (20, 6, 4, False)])
def test_convert_host_multi_is_n_single_calls(lv, orc, torch_cuda, w, h, n, register):
    """CMainDlg::DisplayPic (MainDlg.cpp:946-974): five distinct pictures and destinations through one call; pageable
    and registered buffers; more pictures than pipeline slots; a width that is not a multiple of 16."""
    px = w * h
    frames = orc.gen_batch("uniform", 21, n, 1, w, h).reshape(n, -1)
    ys, us, vs, ds = [], [], [], []
    for i in range(n):
        y, u, v, d = _own(px), _own(px // 4), _own(px // 4), _own(3 * px)
        y[:], u[:], v[:] = frames[i, :px], frames[i, px:px + px // 4], frames[i, px + px // 4:]
        ys.append(y), us.append(u), vs.append(v), ds.append(d)
    bufs = ys + us + vs + ds
    if register:
        for a in bufs:
            lv.host_register(a)
    csc = lv.ColorSpaceConversions()
    for rep in range(2):
        for d in ds:
            d[:] = 0
        csc.YV12_to_RGB24_multi(ys, us, vs, ds, w, h)
        for i in range(n):
            assert np.array_equal(ds[i], orc.convert_frame(frames[i], w, h)), (rep, i)
    if register:
        for a in bufs:
            lv.host_unregister(a)
    with pytest.raises(lv.LiveVideoError):
        lv.capi.check(lv.lib().lv_convert_host_multi(None, None, None, None, 3, w, h), "null arrays")


def test_diff_host_multi_chained_and_explicit_references(lv, orc, torch_cuda):
    """FrameDiff_MB over a sequence in one call: ref[i] == cur[i-1] (every plane uploaded once) and arbitrary references
    give the oracle's masks, counts and maps; optional outputs may be missing."""
    w, h, n = 1280, 720, 6
    px = w * h
    fr = orc.gen_batch("camera", 4, 1, n + 1, w, h).reshape(n + 1, -1)
    planes = [np.ascontiguousarray(fr[i, :px]) for i in range(n + 1)]
    cur, ref = planes[1:], planes[:-1]                                   # chained: ref[i] is cur[i-1] (same object)
    mask, cnt, mp = lv.FrameDiff_MB_multi(cur, ref, w, h, 20, 0)
    for i in range(n):
        m, c, b, _ = orc.frame_diff_mb(cur[i], ref[i], w, h, 20, 0)
        assert np.array_equal(mask[i], m) and np.array_equal(cnt[i], c) and np.array_equal(mp[i], b), i
    ref2 = [planes[(3 * i) % (n + 1)].copy() for i in range(n)]           # unrelated references, separate buffers
    mask, cnt, mp = lv.FrameDiff_MB_multi(cur, ref2, w, h, 7, 3)
    for i in range(n):
        m, c, b, _ = orc.frame_diff_mb(cur[i], ref2[i], w, h, 7, 3)
        assert np.array_equal(mask[i], m) and np.array_equal(cnt[i], c) and np.array_equal(mp[i], b), i
    # counts only (no mask, no map)
    cnt_only = [np.zeros(cnt[0].size, np.uint16) for _ in range(n)]
    arr = lambda xs: (C.c_void_p * n)(*[a.ctypes.data for a in xs])   # noqa: E731
    lv.capi.check(lv.lib().lv_diff_host_multi(arr(cur), arr(ref2), n, w, h, 7, 3, None, arr(cnt_only), None), "counts only")
    for i in range(n):
        assert np.array_equal(cnt_only[i], cnt[i])


# ---- zero-copy: the conversion entries on buffers pinned + mapped in place --------------------------------------------
@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (352, 288), (16, 2), (528, 34)])
def test_zero_copy_entries_match_oracle_in_every_mode(lv, orc, torch_cuda, w, h):
    """YV12_to_RGB24, its five-picture form (MainDlg.cpp:946-974; 9 pictures = two launches) and FrameDiff_MB on buffers
    pinned in place: the kernel addresses host memory directly (lv_zerocopy.cu).  Same bytes with zero copy off, with the
    tile kernel's stores and with the re-dealt stores; a buffer that is NOT pinned, or a misaligned plane, silently takes
    the staged path."""
    px = w * h
    n = 9
    frames = orc.gen_batch("uniform", 33, n, 1, w, h).reshape(n, -1)
    ys, us, vs, ds = [], [], [], []
    for i in range(n):
        y, u, v, d = _own(px), _own(max(px // 4, 1)), _own(max(px // 4, 1)), _own(3 * px)
        y[:], u[:px // 4], v[:px // 4] = frames[i, :px], frames[i, px:px + px // 4], frames[i, px + px // 4:]
        ys.append(y), us.append(u), vs.append(v), ds.append(d)
    mask, cnt, mp = _own(px), np.frombuffer(mmap.mmap(-1, 2 * 4096 * 4), dtype=np.uint16), _own(4096 * 4)
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    mmask = [_own(px) for _ in range(3)]
    mcnt = [np.frombuffer(mmap.mmap(-1, 2 * 4096 * 4), dtype=np.uint16) for _ in range(3)]
    mmap_ = [_own(4096 * 4) for _ in range(3)]
    bufs = ys + us + vs + ds + [mask, cnt, mp] + mmask + mcnt + mmap_
    for a in bufs:
        lv.host_register(a)
    want = [orc.convert_frame(frames[i], w, h) for i in range(n)]
    wm, wc, wb, _ = orc.frame_diff_mb(ys[1], ys[0], w, h, 20, 0)
    csc = lv.ColorSpaceConversions()
    before = lv.get_zero_copy()
    try:
        for mode in ("off", "plain", "staged"):
            lv.set_zero_copy(mode)
            assert lv.get_zero_copy() == mode
            for d in ds:
                d[:] = 0
            csc.YV12_to_RGB24(ys[0], us[0][:px // 4], vs[0][:px // 4], ds[0], w, h)
            assert np.array_equal(ds[0], want[0]), mode
            ds[0][:] = 0
            csc.YV12_to_RGB24_multi(ys, [u[:px // 4] for u in us], [v[:px // 4] for v in vs], ds, w, h)
            for i in range(n):
                assert np.array_equal(ds[i], want[i]), (mode, i)
            mask[:] = 7
            cnt[:] = 7
            mp[:] = 7
            lv.FrameDiff_MB(ys[1], ys[0], w, h, 20, 0, mask=mask, mb_count=cnt[:mbs], mb_map=mp[:mbs])
            assert np.array_equal(mask, wm) and np.array_equal(cnt[:mbs], wc) and np.array_equal(mp[:mbs], wb), mode
            assert cnt[mbs] == 7 and mp[mbs] == 7                      # nothing written past the outputs
            # the same over a sequence in one call (lv_diff_host_multi): every buffer registered
            for m in mmask:
                m[:] = 9
            lv.FrameDiff_MB_multi(ys[1:4], ys[0:3], w, h, 20, 0, mask=mmask, mb_count=[c[:mbs] for c in mcnt], mb_map=[b[:mbs] for b in mmap_])
            for i in range(3):
                m_, c_, b_, _ = orc.frame_diff_mb(ys[i + 1], ys[i], w, h, 20, 0)
                assert np.array_equal(mmask[i], m_) and np.array_equal(mcnt[i][:mbs], c_) and np.array_equal(mmap_[i][:mbs], b_), (mode, i)
        # one pageable destination / one plane at an odd offset: the staged path serves the call, same bytes
        lv.set_zero_copy("staged")
        pageable = np.zeros(3 * px, np.uint8)
        csc.YV12_to_RGB24(ys[2], us[2][:px // 4], vs[2][:px // 4], pageable, w, h)
        assert np.array_equal(pageable, want[2])
        if px > 64:
            odd = _own(px + 64)
            lv.host_register(odd)
            odd[4:4 + px] = ys[3]
            ds[3][:] = 0
            csc.YV12_to_RGB24(odd[4:4 + px], us[3][:px // 4], vs[3][:px // 4], ds[3], w, h)
            assert np.array_equal(ds[3], want[3])
            lv.host_unregister(odd)
    finally:
        lv.set_zero_copy(before)
        for a in bufs:
            lv.host_unregister(a)


# ---- decoder pictures straight into the fused kernel ------------------------------------------------------------------
@pytest.mark.parametrize("sx,sy,cl,ct,w,h,ns,nf", [(1920, 1088, 0, 0, 1920, 1080, 1, 3), (384, 304, 16, 8, 352, 288, 2, 4),
                                                   (64, 48, 0, 0, 64, 48, 3, 2), (1296, 736, 8, 6, 1280, 720, 1, 2),
                                                   (96, 64, 32, 16, 64, 48, 2, 3), (560, 64, 16, 2, 528, 34, 1, 3)])
def test_convert_diff_pictures_is_narrow_then_step(lv, orc, torch_cuda, sx, sy, cl, ct, w, h, ns, nf):
    """lv_convert_diff_pictures (16-bit decoder pictures -> ONE kernel) == the reference's img2buf + write_out_picture
    (lib/JM/output.c:87-178, 362-453) followed by the fused step, bit for bit: BGR, masks, counts, map, summary and the
    carried state, over two calls (state carry), full 16-bit content (the upper byte is dropped, not clipped), layouts the
    one-kernel path takes and one it does not (crop_left = 8: served through the scratch)."""
    torch = torch_cuda
    rng = np.random.default_rng(sx * 7 + w)
    samples = sx * sy * 3 // 2
    ctx = lv.Context(w, h, ns)
    prev = np.zeros((ns, w * h), np.uint8)
    px = w * h
    n = ns * nf
    for call in range(2):
        pics = rng.integers(0, 65536, (ns, nf, samples), dtype=np.uint16)
        pics[:, 1::2, :sx * sy] = (pics[:, 0:-1:2, :sx * sy] & 0xff00) | ((pics[:, 0:-1:2, :sx * sy] + rng.integers(0, 40, (ns, (nf) // 2, sx * sy), dtype=np.uint16)) & 0xff)
        frames = np.stack([[orc.picture16_to_i420(pics[s, t], sx, sy, cl, sx - cl - w, ct, sy - ct - h) for t in range(nf)]
                           for s in range(ns)])
        prev_in = prev.copy()                           # the oracle advances `prev` in place
        ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
        d_pic = torch.from_numpy(np.ascontiguousarray(pics).view(np.int16).reshape(-1)).cuda()
        out = ctx.alloc_outputs(n, mask_bits=True, mask_bytes=True)
        ctx.convert_diff_pictures(d_pic, sx, sy, cl, ct, ns, nf, out)
        torch.cuda.synchronize()
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"]), call
        assert np.array_equal(out["mask_bytes"].cpu().numpy(), ref["mask"]), call
        bits = np.concatenate([orc.pack_mask16(ref["mask"][i * px:(i + 1) * px], w, h) for i in range(n)])
        assert np.array_equal(out["mask_bits"].cpu().numpy().view(np.uint16), bits), call
        assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"]), call
        assert np.array_equal(out["mb_map"].cpu().numpy(), ref["mb_map"]), call
        assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"]), call
        # without the masks (the common instantiation), with either order of its loads forced and the default rule
        for variant in ("tile_late", "tile_early", "tile"):
            lv.set_kernel_variant(variant)
            try:
                ctx2 = lv.Context(w, h, ns)
                for s in range(ns):
                    ctx2.set_prev_luma(s, prev_in[s])
                out2 = ctx2.alloc_outputs(n)
                ctx2.convert_diff_pictures(d_pic, sx, sy, cl, ct, ns, nf, out2)
                torch.cuda.synchronize()
                assert np.array_equal(out2["bgr"].cpu().numpy(), ref["bgr"]), variant
                assert np.array_equal(out2["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"]), variant
                assert np.array_equal(out2["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"]), variant
                assert np.array_equal(ctx2.get_prev_luma(0), frames[0, -1, :px]), variant
                ctx2.close()
            finally:
                lv.set_kernel_variant("tile")
        prev = frames[:, -1, :px].copy()
        for s in range(ns):
            assert np.array_equal(ctx.get_prev_luma(s), prev[s]), (call, s)
    with pytest.raises(lv.LiveVideoError):
        ctx.convert_diff_pictures(d_pic, sx, sy, cl + 1, ct, ns, nf, out)        # odd crop
    ctx.close()


# ---- per-stream state -------------------------------------------------------------------------------------------------
def test_partial_stream_steps_touch_no_other_state(lv, orc, torch_cuda):
    """Streams stepped in arbitrary sub-ranges and orders keep independent previous-luma state (image.c:1553-1561 per
    stream): every stream's outputs equal the oracle run stream by stream, and the device moves no state planes."""
    torch = torch_cuda
    w, h, S, T = 64, 48, 6, 2
    px = w * h
    ctx = lv.Context(w, h, S, 20, 0)
    prev = np.zeros((S, px), np.uint8)
    steps = [(0, 6), (2, 3), (0, 1), (4, 2), (1, 4), (5, 1), (0, 6), (3, 1), (2, 4)]
    frame_no = [0] * S
    for k, (first, cnt) in enumerate(steps):
        frames = np.stack([orc.gen_batch("camera", 5, 1, T, w, h, first_frame=frame_no[s], first_stream=s)[0]
                           for s in range(first, first + cnt)])
        for s in range(first, first + cnt):
            frame_no[s] += T
        sub = prev[first:first + cnt].copy()
        ref = orc.convert_diff_batch(frames, sub, cnt, T, w, h, 20, 0, want_mask=False)
        prev[first:first + cnt] = sub
        out = ctx.alloc_outputs(cnt * T)
        ctx.convert_diff_batch(torch.from_numpy(frames.reshape(-1)).cuda(), cnt, T, out, first_stream=first)
        torch.cuda.synchronize()
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"]), k
        assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"]), k
        assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"]), k
        for s in range(S):
            assert np.array_equal(ctx.get_prev_luma(s), prev[s]), (k, s)
    # set / reset still address the live plane of each stream
    y = np.arange(px, dtype=np.uint32).astype(np.uint8)
    ctx.set_prev_luma(3, y)
    assert np.array_equal(ctx.get_prev_luma(3), y) and np.array_equal(ctx.get_prev_luma(2), prev[2])
    ctx.reset_prev_luma()
    torch.cuda.synchronize()
    for s in range(S):
        assert not ctx.get_prev_luma(s).any()
    ctx.close()


def test_host_step_256_streams_cut_by_streams(lv, orc, torch_cuda):
    """lv_step_host on more streams than one pipeline slot holds (the 256 x 720p config is cut by streams): the same
    results as the oracle, chunk after chunk, with the state carried per stream."""
    w, h, S, T = 352, 288, 40, 3
    os.environ["LV_HOST_CHUNK_MB"] = "1"
    os.environ["LV_HOST_SLOT_MB"] = "2"           # 13 CIF frames per slot: the 40 streams are cut into 13 + 13 + 13 + 1
    try:
        ctx = lv.Context(w, h, S, 20, 0)
        prev = np.zeros((S, w * h), np.uint8)
        for rep in range(2):
            frames = orc.gen_batch("camera", 6, S, T, w, h, first_frame=rep * T)
            ref = orc.convert_diff_batch(frames, prev, S, T, w, h, 20, 0, want_mask=False, n_threads=4)
            n = S * T
            bgr = np.zeros(n * ctx.bgr_bytes, np.uint8)
            cnt = np.zeros(n * ctx.mbs, np.uint16)
            summ = np.zeros(2 * n, np.uint32)
            ctx.step_host(np.ascontiguousarray(frames.reshape(-1)), S, T, bgr=bgr, mb_count=cnt, summary=summ)
            assert np.array_equal(bgr, ref["bgr"]) and np.array_equal(cnt, ref["mb_count"])
            assert np.array_equal(summ.reshape(-1, 2), ref["summary"])
        ctx.close()
    finally:
        del os.environ["LV_HOST_CHUNK_MB"], os.environ["LV_HOST_SLOT_MB"]


# ---- errors are codes ---------------------------------------------------------------------------------------------------
def test_misaligned_device_pointers_are_refused_not_fatal(lv, orc, torch_cuda):
    """A sub-buffer at an odd offset must give LV_E_ARG (-1), not a sticky cudaErrorMisalignedAddress; afterwards the
    context still works.  lv_write_bks (4-byte) / lv_convert_format (any alignment) stay bit-exact off the 16-byte grid."""
    torch = torch_cuda
    w, h = 64, 48
    ctx = lv.Context(w, h, 1, 20, 0)
    L, fb = lv.lib(), ctx.frame_bytes
    big = torch.zeros(4 * fb + 64, dtype=torch.uint8, device="cuda")
    out = ctx.alloc_outputs(1, mask_bytes=True)
    bgr_big = torch.zeros(ctx.bgr_bytes + 64, dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for off in (1, 4, 8):
        assert L.lv_convert_diff_batch(ctx._h, C.c_void_p(big.data_ptr() + off), 0, 1, 1, p(out["bgr"]), None, None, None,
                                       None, None, st) == -1
        assert L.lv_convert_diff_batch(ctx._h, p(big), 0, 1, 1, C.c_void_p(bgr_big.data_ptr() + off), None, None, None,
                                       None, None, st) == -1
        assert L.lv_convert_batch(ctx._h, C.c_void_p(big.data_ptr() + off), 1, p(out["bgr"]), st) == -1
        assert L.lv_diff_batch(ctx._h, C.c_void_p(big.data_ptr() + off), p(big), 1, None, None, None, None, None, st) == -1
    assert L.lv_convert_diff_batch(ctx._h, p(big), 0, 1, 1, None, None, C.c_void_p(out["mask_bytes"].data_ptr() + 2), None,
                                   None, None, st) == -1
    assert L.lv_convert_diff_batch(ctx._h, p(big), 0, 1, 1, None, None, None, None, None,
                                   C.c_void_p(out["summary"].data_ptr() + 2), st) == -1
    assert b"aligned" in L.lv_last_error()
    with pytest.raises(ValueError):
        ctx.convert_batch(big[1:], 1, out["bgr"])                      # the python mirror checks too
    torch.cuda.synchronize()                                           # no sticky error
    # ... and the context still works, bit-exact
    fr = orc.gen_batch("camera", 2, 1, 1, w, h)
    ref = orc.convert_diff_batch(fr, np.zeros((1, w * h), np.uint8), 1, 1, w, h, 20, 0)
    ctx.convert_diff_batch(torch.from_numpy(fr.reshape(-1)).cuda(), 1, 1, out)
    torch.cuda.synchronize()
    assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
    # entries that take narrower alignment: frames off the 16-byte grid
    cur = orc.gen_batch("camera", 2, 1, 2, w, h).reshape(-1)
    assert L.lv_write_bks(ctx._h, C.c_void_p(big.data_ptr() + 1), p(big), 1, 25, C.c_void_p(bgr_big.data_ptr()), None, st) == -1
    for off in (0, 4, 3):
        boff = off & ~3                                   # lv_write_bks: 4-byte grid
        d_cur = big[boff:boff + 2 * fb]
        d_cur.copy_(torch.from_numpy(cur))
        d_bkd = torch.from_numpy(cur[:fb].copy()).cuda()
        d_o = torch.zeros(2 * fb + 16, dtype=torch.uint8, device="cuda")
        ctx.write_bks(d_cur, d_bkd, 2, d_o[boff:boff + 2 * fb], tol=25)
        got = d_o[boff:boff + 2 * fb].cpu().numpy()
        want = np.concatenate([orc.write_bks(cur[i * fb:(i + 1) * fb], cur[:fb], w, h, 25) for i in range(2)])
        assert np.array_equal(got, want), off
        src = torch.zeros(3 * w * h + 16, dtype=torch.uint8, device="cuda")
        img = np.random.default_rng(off).integers(0, 256, 3 * w * h, dtype=np.uint8)
        src[off:off + 3 * w * h].copy_(torch.from_numpy(img))
        dst = torch.zeros(fb + 16, dtype=torch.uint8, device="cuda")
        ctx.convert_format(1, src[off:off + 3 * w * h], dst[off:off + fb], 1)
        assert np.array_equal(dst[off:off + fb].cpu().numpy(), orc.convert_format("rgb24_to_yv12", img, w, h)), off
    ctx.close()


def test_host_step_pitch_above_2gib(lv, orc, torch_cuda):
    """n_frames * 3*w*h > 2 GiB per stream (4K x 87 frames x 2 streams): the strided host copies exceed the device's
    cudaMemcpy2D pitch limit and must fall back, not fail half way.  Checked on sampled frames against the oracle."""
    torch = torch_cuda
    w, h, S, T = 3840, 2160, 2, 87
    px = w * h
    ctx = lv.Context(w, h, S, 20, 0)
    assert T * ctx.bgr_bytes > (1 << 31)
    n = S * T
    d_in = lv.gen_synthetic("camera", 8, 0, 0, S, T, w, h)
    h_in = d_in.cpu().numpy()
    del d_in
    bgr = np.zeros(n * ctx.bgr_bytes, np.uint8)
    cnt = np.zeros(n * ctx.mbs, np.uint16)
    summ = np.zeros(2 * n, np.uint32)
    os.environ["LV_HOST_CHUNK_MB"] = "64"       # 2 frames of each stream per chunk: 2D copies with the full host pitch
    try:
        ctx.step_host(h_in, S, T, bgr=bgr, mb_count=cnt, summary=summ)
    finally:
        del os.environ["LV_HOST_CHUNK_MB"]
    fb = ctx.frame_bytes
    for s, t in ((0, 0), (0, 1), (0, 86), (1, 0), (1, 43), (1, 86)):
        f = s * T + t
        cur = h_in[f * fb:(f + 1) * fb]
        assert np.array_equal(bgr[f * ctx.bgr_bytes:(f + 1) * ctx.bgr_bytes], orc.convert_frame(cur, w, h)), (s, t)
        refy = np.zeros(px, np.uint8) if t == 0 else h_in[(f - 1) * fb:(f - 1) * fb + px]
        _, c, _, total = orc.frame_diff_mb(cur[:px], refy, w, h, 20, 0)
        assert np.array_equal(cnt[f * ctx.mbs:(f + 1) * ctx.mbs], c) and summ[2 * f] == total, (s, t)
    ctx.close()


# ---- the peer exchange across processes (CUDA IPC) on ONE GPU ---------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ipc_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist

    import lvoracle as orc
    from livevideo_b200.shard import ShardedContext, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)                                  # both ranks on the SAME device
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, S, T, steps = 64, 48, 5, 3, 13
        sc = ShardedContext(w, h, S, rank=rank, world=world, device=0, exchange="p2p")
        first, count = shard_range(S, world, rank)
        out = sc.ctx.alloc_outputs(max(count, 1) * T)
        prev = np.zeros((S, w * h), np.uint8)
        ok = True
        for k in range(steps):
            frames = orc.gen_batch("camera", 30 + k, S, T, w, h)
            want = orc.convert_diff_batch(frames, prev, S, T, w, h, 20, 0, want_bgr=False, want_mask=False)["summary"]
            d_in = torch.from_numpy(frames[first:first + count].reshape(-1).copy()).cuda()
            if k == 4:
                sc.gate(hold=(rank == 0))                     # device-side rendezvous, mid-run (rank 0 holds its arrival)
            hd = sc.step_async(d_in, T, out)
            sc.gate_open()                                    # no-op unless a held gate is pending
            if k % 4 == 1:
                # device-side pick-up: the current stream waits for the table, the copy is enqueued behind the wait
                tab = hd.wait_table().clone()
                parts = [tab[r, :shard_range(S, world, r)[1]] for r in range(world)]
                got = torch.cat(parts, dim=0).cpu().numpy().view(np.uint32).reshape(S * T, 2)
                ok = ok and np.array_equal(got, want)
            if k % 4 == 3 or k == steps - 1:
                got = hd.result().cpu().numpy().view(np.uint32).reshape(S * T, 2)
                ok = ok and np.array_equal(got, want)
                hd.wait_summary()
                torch.cuda.synchronize()
                mine = out["summary"][:count * T * 2].cpu().numpy().view(np.uint32).reshape(-1, 2)
                ok = ok and np.array_equal(mine, want[first * T:(first + count) * T])
        assert sc.exchange is not None, sc.exchange_error
        err = sc.exchange.error()
        dist.barrier()
        q.put((rank, bool(ok), err))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_two_processes_one_gpu(torch_cuda):
    """One process per rank, CUDA IPC handles through torch.distributed (gloo), both ranks on cuda:0: the table every rank
    collects equals the single-rank oracle table; gate and wait_summary work across processes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 0), (1, True, 0)]


def test_exchange_dead_peer_gives_up_once(lv, torch_cuda):
    """A peer that never publishes: the collect gives up after ONE 4 s deadline for the whole kernel (not 4 s per word),
    reports error 2, and every later kernel of the exchange returns at once."""
    torch = torch_cuda
    from livevideo_b200.shard import PeerExchange
    xs = [PeerExchange(0, r, 2, 512, depth=4) for r in range(2)]
    PeerExchange.connect_local(xs)
    st = torch.cuda.Stream()
    blk = torch.arange(512, dtype=torch.int32, device="cuda")
    tab = torch.zeros(2 * 512, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xs[0].publish(0, blk, 512, st)
    xs[0].collect(0, tab, st)                                 # rank 1 never publishes step 0
    st.synchronize()
    dt = time.perf_counter() - t0
    assert xs[0].error() == 2
    assert 3.5 < dt < 6.0, dt                                 # one deadline, although 512 words were missing
    t0 = time.perf_counter()
    xs[0].publish(1, blk, 512, st)
    xs[0].collect(1, tab, st)
    st.synchronize()
    assert time.perf_counter() - t0 < 0.5                     # a broken exchange fails fast
    for x in xs:
        x.close()

# ---- randomized sweep ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LV_SWEEP_SEEDS", "12")))))
def test_randomized_geometry_batch_and_threshold_sweep(lv, orc, torch_cuda, seed, monkeypatch):
    """Seeded random cases through the C ABI against the oracle: width a multiple of 16 (16 ... 2304, so both macroblock
    mappings, rows longer and shorter than the early-load-order rule's 2048 pixels, partial last macroblock rows and ragged last
    x tiles), even height, 1 - 5 streams x 1 - 6 frames, tol in 0 ... 255, mb_thresh in 0 ... 256, content with many
    differences at exactly tol and tol + 1; two steps in a row (state carry), the second one through lv_step_host with a
    random chunk size; masks on and off."""
    torch = torch_cuda
    rng = np.random.default_rng(1000 + seed)
    w = 16 * int(rng.integers(1, 145))
    h = 2 * int(rng.integers(1, 1 + max(1, min(300, 600000 // w) // 2)))
    ns, nf = int(rng.integers(1, 6)), int(rng.integers(1, 7))
    tol = int(rng.choice([0, 1, 7, 20, 25, 127, 128, 200, 254, 255]))
    thr = int(rng.choice([0, 0, 1, 17, 128, 255, 256]))
    px = w * h
    monkeypatch.setenv("LV_HOST_CHUNK_MB", str(int(rng.choice([1, 2, 16]))))

    def make(prev_luma):
        fr = rng.integers(0, 256, (ns, nf, px * 3 // 2), dtype=np.uint8)
        # luma of frame t = reference +- {tol, tol + 1, 0, random}: the compare's boundary is hit on a quarter of all pixels
        ref_y = prev_luma
        for t in range(nf):
            kind = rng.integers(0, 4, (ns, px))
            sign = rng.integers(0, 2, (ns, px)) * 2 - 1
            d = np.where(kind == 0, tol, np.where(kind == 1, tol + 1, 0)) * sign
            y = np.clip(ref_y.astype(np.int32) + d, 0, 255).astype(np.uint8)
            y = np.where(kind == 3, fr[:, t, :px], y)
            fr[:, t, :px] = y
            ref_y = y
        return fr

    prev = np.zeros((ns, px), np.uint8)
    ctx = lv.Context(w, h, ns, tol, thr)
    masks = bool(seed & 1)
    # step 1: device entry
    frames = make(prev)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, thr, want_mask=True)       # advances prev in place
    got = run_fused(lv, torch, frames, ns, nf, w, h, tol, thr, ctx=ctx, masks=masks)
    assert_matches_oracle(orc, got, ref, w, h, ns * nf, masks=masks)
    # step 2: host entry on the same context (state carried on the device)
    frames = make(prev)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, thr)
    n = ns * nf
    bgr, cnt = np.zeros(n * ctx.bgr_bytes, np.uint8), np.zeros(n * ctx.mbs, np.uint16)
    mp, summ = np.zeros(n * ctx.mbs, np.uint8), np.zeros(2 * n, np.uint32)
    ctx.step_host(np.ascontiguousarray(frames.reshape(-1)), ns, nf, bgr=bgr, mb_count=cnt, mb_map=mp, summary=summ)
    assert np.array_equal(bgr, ref["bgr"]), (w, h, ns, nf)
    assert np.array_equal(cnt, ref["mb_count"]) and np.array_equal(mp, ref["mb_map"]), (w, h, ns, nf, tol, thr)
    assert np.array_equal(summ.reshape(-1, 2), ref["summary"])
    for s_ in range(ns):
        assert np.array_equal(ctx.get_prev_luma(s_), prev[s_])
    ctx.close()


# ---- LV_AUTO_PIN: a relinked application that never calls lv_host_register -----------------------------------------------
def test_auto_pin_registers_buffers_seen_twice(orc, torch_cuda, tmp_path):
    """LV_AUTO_PIN=1 (read once per process: the case runs in a child): plain malloc'ed planes handed to YV12_to_RGB24 /
    FrameDiff_MB are pageable on the first two calls and pinned + mapped from then on (zero-copy path); every call returns the
    oracle's bytes; a buffer seen with another size is not registered on that call; without the variable nothing is pinned."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import livevideo_b200 as lv, lvoracle as orc
w, h = 640, 360
px = w * h
rng = np.random.default_rng(3)
fr = rng.integers(0, 256, px * 3 // 2, dtype=np.uint8)
y, u, v = fr[:px].copy(), fr[px:px + px // 4].copy(), fr[px + px // 4:].copy()
dst = np.zeros(3 * px, np.uint8)
ref = np.roll(y, 5)
want = orc.convert_frame(fr, w, h)
wm, wc, wb, _ = orc.frame_diff_mb(y, ref, w, h, 20, 0)
csc = lv.ColorSpaceConversions()
states = []
mask, cnt, mp = np.zeros(px, np.uint8), np.zeros(((w + 15) // 16) * ((h + 15) // 16), np.uint16), np.zeros(((w + 15) // 16) * ((h + 15) // 16), np.uint8)
for call in range(4):
    dst[:] = 0
    csc.YV12_to_RGB24(y, u, v, dst, w, h)
    assert np.array_equal(dst, want), call
    lv.FrameDiff_MB(y, ref, w, h, 20, 0, mask=mask, mb_count=cnt, mb_map=mp)
    assert np.array_equal(mask, wm) and np.array_equal(cnt, wc) and np.array_equal(mp, wb), call
    states.append((lv.host_is_mapped(y), lv.host_is_mapped(dst), lv.host_is_mapped(ref), lv.host_is_mapped(mask)))
print("STATES", states)
'''
    import subprocess
    root = ROOT
    script = tmp_path / "autopin.py"
    script.write_text(code % (root, os.path.join(root, "oracle")))
    env = dict(os.environ, LV_AUTO_PIN="1")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    line = [l for l in r.stdout.splitlines() if l.startswith("STATES")][0]
    states = eval(line[len("STATES "):])
    # y goes through both entries, so it has been seen twice after the first round; the others after the second
    assert states[0] == (True, False, False, False), states
    assert all(states[1]) and all(states[3]), states
    env.pop("LV_AUTO_PIN")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    line = [l for l in r.stdout.splitlines() if l.startswith("STATES")][0]
    assert not any(any(s) for s in eval(line[len("STATES "):]))
