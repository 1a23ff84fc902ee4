"""Seeded random sweeps of every entry point beside the fused step, through the C ABI, against the oracle (tests/ only).

The fixed cases in test_gpu_parity.py / test_gpu_round2.py pin the geometries the reference's configurations use; these sweeps
walk the space between them: widths that are and are not multiples of 16 (vector and scalar kernels), one macroblock per row,
ragged last tiles, partial last macroblock rows, every tolerance, windows on and off the picture.  LV_SWEEP_SEEDS raises the
number of cases per sweep (default 8 each; `LV_SWEEP_SEEDS=200` is the soak run recorded in profiles/)."""
import mmap
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEEDS = list(range(int(os.environ.get("LV_SWEEP_SEEDS", "8"))))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _geom16(rng, max_px=400000):
    w = 16 * int(rng.integers(1, 145))
    h = 2 * int(rng.integers(1, 1 + max(1, min(300, max_px // w) // 2)))
    return w, h


def _geom4(rng, max_px=200000):
    """any geometry the legacy entries take: w % 4 == 0, h % 2 == 0"""
    w = 4 * int(rng.integers(1, 500))
    h = 2 * int(rng.integers(1, 1 + max(1, min(200, max_px // w) // 2)))
    return w, h


def _own(n, dtype=np.uint8):
    """page-aligned anonymous memory the test owns (so that it can be pinned in place)"""
    item = np.dtype(dtype).itemsize
    return np.frombuffer(mmap.mmap(-1, max(4096, (n * item + 4095) // 4096 * 4096)), dtype=dtype)[:n]


# ---- legacy drop-in entries: YV12_to_RGB24 / FrameDiff_MB on host pointers ---------------------------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_legacy_entries(lv, orc, torch_cuda, seed):
    rng = np.random.default_rng(2000 + seed)
    w, h = _geom4(rng) if seed % 3 else _geom16(rng, 200000)
    px = w * h
    tol, thr = int(rng.integers(0, 256)), int(rng.choice([0, 0, 1, 100, 255, 256]))
    fr = rng.integers(0, 256, px * 3 // 2, dtype=np.uint8)
    y, u, v, dst = _own(px), _own(px // 4), _own(px // 4), _own(3 * px)
    y[:], u[:], v[:] = fr[:px], fr[px:px + px // 4], fr[px + px // 4:]
    ref_y = np.clip(y.astype(int) + rng.choice([-tol - 1, -tol, 0, tol, tol + 1], px), 0, 255).astype(np.uint8)
    want = orc.convert_frame(fr, w, h)
    wm, wc, wb, _ = orc.frame_diff_mb(y, ref_y, w, h, tol, thr)
    csc = lv.ColorSpaceConversions()
    registered = []
    before = lv.get_zero_copy()
    try:
        for mode in ("pageable", "off", "plain", "staged"):
            if mode == "off":                      # from here on the buffers are pinned + mapped in place
                ry = _own(px)
                ry[:] = ref_y
                ref_y = ry
                for a in (y, u, v, dst, ref_y):
                    lv.host_register(a)
                    registered.append(a)
            if mode != "pageable":
                lv.set_zero_copy(mode)
            dst[:] = 0
            csc.YV12_to_RGB24(y, u, v, dst, w, h)
            assert np.array_equal(dst, want), (mode, w, h)
            mask, cnt, mp = lv.FrameDiff_MB(y, ref_y, w, h, tol, thr)
            assert np.array_equal(mask, wm) and np.array_equal(cnt, wc) and np.array_equal(mp, wb), (mode, w, h, tol, thr)
    finally:
        lv.set_zero_copy(before)
        for a in registered:
            lv.host_unregister(a)


# ---- conversion only / difference only over batches (K2, K3), both kernel families ------------------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_convert_batch_and_diff_batch(lv, orc, torch_cuda, seed):
    torch = torch_cuda
    rng = np.random.default_rng(3000 + seed)
    w, h = _geom16(rng)
    n = int(rng.integers(1, 6))
    px = w * h
    tol, thr = int(rng.integers(0, 256)), int(rng.choice([0, 0, 3, 128, 256]))
    frames = rng.integers(0, 256, (n, px * 3 // 2), dtype=np.uint8)
    cur = np.ascontiguousarray(frames[:, :px])
    refy = np.clip(cur.astype(int) + rng.choice([-tol - 1, -tol, 0, 0, tol, tol + 1], (n, px)), 0, 255).astype(np.uint8)
    lv.set_kernel_variant("tma" if seed % 4 == 3 else "tile")
    try:
        ctx = lv.Context(w, h, 1, tol, thr)
        d_out = torch.empty(n * ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
        ctx.convert_batch(torch.from_numpy(frames.reshape(-1)).cuda(), n, d_out)
        got = d_out.cpu().numpy().reshape(n, -1)
        for i in range(n):
            assert np.array_equal(got[i], orc.convert_frame(frames[i], w, h)), (w, h, i)
        masks = bool(seed & 1)
        out = ctx.alloc_outputs(n, bgr=False, mask_bits=masks, mask_bytes=masks)
        ctx.diff_batch(torch.from_numpy(cur.reshape(-1)).cuda(), torch.from_numpy(refy.reshape(-1)).cuda(), n, out)
        torch.cuda.synchronize()
        cnt, mp = out["mb_count"].cpu().numpy().view(np.uint16), out["mb_map"].cpu().numpy()
        summ = out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2)
        for i in range(n):
            m, c, b, total = orc.frame_diff_mb(cur[i], refy[i], w, h, tol, thr)
            assert np.array_equal(cnt[i * ctx.mbs:(i + 1) * ctx.mbs], c), (w, h, tol, i)
            assert np.array_equal(mp[i * ctx.mbs:(i + 1) * ctx.mbs], b), (w, h, tol, thr, i)
            assert summ[i, 0] == total and summ[i, 1] == int(b.sum())
            if masks:
                assert np.array_equal(out["mask_bytes"].cpu().numpy()[i * px:(i + 1) * px], m)
                assert np.array_equal(out["mask_bits"].cpu().numpy().view(np.uint16)[i * ctx.mask_units:(i + 1) * ctx.mask_units],
                                      orc.pack_mask16(m, w, h))
        ctx.close()
    finally:
        lv.set_kernel_variant("tile")


# ---- the box from the fused kernel, random windows (on, across and off the picture's edges) --------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_fused_box(lv, orc, torch_cuda, seed):
    torch = torch_cuda
    rng = np.random.default_rng(4000 + seed)
    w, h = _geom16(rng, 250000)
    ns, nf, k = int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 17))
    n, px = ns * nf, w * h
    tol = int(rng.choice([0, 5, 20, 200]))
    frames = rng.integers(0, 256, (ns, nf, px * 3 // 2), dtype=np.uint8)
    # sparse changes: most pixels repeat the previous frame, so the boxes are not always the whole window
    for t in range(1, nf):
        keep = rng.random((ns, px)) < 0.995
        frames[:, t, :px] = np.where(keep, frames[:, t - 1, :px], frames[:, t, :px])
    prev = frames[:, 0, :px].copy()
    prev[:, :: max(1, px // 7)] ^= 0x80                                   # a handful of changes in every stream's first frame
    state = prev.copy()
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, 0, want_mask=True)
    wins = np.zeros((n, k, 4), np.int32)
    for f in range(n):
        for j in range(k):
            xa, xb = sorted(int(v) for v in rng.integers(0, w, 2))
            ya, yb = sorted(int(v) for v in rng.integers(0, h, 2))
            wins[f, j] = (xa, xb, ya, yb)
    wins[0, 0] = (0, w - 1, 0, h - 1)
    want = np.zeros((n, k, 4), np.int32)
    for f in range(n):
        for j in range(k):
            want[f, j] = orc.object_box(ref["mask"][f * px:(f + 1) * px], w, *[int(v) for v in wins[f, j]])
    for windowed in (False, True):
        ctx = lv.Context(w, h, ns, tol, 0)
        for s in range(ns):
            ctx.set_prev_luma(s, state[s])
        out = ctx.alloc_outputs(n)
        boxes = ctx.convert_diff_box_batch(torch.from_numpy(frames.reshape(-1)).cuda(), ns, nf, out, torch.from_numpy(wins).cuda(),
                                           windowed=windowed)
        torch.cuda.synchronize()
        assert np.array_equal(boxes.cpu().numpy(), want), (w, h, ns, nf, k, windowed)
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
        assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"])
        assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"])
        for s in range(ns):
            assert np.array_equal(ctx.get_prev_luma(s), prev[s]), (windowed, s)
        ctx.close()


# ---- decoder pictures (16-bit planes, crop) into the fused kernel and through the narrowing kernel ---------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_pictures(lv, orc, torch_cuda, seed):
    torch = torch_cuda
    rng = np.random.default_rng(5000 + seed)
    w, h = _geom16(rng, 150000)
    cl, ct = 2 * int(rng.integers(0, 20)), 2 * int(rng.integers(0, 12))
    if seed % 2 == 0:
        cl = cl // 16 * 16                          # the one-kernel form takes crops that keep 16-byte alignment
    sx = w + cl + 2 * int(rng.integers(0, 9))
    if seed % 4 != 1:
        sx = (sx + 15) // 16 * 16                   # pitches that are not multiples of 16 go through the scratch
    sy = (h + ct + 2 * int(rng.integers(0, 9)) + 1) // 2 * 2
    ns, nf = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    n, px = ns * nf, w * h
    samples = sx * sy * 3 // 2
    pics = rng.integers(0, 65536, (ns, nf, samples), dtype=np.uint16)
    frames = np.stack([[orc.picture16_to_i420(pics[s, t], sx, sy, cl, sx - cl - w, ct, sy - ct - h) for t in range(nf)]
                       for s in range(ns)])
    prev = np.zeros((ns, px), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
    ctx = lv.Context(w, h, ns)
    d_pic = torch.from_numpy(np.ascontiguousarray(pics).view(np.int16).reshape(-1)).cuda()
    out = ctx.alloc_outputs(n)
    ctx.convert_diff_pictures(d_pic, sx, sy, cl, ct, ns, nf, out)
    torch.cuda.synchronize()
    assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"]), (sx, sy, cl, ct, w, h)
    assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"])
    assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"])
    for s in range(ns):
        assert np.array_equal(ctx.get_prev_luma(s), prev[s])
    d_i420 = torch.empty(n * ctx.frame_bytes, dtype=torch.uint8, device="cuda")
    ctx.narrow_pictures(d_pic, n, sx, sy, cl, ct, d_i420)
    assert np.array_equal(d_i420.cpu().numpy(), frames.reshape(-1)), (sx, sy, cl, ct, w, h)
    # and from host memory: lv_step_host_pictures on a fresh context
    ctx2 = lv.Context(w, h, ns)
    bgr, cnt = np.zeros(n * ctx.bgr_bytes, np.uint8), np.zeros(n * ctx.mbs, np.uint16)
    ctx2.step_host_pictures(np.ascontiguousarray(pics).reshape(-1), sx, sy, cl, ct, ns, nf, bgr=bgr, mb_count=cnt)
    assert np.array_equal(bgr, ref["bgr"]) and np.array_equal(cnt, ref["mb_count"])
    ctx2.close()
    ctx.close()


# ---- WriteBks and the stand-alone object box ----------------------------------------------------------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_write_bks_and_object_box(lv, orc, torch_cuda, seed):
    torch = torch_cuda
    rng = np.random.default_rng(6000 + seed)
    w, h = _geom4(rng, 150000) if seed % 2 else _geom16(rng, 150000)
    n = int(rng.integers(1, 5))
    tol = int(rng.choice([0, 1, 20, 25, 128, 255]))
    fb = w * h * 3 // 2
    bkd = rng.integers(0, 256, fb, dtype=np.uint8)
    cur = np.clip(bkd.astype(int)[None] + rng.choice([-tol - 1, -tol, 0, 0, 0, tol, tol + 1], (n, fb)), 0, 255).astype(np.uint8)
    ctx = lv.Context(w, h)
    d_out = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    bits = torch.zeros(n * ctx.mask_units, dtype=torch.int16, device="cuda") if w % 16 == 0 else None
    ctx.write_bks(torch.from_numpy(cur.reshape(-1)).cuda(), torch.from_numpy(bkd).cuda(), n, d_out, tol=tol, obj_bits=bits)
    got = d_out.cpu().numpy().reshape(n, fb)
    for i in range(n):
        assert np.array_equal(got[i], orc.write_bks(cur[i], bkd, w, h, tol)), (w, h, tol, i)
    masks = np.stack([(rng.random(w * h) < d).astype(np.uint8) for d in rng.choice([0.0, 0.0003, 0.02, 0.6], n)])
    d_bytes = torch.from_numpy(masks.reshape(-1)).cuda()
    for _ in range(4):
        xa, xb = sorted(int(v) for v in rng.integers(0, w, 2))
        ya, yb = sorted(int(v) for v in rng.integers(0, h, 2))
        want = np.array([orc.object_box(masks[i], w, xa, xb, ya, yb) for i in range(n)], np.int32)
        assert np.array_equal(ctx.object_box_batch(d_bytes, n, xa, xb, ya, yb).cpu().numpy(), want), (w, h, xa, xb, ya, yb)
        if w % 16 == 0:
            mb = np.concatenate([orc.pack_mask16(masks[i], w, h) for i in range(n)])
            got = ctx.object_box_batch(torch.from_numpy(mb.view(np.int16)).cuda(), n, xa, xb, ya, yb, mask_is_bits=True)
            assert np.array_equal(got.cpu().numpy(), want), ("bits", w, h, xa, xb, ya, yb)
    ctx.close()


# ---- the five other Convert.h converters --------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_other_converters(lv, orc, torch_cuda, seed):
    torch = torch_cuda
    rng = np.random.default_rng(7000 + seed)
    names = ["rgb24_to_yv12", "yvu9_to_yv12", "yuy2_to_yv12", "yv12_to_yvu9", "yv12_to_yuy2"]
    w = 16 * int(rng.integers(1, 130)) if seed % 2 else 8 * int(rng.integers(1, 260))
    h = 8 * int(rng.integers(1, 1 + max(1, min(40, 200000 // w // 8))))
    n = int(rng.integers(1, 4))
    for kind, name in enumerate(names, start=1):
        mode, inb, outb = orc.FORMATS[name]
        data = rng.integers(0, 256, (n, w * h // 8 * inb), dtype=np.uint8)
        want = np.stack([orc.convert_format(name, data[i], w, h) for i in range(n)])
        ctx = lv.Context(w, h)
        d_out = torch.zeros(n * w * h // 8 * outb, dtype=torch.uint8, device="cuda")
        ctx.convert_format(kind, torch.from_numpy(data.reshape(-1)).cuda(), d_out, n)
        assert np.array_equal(d_out.cpu().numpy(), want.reshape(-1)), (name, w, h, n)
        ctx.close()


# ---- several pictures per call on host pointers (DisplayPic's pattern) ------------------------------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_multi_picture_entries(lv, orc, torch_cuda, seed):
    rng = np.random.default_rng(8000 + seed)
    w, h = _geom16(rng, 120000) if seed % 2 else _geom4(rng, 120000)
    px = w * h
    n = int(rng.integers(1, 12))
    tol = int(rng.integers(0, 256))
    frames = rng.integers(0, 256, (n, px * 3 // 2), dtype=np.uint8)
    ys, us, vs, ds = [], [], [], []
    for i in range(n):
        y, u, v, d = _own(px), _own(px // 4), _own(px // 4), _own(3 * px)
        y[:], u[:], v[:] = frames[i, :px], frames[i, px:px + px // 4], frames[i, px + px // 4:]
        ys.append(y), us.append(u), vs.append(v), ds.append(d)
    pinned = seed % 3 != 0
    if pinned:
        for a in ys + us + vs + ds:
            lv.host_register(a)
    try:
        lv.ColorSpaceConversions().YV12_to_RGB24_multi(ys, us, vs, ds, w, h)
        for i in range(n):
            assert np.array_equal(ds[i], orc.convert_frame(frames[i], w, h)), (w, h, n, i, pinned)
        if n > 1:
            # picture i against picture i - 1 (the chained form), all outputs
            mask, cnt, mp = lv.FrameDiff_MB_multi(ys[1:], ys[:-1], w, h, tol, 0)
            for i in range(n - 1):
                wm, wc, wb, _ = orc.frame_diff_mb(ys[i + 1], ys[i], w, h, tol, 0)
                assert np.array_equal(mask[i], wm) and np.array_equal(cnt[i], wc) and np.array_equal(mp[i], wb), (w, h, n, i)
    finally:
        if pinned:
            for a in ys + us + vs + ds:
                lv.host_unregister(a)


# ---- a random SESSION on one context: the per-stream state under every call that touches it ----------------------------
@pytest.mark.parametrize("seed", SEEDS)
def test_random_stateful_session(lv, orc, torch_cuda, seed):
    """A context with several streams driven by a random sequence of calls -- device steps and host steps over random stream
    ranges and frame counts, decoder-picture steps, tol / mb_thresh changes, the static-background mode switched on and off,
    set / reset of a stream's previous luma -- against a host model that keeps one previous-luma plane per stream
    (lib/JM/image.c:1553-1561; static mode: MainDlg.cpp:238-246).  After every call all outputs and every stream's state must
    equal the model's."""
    torch = torch_cuda
    rng = np.random.default_rng(9000 + seed)
    w, h = _geom16(rng, 60000)
    S = int(rng.integers(1, 7))
    px = w * h
    tol, thr = 20, 0
    ctx = lv.Context(w, h, S, tol, thr)
    prev = np.zeros((S, px), np.uint8)
    static = False

    def model(frames, first, cnt, T):
        """advance the model over streams [first, first + cnt) and return the expected outputs"""
        if not static:
            sub = prev[first:first + cnt].copy()
            ref = orc.convert_diff_batch(frames, sub, cnt, T, w, h, tol, thr)
            prev[first:first + cnt] = sub
            return ref["bgr"], ref["mb_count"], ref["mb_map"], ref["summary"]
        bgr, cnts, maps, summ = [], [], [], []
        for s in range(cnt):
            for t in range(T):
                _, c, m, changed = orc.frame_diff_mb(frames[s, t, :px], prev[first + s], w, h, tol, thr)
                bgr.append(orc.convert_frame(frames[s, t], w, h)), cnts.append(c), maps.append(m)
                summ.append((changed, int(m.sum())))
        return np.concatenate(bgr), np.concatenate(cnts), np.concatenate(maps), np.array(summ, np.uint32)

    for step in range(int(rng.integers(6, 14))):
        op = int(rng.integers(0, 10))
        if op == 0:
            tol, thr = int(rng.integers(0, 256)), int(rng.choice([0, 0, 2, 256]))
            ctx.set_params(tol, thr)
            continue
        if op == 1:
            static = not static
            ctx.set_ref_mode(static)
            continue
        if op == 2:
            s = int(rng.integers(0, S))
            if rng.integers(0, 2):
                y = rng.integers(0, 256, px, dtype=np.uint8)
                ctx.set_prev_luma(s, y)
                prev[s] = y
            else:
                ctx.reset_prev_luma()
                prev[:] = 0
            continue
        first = int(rng.integers(0, S))
        cnt = int(rng.integers(1, S - first + 1))
        T = int(rng.integers(1, 5))
        n = cnt * T
        frames = rng.integers(0, 256, (cnt, T, px * 3 // 2), dtype=np.uint8)
        # luma close to the model's reference so that differences straddle the tolerance
        base = prev[first:first + cnt]
        for t in range(T):
            d = rng.choice([-tol - 1, -tol, 0, 0, tol, tol + 1], (cnt, px))
            y = np.clip(base.astype(int) + d, 0, 255).astype(np.uint8)
            frames[:, t, :px] = np.where(rng.random((cnt, px)) < 0.9, y, frames[:, t, :px])
            if not static:
                base = frames[:, t, :px]
        if op <= 5:                                  # device step
            want = model(frames, first, cnt, T)
            out = ctx.alloc_outputs(n)
            ctx.convert_diff_batch(torch.from_numpy(frames.reshape(-1)).cuda(), cnt, T, out, first_stream=first)
            torch.cuda.synchronize()
            got = (out["bgr"].cpu().numpy(), out["mb_count"].cpu().numpy().view(np.uint16), out["mb_map"].cpu().numpy(),
                   out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2))
        elif op <= 7:                                # host step
            want = model(frames, first, cnt, T)
            bgr, c16 = np.zeros(n * ctx.bgr_bytes, np.uint8), np.zeros(n * ctx.mbs, np.uint16)
            mp, summ = np.zeros(n * ctx.mbs, np.uint8), np.zeros(2 * n, np.uint32)
            ctx.step_host(np.ascontiguousarray(frames.reshape(-1)), cnt, T, bgr=bgr, mb_count=c16, mb_map=mp, summary=summ,
                          first_stream=first)
            got = (bgr, c16, mp, summ.reshape(-1, 2))
        else:                                        # decoder pictures (upper bytes random: they are dropped, not clipped)
            want = model(frames, first, cnt, T)
            pics = frames.astype(np.uint16) | (rng.integers(0, 256, frames.shape, dtype=np.uint16) << 8)
            out = ctx.alloc_outputs(n)
            ctx.convert_diff_pictures(torch.from_numpy(np.ascontiguousarray(pics).view(np.int16).reshape(-1)).cuda(), w, h, 0, 0,
                                      cnt, T, out, first_stream=first)
            torch.cuda.synchronize()
            got = (out["bgr"].cpu().numpy(), out["mb_count"].cpu().numpy().view(np.uint16), out["mb_map"].cpu().numpy(),
                   out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2))
        for name, g, e in zip(("bgr", "mb_count", "mb_map", "summary"), got, want):
            assert np.array_equal(g, e), (seed, step, op, name, w, h, first, cnt, T, static, tol, thr)
        for s in range(S):
            assert np.array_equal(ctx.get_prev_luma(s), prev[s]), (seed, step, op, "state", s, static)
    ctx.close()
