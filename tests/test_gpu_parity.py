"""Parity of the CUDA path against the oracle, through the C ABI (run on the B200 box: -m gpu).

Bar: bit-exact for every output -- BGR bytes, difference masks (bytes and bits), per-macroblock counts, the
block map, the per-frame summaries and the carried previous-luma state."""
import ctypes as C
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(autouse=True)
def _device_unchanged_by_the_library():
    """No entry of the C ABI may leave a different CUDA device current than the one it found."""
    import torch
    if not torch.cuda.is_available():
        yield
        return
    torch.cuda.set_device(0)
    yield
    assert torch.cuda.current_device() == 0, "a library call left another CUDA device current"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_fused(lv, torch, frames, ns, nf, w, h, tol, thr, ctx=None, masks=True):
    own = ctx is None
    ctx = ctx or lv.Context(w, h, ns, tol, thr)
    d_in = torch.from_numpy(np.ascontiguousarray(frames).reshape(-1)).cuda()
    out = ctx.alloc_outputs(ns * nf, mask_bits=masks, mask_bytes=masks)
    ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    res = {k: (None if v is None else v.cpu().numpy()) for k, v in out.items()}
    res["mb_count"] = res["mb_count"].view(np.uint16)
    res["summary"] = res["summary"].view(np.uint32).reshape(-1, 2)
    if masks:
        res["mask_bits"] = res["mask_bits"].view(np.uint16)
    if own:
        res["state"] = np.stack([ctx.get_prev_luma(s) for s in range(ns)])
        ctx.close()
    return res


def assert_matches_oracle(orc, got, ref, w, h, n, masks=True):
    assert np.array_equal(got["bgr"], ref["bgr"]), "BGR bytes differ"
    assert np.array_equal(got["mb_count"], ref["mb_count"]), "macroblock counts differ"
    assert np.array_equal(got["mb_map"], ref["mb_map"]), "block map differs"
    assert np.array_equal(got["summary"], ref["summary"]), "summaries differ"
    if masks:
        assert np.array_equal(got["mask_bytes"], ref["mask"]), "byte mask differs"
        px = w * h
        bits = np.concatenate([orc.pack_mask16(ref["mask"][i * px:(i + 1) * px], w, h) for i in range(n)])
        assert np.array_equal(got["mask_bits"], bits), "bit mask differs"


# ---- conversion -------------------------------------------------------------------------------------------

def test_convert_exhaustive_2_24(lv, orc, torch_cuda):
    """Every (Y,U,V) triple: 256 frames of 512x512 (Y constant per frame, U = column, V = row)."""
    torch = torch_cuda
    u = np.repeat(np.arange(256, dtype=np.uint8)[None, :], 256, 0).ravel()
    v = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1).ravel()
    ctx = lv.Context(512, 512)
    bad = 0
    for y0 in range(0, 256, 64):
        frames = np.stack([np.concatenate([np.full(512 * 512, y, np.uint8), u, v]) for y in range(y0, y0 + 64)])
        d_in = torch.from_numpy(frames.reshape(-1)).cuda()
        d_out = torch.empty(64 * ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
        ctx.convert_batch(d_in, 64, d_out)
        got = d_out.cpu().numpy().reshape(64, -1)
        for i in range(64):
            bad += int((got[i] != orc.convert_frame(frames[i], 512, 512)).sum())
    assert bad == 0


def test_convert_golden_digests_device_and_legacy_entry(lv, orc, golden, torch_cuda):
    """Digests recorded from the reference's original object code (tests/golden + SURVEY.md 8(c))."""
    torch = torch_cuda
    csc = lv.ColorSpaceConversions()
    for e in golden["convert"]:
        w, h = e["w"], e["h"]
        fr = orc.gen(e["kind"], e["seed"], e["stream"], e["frame"], w, h)
        px = w * h
        dst = np.zeros(3 * px, dtype=np.uint8)
        csc.YV12_to_RGB24(fr[:px].copy(), fr[px:px + px // 4].copy(), fr[px + px // 4:].copy(), dst, w, h)
        assert dst[:6].tolist() == e["first6"], e
        assert sha(dst) == e["sha256"], ("legacy entry", e)
        if w % 16 == 0:
            ctx = lv.Context(w, h)
            d_out = torch.empty(ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
            ctx.convert_batch(torch.from_numpy(fr).cuda(), 1, d_out)
            assert sha(d_out.cpu().numpy()) == e["sha256"], ("device entry", e)
            ctx.close()


def test_legacy_kats_and_layout(lv, golden, torch_cuda):
    csc = lv.ColorSpaceConversions()
    for k in golden["kat"]:
        y, u, v = k["yuv"]
        dst = np.zeros(24, np.uint8)
        csc.YV12_to_RGB24(np.full(8, y, np.uint8), np.full(2, u, np.uint8), np.full(2, v, np.uint8), dst, 4, 2)
        assert dst.reshape(8, 3).tolist() == [k["bgr"]] * 8
    y = np.repeat(np.array([16, 80, 160, 235], dtype=np.uint8), 4)
    dst = np.zeros(48, np.uint8)
    csc.YV12_to_RGB24(y, np.full(4, 128, np.uint8), np.full(4, 128, np.uint8), dst, 4, 4)
    assert [int(dst[12 * r]) for r in range(4)] == golden["layout_kat"]["first_byte_of_dst_rows"]   # bottom-up
    dst = np.zeros(24, np.uint8)
    csc.YV12_to_RGB24(np.full(8, 128, np.uint8), np.array([0, 255], np.uint8), np.full(2, 128, np.uint8), dst, 4, 2)
    assert [int(dst[3 * p]) for p in range(4)] == golden["chroma_kat"]["b_row0"]                     # 2x2 replication
    assert [int(dst[12 + 3 * p]) for p in range(4)] == golden["chroma_kat"]["b_row1"]
    with pytest.raises(ValueError):
        csc.YV12_to_RGB24(np.zeros(6, np.uint8), np.zeros(2, np.uint8), np.zeros(2, np.uint8), np.zeros(18, np.uint8), 3, 2)
    with pytest.raises(ValueError):
        csc.RGB24_to_YV12(np.zeros(18, np.uint8), np.zeros(9, np.uint8), 3, 2)


@pytest.mark.parametrize("w,h", [(4, 2), (8, 2), (12, 10), (20, 6), (36, 4), (100, 50), (16, 2), (32, 18), (528, 34)])
def test_legacy_convert_odd_geometries(lv, orc, torch_cuda, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    px = w * h
    fr = rng.integers(0, 256, px * 3 // 2, dtype=np.uint8)
    dst = np.zeros(3 * px, np.uint8)
    lv.ColorSpaceConversions().YV12_to_RGB24(fr[:px].copy(), fr[px:px + px // 4].copy(), fr[px + px // 4:].copy(),
                                             dst, w, h)
    assert np.array_equal(dst, orc.convert_frame(fr, w, h))


# ---- fused step ----------------------------------------------------------------------------------------------

def test_fused_golden_entries(lv, orc, golden, torch_cuda):
    for e in golden["diff"]:
        w, h, ns, nf = e["w"], e["h"], e["n_streams"], e["n_frames"]
        frames = orc.gen_batch(e["kind"], e["seed"], ns, nf, w, h)
        got = run_fused(lv, torch_cuda, frames, ns, nf, w, h, e["tol"], e["mb_thresh"])
        assert sha(got["bgr"]) == e["bgr"]["sha256"], e
        assert sha(got["mask_bytes"]) == e["mask"]["sha256"], e
        assert sha(got["mb_count"]) == e["mb_count"]["sha256"], e
        assert sha(got["mb_map"]) == e["mb_map"]["sha256"], e
        assert got["summary"].tolist() == e["summary"], e
        assert sha(got["state"]) == e["state"]["sha256"], e


@pytest.mark.parametrize("tol", [0, 1, 20, 25, 254, 255])
@pytest.mark.parametrize("thr", [0, 1, 128, 255, 256])
def test_fused_threshold_sweep(lv, orc, torch_cuda, tol, thr):
    w, h, ns, nf = 176, 72, 2, 3          # 72 rows: partial last macroblock row (8 lines), like 1080
    rng = np.random.default_rng(tol * 1000 + thr)
    frames = rng.integers(0, 256, (ns, nf, w * h * 3 // 2), dtype=np.uint8)
    # force many |d| in {tol, tol+1} and whole macroblocks at exactly 0 / 256 changed pixels
    base = frames[:, 0, :w * h].astype(int)
    frames[:, 1, :w * h:2] = np.clip(base[:, ::2] + tol, 0, 255)
    frames[:, 1, 1:w * h:2] = np.clip(base[:, 1::2] - tol - 1, 0, 255)
    frames[:, 2, :16 * w] = frames[:, 1, :16 * w]                 # first MB row unchanged
    prev = np.zeros((ns, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, thr)
    got = run_fused(lv, torch_cuda, frames, ns, nf, w, h, tol, thr)
    assert_matches_oracle(orc, got, ref, w, h, ns * nf)
    assert np.array_equal(got["state"], prev)


def test_fused_ragged_widths_and_tiny_frames(lv, orc, torch_cuda):
    # widths that do not fill a 512-pixel block tile, heights with 1..8 row pairs in the last MB row
    for (w, h) in [(16, 2), (16, 16), (48, 18), (528, 20), (1040, 30), (4112, 2), (64, 18), (96, 20), (160, 8), (32, 2),
                   (352, 30), (4096, 24)]:
        frames = orc.gen_batch("camera" if h >= 16 and w >= 16 and h // 8 > 0 and w // 8 > 0 and w > w // 8 else "uniform",
                               3, 2, 2, w, h) if h >= 16 else orc.gen_batch("uniform", 3, 2, 2, w, h)
        prev = np.zeros((2, w * h), np.uint8)
        ref = orc.convert_diff_batch(frames, prev, 2, 2, w, h, 20, 0)
        got = run_fused(lv, torch_cuda, frames, 2, 2, w, h, 20, 0)
        assert_matches_oracle(orc, got, ref, w, h, 4)


def test_state_carry_split_batches_and_stream_ranges(lv, orc, torch_cuda):
    torch = torch_cuda
    w, h, ns, nf = 352, 288, 3, 6
    frames = orc.gen_batch("camera", 7, ns, nf, w, h)
    prev = np.zeros((ns, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
    # (a) two launches of 3 frames == one launch of 6 (the state carries frame 2's luma)
    ctx = lv.Context(w, h, ns, 20, 0)
    a = run_fused(lv, torch, frames[:, :3], ns, 3, w, h, 20, 0, ctx=ctx)
    b = run_fused(lv, torch, frames[:, 3:], ns, 3, w, h, 20, 0, ctx=ctx)
    cnt = np.concatenate([a["mb_count"].reshape(ns, 3, -1), b["mb_count"].reshape(ns, 3, -1)], axis=1)
    assert np.array_equal(cnt.reshape(-1), ref["mb_count"])
    bgr = np.concatenate([a["bgr"].reshape(ns, 3, -1), b["bgr"].reshape(ns, 3, -1)], axis=1)
    assert np.array_equal(bgr.reshape(-1), ref["bgr"])
    for s in range(ns):
        assert np.array_equal(ctx.get_prev_luma(s), prev[s])
    ctx.close()
    # (b) stepping only stream 1 leaves the state of streams 0 and 2 alone
    ctx = lv.Context(w, h, ns, 20, 0)
    for s in range(ns):
        ctx.set_prev_luma(s, frames[s, 0, :w * h].copy())
    out = ctx.alloc_outputs(2)
    ctx.convert_diff_batch(torch.from_numpy(frames[1, 1:3].reshape(-1).copy()).cuda(), 1, 2, out, first_stream=1)
    torch.cuda.synchronize()
    assert np.array_equal(ctx.get_prev_luma(0), frames[0, 0, :w * h])
    assert np.array_equal(ctx.get_prev_luma(2), frames[2, 0, :w * h])
    assert np.array_equal(ctx.get_prev_luma(1), frames[1, 2, :w * h])
    p1 = frames[1, 0, :w * h].copy().reshape(1, -1)
    r1 = orc.convert_diff_batch(frames[1, 1:3], p1, 1, 2, w, h, 20, 0)
    assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), r1["mb_count"])
    # (c) static background as the reference (MainDlg.cpp:238-246): reset + set
    ctx.reset_prev_luma()
    torch.cuda.synchronize()
    assert not ctx.get_prev_luma(1).any()
    ctx.close()


def test_identical_frames_and_idempotence(lv, orc, torch_cuda):
    w, h = 352, 288
    fr = orc.gen("camera", 2, 0, 0, w, h)
    frames = np.stack([fr, fr, fr])[None]
    got = run_fused(lv, torch_cuda, frames, 1, 3, w, h, 0, 0)
    assert got["summary"][1:].tolist() == [[0, 0], [0, 0]]          # tol = 0: identical frames never change
    assert not got["mask_bytes"][w * h:].any()
    again = run_fused(lv, torch_cuda, frames, 1, 3, w, h, 0, 0)
    for k in ("bgr", "mask_bytes", "mask_bits", "mb_count", "mb_map", "summary"):
        assert np.array_equal(got[k], again[k])


def test_diff_only_entry_and_legacy_framediff(lv, orc, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(5)
    w, h, n = 352, 288, 3
    cur = rng.integers(0, 256, (n, w * h), dtype=np.uint8)
    ref = np.clip(cur.astype(int) + rng.integers(-30, 31, (n, w * h)), 0, 255).astype(np.uint8)
    ctx = lv.Context(w, h, 1, 20, 3)
    out = ctx.alloc_outputs(n, bgr=False, mask_bits=True, mask_bytes=True)
    ctx.diff_batch(torch.from_numpy(cur.reshape(-1)).cuda(), torch.from_numpy(ref.reshape(-1)).cuda(), n, out)
    torch.cuda.synchronize()
    for i in range(n):
        mask, cnt, mp, total = orc.frame_diff_mb(cur[i], ref[i], w, h, 20, 3)
        assert np.array_equal(out["mask_bytes"].cpu().numpy()[i * w * h:(i + 1) * w * h], mask)
        assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16)[i * ctx.mbs:(i + 1) * ctx.mbs], cnt)
        assert np.array_equal(out["mb_map"].cpu().numpy()[i * ctx.mbs:(i + 1) * ctx.mbs], mp)
        assert out["summary"].cpu().numpy().view(np.uint32)[2 * i] == total
    ctx.close()
    # the host-pointer legacy entry, incl. sizes that are not multiples of 16 and NULL outputs
    for (w, h, tol, thr) in [(352, 288, 20, 0), (100, 50, 25, 10), (20, 6, 0, 0), (1920, 1080, 20, 128), (4, 2, 255, 0)]:
        cur = rng.integers(0, 256, w * h, dtype=np.uint8)
        ref = np.clip(cur.astype(int) + rng.integers(-40, 41, w * h), 0, 255).astype(np.uint8)
        mask, cnt, mp = lv.FrameDiff_MB(cur, ref, w, h, tol, thr)
        emask, ecnt, emp, _ = orc.frame_diff_mb(cur, ref, w, h, tol, thr)
        assert np.array_equal(mask, emask) and np.array_equal(cnt, ecnt) and np.array_equal(mp, emp)
    L = lv.lib()
    cnt = np.zeros(396, np.uint16)
    cur = rng.integers(0, 256, 352 * 288, dtype=np.uint8)
    L.FrameDiff_MB(cur.ctypes.data, cur.ctypes.data, 352, 288, 20, 0, None, cnt.ctypes.data, None)   # NULL mask/map
    assert not cnt.any()


def test_host_step_equals_device_step(lv, orc, torch_cuda):
    """lv_step_host (pinned, chunked, 3-slot pipeline) returns exactly what one device launch computes."""
    w, h, ns, nf = 1280, 720, 5, 7            # > 32 MiB of input: several chunks, ragged last chunk
    frames = orc.gen_batch("camera", 11, ns, nf, w, h)
    prev = np.zeros((ns, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0, want_mask=True, n_threads=4)
    ctx = lv.Context(w, h, ns, 20, 0)
    n = ns * nf
    pin = lv.pinned(n * ctx.frame_bytes)
    pin.array[:] = frames.reshape(-1)
    bgr = lv.pinned(n * ctx.bgr_bytes)
    bits = np.zeros(n * ctx.mask_units, np.uint16)
    cnt = np.zeros(n * ctx.mbs, np.uint16)
    mp = np.zeros(n * ctx.mbs, np.uint8)
    summ = np.zeros(2 * n, np.uint32)
    ctx.step_host(pin.array, ns, nf, bgr=bgr.array, mask_bits=bits, mb_count=cnt, mb_map=mp, summary=summ)
    assert np.array_equal(bgr.array, ref["bgr"])
    assert np.array_equal(cnt, ref["mb_count"]) and np.array_equal(mp, ref["mb_map"])
    assert np.array_equal(summ.reshape(-1, 2), ref["summary"])
    px = w * h
    assert np.array_equal(bits, np.concatenate([orc.pack_mask16(ref["mask"][i * px:(i + 1) * px], w, h) for i in range(n)]))
    for s in range(ns):
        assert np.array_equal(ctx.get_prev_luma(s), prev[s])
    # pageable host memory and a subset of outputs
    summ2 = np.zeros(2 * n, np.uint32)
    ctx.reset_prev_luma()
    ctx.step_host(frames.reshape(-1).copy(), ns, nf, summary=summ2)
    assert np.array_equal(summ2, summ)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ns,nf,chunk_mb", [(1, 30, None), (3, 4, "64"), (4, 9, "4")])
def test_host_step_small_outputs_whole_step_staging(lv, orc, torch_cuda, monkeypatch, ns, nf, chunk_mb):
    """The macroblock counts / map / summaries of a host step are staged for the whole step on the device and downloaded
    once: a single stream cut into several chunks and a chunk that holds every frame of its streams write straight into
    that staging, any other chunk shape is placed by k_place_small.  All three against the oracle, twice in a row (the
    second call reuses the staging and carries the state)."""
    if chunk_mb:
        monkeypatch.setenv("LV_HOST_CHUNK_MB", chunk_mb)
    w, h = 1280, 720
    frames = orc.gen_batch("camera", 23, ns, 2 * nf, w, h)
    prev = np.zeros((ns, w * h), np.uint8)
    ctx = lv.Context(w, h, ns, 20, 0)
    n = ns * nf
    for half in range(2):
        part = np.ascontiguousarray(frames[:, half * nf:(half + 1) * nf])
        ref = orc.convert_diff_batch(part, prev, ns, nf, w, h, 20, 0, n_threads=4)
        bgr = np.zeros(n * ctx.bgr_bytes, np.uint8)
        cnt = np.zeros(n * ctx.mbs, np.uint16)
        mp = np.full(n * ctx.mbs, 7, np.uint8)
        summ = np.zeros(2 * n, np.uint32)
        ctx.step_host(part.reshape(-1), ns, nf, bgr=bgr, mb_count=cnt, mb_map=mp, summary=summ)
        assert np.array_equal(bgr, ref["bgr"])
        assert np.array_equal(cnt, ref["mb_count"]) and np.array_equal(mp, ref["mb_map"])
        assert np.array_equal(summ.reshape(-1, 2), ref["summary"])
        # only one of the small outputs asked for
        if half == 0:
            ctx2 = lv.Context(w, h, ns, 20, 0)
            mp2 = np.zeros(n * ctx.mbs, np.uint8)
            ctx2.step_host(part.reshape(-1), ns, nf, mb_map=mp2)
            assert np.array_equal(mp2, ref["mb_map"])
            ctx2.close()
    ctx.close()


def test_device_generator_matches_oracle_generator(lv, orc, torch_cuda):
    for kind in ("uniform", "camera"):
        d = lv.gen_synthetic(kind, 1, 2, 5, 2, 3, 352, 288)
        torch_cuda.cuda.synchronize()
        want = orc.gen_batch(kind, 1, 2, 3, 352, 288, first_frame=5, first_stream=2)
        assert np.array_equal(d.cpu().numpy(), want.reshape(-1)), kind


def test_errors_are_codes_not_crashes(lv, torch_cuda):
    torch = torch_cuda
    ctx = lv.Context(20, 6)                       # legal for the legacy entries, not for the batched kernels
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    with pytest.raises(lv.LiveVideoError) as ei:
        ctx.convert_batch(buf, 1, buf)
    assert ei.value.code == -2
    ctx.close()
    ctx = lv.Context(352, 288, 2)
    out = ctx.alloc_outputs(2)
    d = torch.zeros(2 * ctx.frame_bytes, dtype=torch.uint8, device="cuda")
    with pytest.raises(lv.LiveVideoError) as ei:
        ctx.convert_diff_batch(d, 2, 1, out, first_stream=1)      # stream range exceeds max_streams
    assert ei.value.code == -1
    L = lv.lib()
    assert L.lv_convert_diff_batch(ctx._h, None, 0, 1, 1, None, None, None, None, None, None, None) == -1
    with pytest.raises(ValueError):
        ctx.convert_batch(torch.zeros(10, dtype=torch.uint8, device="cuda"), 1, out["bgr"])       # too small
    ctx.convert_diff_batch(d, 2, 0, out)          # empty batch: a no-op, not an error
    ctx.close()
    h = C.c_void_p()
    assert L.lv_create(99, 352, 288, 1, 20, 0, C.byref(h)) == -3


# ---- WriteBks background subtraction and the object box ---------------------------------

@pytest.mark.parametrize("w,h", [(352, 288), (1920, 1080), (320, 240), (20, 6), (4, 2), (1280, 720)])
def test_write_bks_matches_reference_loops(lv, orc, torch_cuda, w, h):
    torch = torch_cuda
    rng = np.random.default_rng(w + h)
    n = 3
    fb = w * h * 3 // 2
    bkd = rng.integers(0, 256, fb, dtype=np.uint8)
    cur = np.clip(bkd.astype(int)[None] + rng.integers(-40, 41, (n, fb)), 0, 255).astype(np.uint8)
    ctx = lv.Context(w, h)
    for tol in (0, 20, 25, 255):
        d_out = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
        bits = torch.zeros(n * ctx.mask_units, dtype=torch.int16, device="cuda") if w % 16 == 0 else None
        ctx.write_bks(torch.from_numpy(cur.reshape(-1)).cuda(), torch.from_numpy(bkd).cuda(), n, d_out, tol=tol, obj_bits=bits)
        got = d_out.cpu().numpy().reshape(n, fb)
        for i in range(n):
            assert np.array_equal(got[i], orc.write_bks(cur[i], bkd, w, h, tol)), (w, h, tol, i)
            if bits is not None:
                mask, _, _, _ = orc.frame_diff_mb(cur[i, :w * h], bkd[:w * h], w, h, tol, 0)
                want = orc.pack_mask16(mask, w, h)
                assert np.array_equal(bits.cpu().numpy().view(np.uint16)[i * ctx.mask_units:(i + 1) * ctx.mask_units], want)
    ctx.close()


def test_object_box_matches_reference_scan(lv, orc, torch_cuda):
    torch = torch_cuda
    w, h = 352, 288
    rng = np.random.default_rng(9)
    ctx = lv.Context(w, h)
    for density in (0.0, 0.0005, 0.05):
        mask = (rng.random(w * h) < density).astype(np.uint8)
        bits = orc.pack_mask16(mask, w, h)
        d_mask = torch.from_numpy(mask).cuda()
        d_bits = torch.from_numpy(bits.view(np.int16)).cuda()
        for (x0, x1, y0, y1) in [(0, w - 1, 0, h - 1), (10, 200, 5, 100), (17, 17, 40, 40), (300, 351, 250, 287)]:
            want = orc.object_box(mask, w, x0, x1, y0, y1)
            assert ctx.object_box(d_mask, x0, x1, y0, y1) == want
            assert ctx.object_box(d_bits, x0, x1, y0, y1, mask_is_bits=True) == want
    ctx.close()


@pytest.mark.parametrize("w,h", [(352, 288), (1920, 1080), (64, 48), (20, 12), (4096, 24)])
def test_object_box_batch_matches_reference_scan(lv, orc, torch_cuda, w, h):
    """One pass over a batch of masks (the vector kernels when w % 16 == 0, 128-pixel items for the bit mask with ragged
    row tails at 352 and 1920 -> w/16 not a multiple of 8) against the reference's per-picture scan; values other than
    1 in the byte mask are not 'on' (image.c tests == 1)."""
    torch = torch_cuda
    rng = np.random.default_rng(w + h)
    n = 5
    ctx = lv.Context(w, h)
    dens = [0.0, 0.0002, 0.01, 0.3, 1.0]
    masks = np.stack([(rng.random(w * h) < d).astype(np.uint8) for d in dens])
    noisy = masks.copy()
    noisy[2][masks[2] == 0] = rng.choice(np.array([0, 2, 255], np.uint8), size=int((masks[2] == 0).sum()))
    windows = [(0, w - 1, 0, h - 1), (w // 3, w // 3, h // 2, h // 2), (1, w - 2, 1, h - 2), (w // 2 - 3, w // 2 + 9, 0, h - 1),
               (w - 5, w - 1, h - 3, h - 1), (0, min(w - 1, 130), 2, min(h - 1, 9))]
    d_bytes = torch.from_numpy(noisy.reshape(-1)).cuda()
    for (x0, x1, y0, y1) in windows:
        want = np.array([orc.object_box(masks[i], w, x0, x1, y0, y1) for i in range(n)], np.int32)
        got = ctx.object_box_batch(d_bytes, n, x0, x1, y0, y1).cpu().numpy()
        assert np.array_equal(got, want), ("bytes", w, h, x0, x1, y0, y1)
        if w % 16 == 0:
            bits = np.concatenate([orc.pack_mask16(masks[i], w, h) for i in range(n)])
            d_bits = torch.from_numpy(bits.view(np.int16)).cuda()
            got = ctx.object_box_batch(d_bits, n, x0, x1, y0, y1, mask_is_bits=True).cpu().numpy()
            assert np.array_equal(got, want), ("bits", w, h, x0, x1, y0, y1)
    ctx.close()


# ---- every kernel variant computes the same bytes --------------------------------------------------------------

@pytest.mark.parametrize("variant", ["tile", "tma", "tile_late", "tile_early"])
def test_kernel_variants_are_bit_identical(lv, orc, torch_cuda, variant):
    """tile (default) and the TMA pipeline, over aligned, ragged and partial-band geometries,
    with and without the mask outputs, plus the conversion-only and difference-only entries."""
    torch = torch_cuda
    lv.set_kernel_variant(variant)
    try:
        assert lv.get_kernel_variant() == variant
        for (w, h, ns, nf) in [(1920, 1080, 1, 3), (352, 288, 2, 3), (1280, 720, 3, 2), (96, 20, 2, 2), (64, 18, 1, 3),
                               (4096, 24, 1, 2), (160, 8, 2, 2), (48, 18, 1, 2), (16, 216, 2, 2), (32, 40, 1, 3)]:       # 16 x 216: one macroblock per row
            frames = orc.gen_batch("camera" if min(w, h) >= 16 else "uniform", 5, ns, nf, w, h)
            prev = np.zeros((ns, w * h), np.uint8)
            ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
            for masks in (True, False):
                got = run_fused(lv, torch, frames, ns, nf, w, h, 20, 0, masks=masks)
                assert_matches_oracle(orc, got, ref, w, h, ns * nf, masks=masks)
                assert np.array_equal(got["state"], prev), (variant, w, h)
            ctx = lv.Context(w, h)
            d_out = torch.empty(ns * nf * ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
            ctx.convert_batch(torch.from_numpy(frames.reshape(-1)).cuda(), ns * nf, d_out)
            assert np.array_equal(d_out.cpu().numpy(), ref["bgr"]), (variant, "convert", w, h)
            cur = np.ascontiguousarray(frames[:, :, :w * h]).reshape(-1)
            refy = np.roll(cur, 7)
            out = ctx.alloc_outputs(ns * nf, bgr=False)
            ctx.diff_batch(torch.from_numpy(cur).cuda(), torch.from_numpy(refy).cuda(), ns * nf, out)
            torch.cuda.synchronize()
            for i in range(ns * nf):
                _, cnt, mp, _ = orc.frame_diff_mb(cur[i * w * h:(i + 1) * w * h], refy[i * w * h:(i + 1) * w * h], w, h, 20, 0)
                assert np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16)[i * ctx.mbs:(i + 1) * ctx.mbs], cnt)
            ctx.close()
    finally:
        lv.set_kernel_variant("tile")


# ---- BASELINE.json configurations at full size --------------------------------------------------------------

def _full_size(lv, orc, torch, w, h, ns, nf, kind="camera"):
    n = ns * nf
    ctx = lv.Context(w, h, ns, 20, 0)
    if min(w, h) >= 16:
        d_in = lv.gen_synthetic(kind, 1, 0, 0, ns, nf, w, h)
    else:                                   # the device generator works on whole macroblocks: generate on the host
        d_in = torch.from_numpy(orc.gen_batch("uniform", 1, ns, nf, w, h).reshape(-1)).cuda()
    out = ctx.alloc_outputs(n, mask_bits=True)
    ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    frames = d_in.cpu().numpy().reshape(ns, nf, -1)
    prev = np.zeros((ns, w * h), np.uint8)
    import os
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0, want_mask=False,
                                 n_threads=min(16, os.cpu_count() or 1))
    cnt = out["mb_count"].cpu().numpy().view(np.uint16)
    mp = out["mb_map"].cpu().numpy()
    summ = out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2)
    assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
    assert np.array_equal(cnt, ref["mb_count"]) and np.array_equal(mp, ref["mb_map"])
    assert np.array_equal(summ, ref["summary"])
    # size-independent properties
    assert np.array_equal(summ[:, 0], cnt.reshape(n, -1).sum(1))                  # checksum of checksums
    assert np.array_equal(summ[:, 1], mp.reshape(n, -1).sum(1))
    assert np.array_equal(mp, (cnt > 0).astype(np.uint8))
    bits = out["mask_bits"].cpu().numpy().view(np.uint16)
    pop = np.unpackbits(bits.view(np.uint8)).reshape(n, -1).sum(1)
    assert np.array_equal(pop, summ[:, 0])                                        # mask popcount == changed pixels
    for s in range(ns):
        assert np.array_equal(ctx.get_prev_luma(s), prev[s])
    ctx.close()


def test_config2_1080p_batch64(lv, orc, torch_cuda):
    _full_size(lv, orc, torch_cuda, 1920, 1080, 1, 64)


def test_config3_4k_batch32(lv, orc, torch_cuda):
    _full_size(lv, orc, torch_cuda, 3840, 2160, 1, 32)


@pytest.mark.parametrize("w,h,ns,nf", [(7680, 4320, 1, 3), (16384, 8192, 1, 2), (65536, 2, 2, 3), (16, 32768, 1, 2),
                                        (32768, 16, 3, 2), (2048, 2048, 1, 5), (2064, 1038, 2, 3)])
def test_extreme_geometries_full_outputs(lv, orc, torch_cuda, w, h, ns, nf):
    """8K, a 128-megapixel picture, one row pair of 65 536 pixels, a 16-pixel column of 32 768 rows, the two sides of the
    2 048-pixel load-order rule: every output against the oracle plus the size-independent properties."""
    _full_size(lv, orc, torch_cuda, w, h, ns, nf, kind="camera" if min(w, h) >= 64 else "uniform")


def test_largest_picture_the_api_accepts(lv, orc, torch_cuda):
    """2^30 pixels (32 768 x 32 768), the largest picture lv_create takes: the kernel's 32-bit in-frame offsets reach 3 * 2^30 in
    the BGR plane.  One frame, every output against the oracle; one pixel more is refused."""
    import psutil
    if psutil.virtual_memory().available < 24 << 30:
        pytest.skip("needs ~16 GB of host memory for the picture, its BGR copy and the oracle's")
    with pytest.raises(lv.LiveVideoError):
        lv.Context(32768 + 16, 32768)
    _full_size(lv, orc, torch_cuda, 32768, 32768, 1, 1)


def test_batch_whose_outputs_pass_4_gib(lv, orc, torch_cuda):
    """180 frames of 4K in one launch: the BGR batch is 4.48 GB, so the last frames sit beyond 2^32 bytes (the kernel's 64-bit
    frame bases + 32-bit in-frame offsets); every output against the oracle."""
    import psutil
    if psutil.virtual_memory().available < 32 << 30:
        pytest.skip("needs ~16 GB of host memory")
    _full_size(lv, orc, torch_cuda, 3840, 2160, 3, 60)


def test_config4_one_gpu_share_of_256_720p_streams(lv, orc, torch_cuda):
    # the 8-GPU share of config 4: 32 streams x 16 frames of 720p in one launch
    _full_size(lv, orc, torch_cuda, 1280, 720, 32, 16)


def test_virtual_ranks_on_one_gpu_equal_single_rank(lv, orc, torch_cuda):
    """Stream sharding is embarrassingly parallel: G virtual ranks (one context each, same GPU) produce, block by
    block, exactly the single-rank result ('Multi-GPU without a cluster')."""
    torch = torch_cuda
    from livevideo_b200.shard import ShardedContext, shard_range
    w, h, S, T = 352, 288, 7, 4
    whole = ShardedContext(w, h, S, rank=0, world=1, device=0)
    d_in = lv.gen_synthetic("camera", 3, 0, 0, S, T, w, h)
    out = whole.ctx.alloc_outputs(S * T)
    full = whole.step(d_in, T, out).cpu().numpy()
    cnt_full = out["mb_count"].cpu().numpy().reshape(S, T, -1)
    for G in (2, 3):
        parts = []
        for r in range(G):
            first, count = shard_range(S, G, r)
            sc = ShardedContext(w, h, S, rank=r, world=G, device=0)
            fb = sc.ctx.frame_bytes
            o = sc.ctx.alloc_outputs(max(count, 1) * T)
            sc.step(d_in[first * T * fb:(first + count) * T * fb], T, o, gather=False)
            torch.cuda.synchronize()
            parts.append(o["summary"][:count * T * 2].cpu().numpy().reshape(count, T, 2))
            assert np.array_equal(o["mb_count"][:count * T * sc.ctx.mbs].cpu().numpy().reshape(count, T, -1),
                                  cnt_full[first:first + count])
        assert np.array_equal(np.concatenate(parts), full)


def _single_process_shard(lv, orc, torch, devices, S=5, T=3, w=352, h=288):
    from livevideo_b200.shard import SingleProcessShard
    sh = SingleProcessShard(w, h, S, devices)
    frames = orc.gen_batch("camera", 4, S, T, w, h)
    prev = np.zeros((S, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, S, T, w, h, 20, 0, want_mask=False)
    fb = w * h * 3 // 2
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    ins, bgrs, cnts, maps = [], [], [], []
    for (dev, first, n) in sh.ranges:
        d = torch.device("cuda", dev)
        ins.append(torch.from_numpy(frames[first:first + n].reshape(-1).copy()).to(d) if n else None)
        bgrs.append(torch.empty(max(n, 1) * T * 3 * w * h, dtype=torch.uint8, device=d))
        cnts.append(torch.empty(max(n, 1) * T * mbs, dtype=torch.int16, device=d))
        maps.append(torch.empty(max(n, 1) * T * mbs, dtype=torch.uint8, device=d))
    for _ in range(2):                      # second step: frame 0 differenced against the carried state
        sh.step(ins, T, bgrs, cnts, maps)
        table = sh.gather(T)
    sh.sync()
    ref2 = orc.convert_diff_batch(frames, prev, S, T, w, h, 20, 0, want_mask=False)   # prev now holds the last luma
    assert np.array_equal(table, ref2["summary"].reshape(S, T, 2))
    for i, (dev, first, n) in enumerate(sh.ranges):
        if n:
            assert np.array_equal(bgrs[i][: n * T * 3 * w * h].cpu().numpy(), ref2["bgr"][first * T * 3 * w * h:(first + n) * T * 3 * w * h])
            assert np.array_equal(cnts[i][: n * T * mbs].cpu().numpy().view(np.uint16), ref2["mb_count"][first * T * mbs:(first + n) * T * mbs])
    sh.close()
    return ref


def test_single_process_shard_one_device(lv, orc, torch_cuda):
    _single_process_shard(lv, orc, torch_cuda, [0])


def test_single_process_shard_all_devices_nccl(lv, orc, torch_cuda):
    n = torch_cuda.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2): NCCL cannot build a communicator over one device twice")
    _single_process_shard(lv, orc, torch_cuda, list(range(n)))
    _single_process_shard(lv, orc, torch_cuda, list(range(n)), S=1)      # more devices than streams: empty blocks


def test_upload_slots_double_buffering(lv, orc, torch_cuda):
    """lv_upload_async / wait / release: host batches alternate between two device slots; results equal one big launch."""
    torch = torch_cuda
    w, h, T, B = 352, 288, 4, 5
    frames = orc.gen_batch("camera", 6, 1, T * B, w, h)
    prev = np.zeros((1, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, 1, T * B, w, h, 20, 0, want_mask=False)
    ctx = lv.Context(w, h)
    pins = [lv.pinned(T * ctx.frame_bytes) for _ in range(2)]
    outs = [ctx.alloc_outputs(T) for _ in range(2)]
    got_cnt, got_bgr = [], []
    for b in range(B):
        s = b & 1
        if b >= 2:
            torch.cuda.current_stream().synchronize()      # host buffer s is free again (its copy has completed)
        pins[s].array[:] = frames[0, b * T:(b + 1) * T].reshape(-1)
        d = ctx.upload_async(pins[s].array, T, s)
        ctx.upload_wait(s)
        ctx.convert_diff_batch(d, 1, T, outs[s])
        ctx.upload_release(s)
        got_cnt.append(outs[s]["mb_count"].cpu().numpy().view(np.uint16).copy())
        got_bgr.append(outs[s]["bgr"].cpu().numpy().copy())
    assert np.array_equal(np.concatenate(got_cnt), ref["mb_count"])
    assert np.array_equal(np.concatenate(got_bgr), ref["bgr"])
    ctx.close()


@pytest.mark.parametrize("name,kind", [("rgb24_to_yv12", 1), ("yvu9_to_yv12", 2), ("yuy2_to_yv12", 3), ("yv12_to_yvu9", 4),
                                       ("yv12_to_yuy2", 5)])
def test_other_converters_match_object_code_semantics(lv, orc, torch_cuda, name, kind):
    """The five Convert.h methods the application never calls: legacy host entry and batched device entry vs the
    oracle (itself pinned to the object code on all 2^24 RGB triples / random planes)."""
    torch = torch_cuda
    csc = lv.ColorSpaceConversions()
    mode, inb, outb = orc.FORMATS[name]
    rng = np.random.default_rng(kind)
    method = getattr(csc, {"rgb24_to_yv12": "RGB24_to_YV12", "yvu9_to_yv12": "YVU9_to_YV12", "yuy2_to_yv12": "YUY2_to_YV12",
                           "yv12_to_yvu9": "YV12_to_YVU9", "yv12_to_yuy2": "YV12_to_YUY2"}[name])
    for (w, h) in [(16, 8), (64, 48), (352, 288), (1920, 1080), (20, 12)]:
        n = 3
        data = rng.integers(0, 256, (n, w * h // 8 * inb), dtype=np.uint8)
        want = np.stack([orc.convert_format(name, data[i], w, h) for i in range(n)])
        out = np.zeros(w * h // 8 * outb, np.uint8)
        method(data[0].copy(), out, w, h)
        assert np.array_equal(out, want[0]), (name, "legacy", w, h)
        ctx = lv.Context(w, h)
        d_out = torch.zeros(n * w * h // 8 * outb, dtype=torch.uint8, device="cuda")
        ctx.convert_format(kind, torch.from_numpy(data.reshape(-1)).cuda(), d_out, n)
        assert np.array_equal(d_out.cpu().numpy(), want.reshape(-1)), (name, "device", w, h)
        ctx.close()
    if name == "rgb24_to_yv12":       # every (R,G,B) once as a chroma-producing pixel
        vals = np.arange(1 << 24, dtype=np.uint32)
        ctx = lv.Context(2048, 2048)
        for c in range(16):
            v = vals[c << 20:(c + 1) << 20]
            blk = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(1024, 1024, 3)
            img = np.repeat(np.repeat(blk, 2, axis=0), 2, axis=1).reshape(-1)
            d_out = torch.zeros(2048 * 2048 * 3 // 2, dtype=torch.uint8, device="cuda")
            ctx.convert_format(1, torch.from_numpy(img).cuda(), d_out, 1)
            assert np.array_equal(d_out.cpu().numpy(), orc.convert_format(name, img, 2048, 2048))
        ctx.close()


# ---- peer-memory exchange of the summaries (lv_exchange_*) -------------------------------------------------------
def _local_exchanges(world, slot_words, depth=8):
    from livevideo_b200.shard import PeerExchange
    xs = [PeerExchange(0, r, world, slot_words, depth=depth) for r in range(world)]
    PeerExchange.connect_local(xs)
    return xs


def test_peer_exchange_ring_on_one_gpu(lv, torch_cuda):
    """The two halves on their own, G ranks in one process on one GPU (raw pointers instead of IPC): over many more
    steps than the ring is deep, every rank's table of every step holds every rank's block (short blocks are padded
    with zeros), and no wait ever gives up."""
    torch = torch_cuda
    torch.cuda.set_device(0)                        # an earlier multi-GPU test may have left another device current
    G, words, depth, steps = 3, 20, 4, 23           # not a multiple of the ring
    xs = _local_exchanges(G, words, depth)
    streams = [torch.cuda.Stream() for _ in range(G)]
    tables = [[torch.zeros(G * words, dtype=torch.int32, device="cuda") for _ in range(steps)] for _ in range(G)]
    rng = np.random.default_rng(5)
    blocks = rng.integers(-2**31, 2**31 - 1, size=(steps, G, words), dtype=np.int64).astype(np.int32)
    used = [words, words - 3, 1]                    # rank r publishes used[r] words; the rest of its row must read 0
    want = blocks.copy()
    for r in range(G):
        want[:, r, used[r]:] = 0
    d_blocks = torch.from_numpy(blocks).cuda()
    torch.cuda.synchronize()
    lag = 2
    for k in range(steps + lag):
        for r in range(G):
            if k < steps:
                xs[r].publish(k, d_blocks[k, r], used[r], streams[r])
            if k >= lag:
                xs[r].collect(k - lag, tables[r][k - lag], streams[r])
    torch.cuda.synchronize()
    for r in range(G):
        assert xs[r].error() == 0
        for k in range(steps):
            assert np.array_equal(tables[r][k].cpu().numpy().reshape(G, words), want[k]), (r, k)
    # call-order errors are reported, not executed
    from livevideo_b200.capi import LiveVideoError
    with pytest.raises(LiveVideoError):
        xs[0].publish(steps + 5, d_blocks[0, 0], words, streams[0])      # not the next step
    with pytest.raises(LiveVideoError):
        xs[0].collect(steps, None, streams[0])                           # not published yet
    for x in xs:
        x.close()


def test_sharded_step_async_over_peer_exchange_equals_single_rank(lv, orc, torch_cuda):
    """ShardedContext.step_async over the peer exchange (virtual ranks on one GPU): every rank ends up with the table a
    single rank computes, step after step, across more steps than the ring is deep; out['summary'] keeps the local
    block."""
    torch = torch_cuda
    torch.cuda.set_device(0)
    from livevideo_b200.shard import PeerExchange, ShardedContext, max_shard, shard_range
    w, h, S, T, G, steps = 64, 48, 5, 3, 2, 11
    frames = [orc.gen_batch("camera", 7 + k, S, T, w, h) for k in range(steps)]
    prev = np.zeros((S, w * h), np.uint8)
    want = [orc.convert_diff_batch(f, prev, S, T, w, h, 20, 0, want_mask=False)["summary"].reshape(S, T, 2)
            for f in frames]      # the oracle carries `prev` from step to step
    words = (max_shard(S, G) * T * 2 + 3) & ~3
    xs = [PeerExchange(0, r, G, words, depth=ShardedContext.RING) for r in range(G)]
    PeerExchange.connect_local(xs)
    scs = [ShardedContext(w, h, S, rank=r, world=G, device=0, exchange=xs[r]) for r in range(G)]
    outs = [sc.ctx.alloc_outputs(max(sc.local_streams, 1) * T) for sc in scs]
    fb = scs[0].ctx.frame_bytes
    streams = [torch.cuda.Stream() for _ in range(G)]
    for k in range(steps):
        handles = []
        for r, sc in enumerate(scs):
            first, count = shard_range(S, G, r)
            d_in = torch.from_numpy(frames[k][first:first + count].reshape(-1).copy()).cuda()
            with torch.cuda.stream(streams[r]):
                handles.append(sc.step_async(d_in, T, outs[r]))
        if k % 3 == 2 or k == steps - 1:                     # ask only now and then: the others are retired unread
            for hd in handles:                               # one host thread drives both ranks: push both blocks first
                hd.flush()
            for r, hd in enumerate(handles):
                got = hd.result().cpu().numpy().view(np.uint32)
                assert np.array_equal(got, want[k]), (k, r)
                first, count = shard_range(S, G, r)
                torch.cuda.synchronize()
                mine = outs[r]["summary"][:count * T * 2].cpu().numpy().view(np.uint32).reshape(count, T, 2)
                assert np.array_equal(mine, want[k][first:first + count])
    for x in xs:
        assert x.error() == 0


# ---- the drop-in entries as an application calls them -------------------------------------------------------------
def test_legacy_entries_from_several_threads_and_registered_buffers(lv, orc, torch_cuda):
    """YV12_to_RGB24 / FrameDiff_MB are thread-safe (a context per thread) and give the same bytes whether the
    caller's buffers are pageable or pinned in place with lv_host_register (the reference mallocs its planes once,
    MainDlg.cpp:39-51)."""
    import mmap
    import threading
    w, h = 1280, 720
    px = w * h
    csc = lv.ColorSpaceConversions()
    frames = orc.gen_batch("camera", 11, 4, 2, w, h)          # 4 threads x (reference frame, current frame)
    prev = np.zeros((4, px), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, 4, 2, w, h, 20, 0)
    errors = []

    def worker(i, register):
        try:
            bufs = [np.frombuffer(mmap.mmap(-1, n), dtype=np.uint8, count=n) for n in (px, px // 4, px // 4, 3 * px, px, px)]
            y, u, v, dst, ry, mask = bufs
            cur = frames[i, 1]
            y[:], u[:], v[:], ry[:] = cur[:px], cur[px:px + px // 4], cur[px + px // 4:], frames[i, 0, :px]
            if register:
                for a in bufs:
                    lv.host_register(a)
            for _ in range(3):
                dst[:] = 0
                csc.YV12_to_RGB24(y, u, v, dst, w, h)
                m, cnt, mp = lv.FrameDiff_MB(y, ry, w, h, 20, 0, mask=mask)
                f = i * 2 + 1
                assert np.array_equal(dst, ref["bgr"][f * 3 * px:(f + 1) * 3 * px])
                assert np.array_equal(m, ref["mask"][f * px:(f + 1) * px])
                mbs = cnt.size
                assert np.array_equal(cnt, ref["mb_count"][f * mbs:(f + 1) * mbs])
                assert np.array_equal(mp, ref["mb_map"][f * mbs:(f + 1) * mbs])
            if register:
                for a in bufs:
                    lv.host_unregister(a)
        except Exception as e:  # noqa: BLE001
            errors.append((i, register, repr(e)))

    for register in (False, True):
        ths = [threading.Thread(target=worker, args=(i, register)) for i in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    assert not errors, errors
    # the calling thread's stream and scratch stay on the device of its first call even when something else in the
    # process makes another device current in between
    torch = torch_cuda
    if torch.cuda.device_count() >= 2:
        cur = frames[0, 1]
        dst = np.zeros(3 * px, np.uint8)
        csc.YV12_to_RGB24(cur[:px].copy(), cur[px:px + px // 4].copy(), cur[px + px // 4:].copy(), dst, w, h)
        torch.cuda.set_device(1)
        try:
            dst2 = np.zeros(3 * px, np.uint8)
            csc.YV12_to_RGB24(cur[:px].copy(), cur[px:px + px // 4].copy(), cur[px + px // 4:].copy(), dst2, w, h)
            assert torch.cuda.current_device() == 1
        finally:
            torch.cuda.set_device(0)
        assert np.array_equal(dst, dst2) and np.array_equal(dst, ref["bgr"][3 * px:6 * px])


@pytest.mark.parametrize("sx,sy,cl,ct,w,h", [(1920, 1088, 0, 0, 1920, 1080), (384, 304, 16, 8, 352, 288), (64, 48, 0, 0, 64, 48),
                                             (48, 32, 4, 2, 40, 24), (1296, 736, 8, 6, 1280, 720), (96, 64, 32, 16, 64, 48)])
def test_narrow_pictures_matches_img2buf(lv, orc, torch_cuda, sx, sy, cl, ct, w, h):
    """lv_narrow_pictures = the reference's img2buf + write_out_picture (lib/JM/output.c:87-178, 362-453) on the device:
    vector kernel (16-byte aligned crops) and the generic one, 10-bit and full 16-bit content (upper byte dropped, not
    clipped), several pictures per launch; then straight into the fused step."""
    torch = torch_cuda
    n = 3
    rng = np.random.default_rng(sx + w)
    samples = sx * sy * 3 // 2
    pics = rng.integers(0, 65536, (n, samples), dtype=np.uint16)
    pics[1] &= 0x3ff
    want = np.stack([orc.picture16_to_i420(pics[i], sx, sy, cl, sx - cl - w, ct, sy - ct - h) for i in range(n)])
    ctx = lv.Context(w, h)
    d_pic = torch.from_numpy(pics.view(np.int16).reshape(-1)).cuda()
    d_out = torch.zeros(n * ctx.frame_bytes, dtype=torch.uint8, device="cuda")
    ctx.narrow_pictures(d_pic, n, sx, sy, cl, ct, d_out)
    assert np.array_equal(d_out.cpu().numpy(), want.reshape(-1)), (sx, sy, cl, ct, w, h)
    if w % 16 == 0:
        out = ctx.alloc_outputs(n)
        ctx.convert_diff_batch(d_out, 1, n, out)
        ref = orc.convert_diff_batch(want.reshape(1, n, -1), np.zeros((1, w * h), np.uint8), 1, n, w, h, 20, 0, want_mask=False)
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
    from livevideo_b200.capi import LiveVideoError
    with pytest.raises(LiveVideoError):
        ctx.narrow_pictures(d_pic, n, sx, sy, cl + 1, ct, d_out)        # odd crop
    with pytest.raises(LiveVideoError):
        ctx.narrow_pictures(d_pic, n, sx, sy, sx - w + 2, ct, d_out)    # rectangle leaves the picture
    ctx.close()


def test_step_host_pictures_equals_narrow_then_step(lv, orc, torch_cuda):
    """lv_step_host_pictures: raw decoder pictures in host memory -> (device) img2buf/write_out_picture -> fused step,
    chunked; equals the oracle's picture16_to_i420 followed by the fused oracle step, state carried across two calls."""
    sx, sy, cl, ct, w, h = 384, 304, 16, 8, 352, 288
    ns, nf = 2, 5
    rng = np.random.default_rng(21)
    samples = sx * sy * 3 // 2
    ctx = lv.Context(w, h, ns)
    prev = np.zeros((ns, w * h), np.uint8)
    import os
    os.environ["LV_HOST_CHUNK_MB"] = "1"           # several chunks per call (read once per process: harmless if already set)
    for call in range(2):
        pics = (rng.integers(0, 256, (ns, nf, samples), dtype=np.uint16) | (rng.integers(0, 4, (ns, nf, samples), dtype=np.uint16) << 8))
        frames = np.stack([[orc.picture16_to_i420(pics[s, t], sx, sy, cl, sx - cl - w, ct, sy - ct - h) for t in range(nf)]
                           for s in range(ns)])
        ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0, want_mask=False)
        n = ns * nf
        bgr, cnt = np.zeros(n * ctx.bgr_bytes, np.uint8), np.zeros(n * ctx.mbs, np.uint16)
        mp, sm = np.zeros(n * ctx.mbs, np.uint8), np.zeros(2 * n, np.uint32)
        ctx.step_host_pictures(np.ascontiguousarray(pics).reshape(-1), sx, sy, cl, ct, ns, nf, bgr=bgr, mb_count=cnt, mb_map=mp,
                               summary=sm)
        assert np.array_equal(bgr, ref["bgr"]) and np.array_equal(cnt, ref["mb_count"])
        assert np.array_equal(mp, ref["mb_map"]) and np.array_equal(sm.reshape(-1, 2), ref["summary"])
    ctx.close()


def test_more_than_65535_frames_per_call(lv, orc, torch_cuda):
    """gridDim.z is limited to 65535: long batches are cut into launch groups (by streams for the fused step, by frames
    for the single-purpose entries) without changing any result -- 3 streams x 30 000 tiny frames = 90 000 frames."""
    torch = torch_cuda
    w, h, ns, nf = 16, 2, 3, 30000
    rng = np.random.default_rng(65535)
    frames = rng.integers(0, 256, (ns, nf, w * h * 3 // 2), dtype=np.uint8)
    frames[:, 1::2, :w * h] = frames[:, 0:-1:2, :w * h]              # every other frame repeats its predecessor's luma
    prev = np.zeros((ns, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0, want_mask=False, n_threads=8)
    got = run_fused(lv, torch, frames, ns, nf, w, h, 20, 0, masks=False)
    assert np.array_equal(got["bgr"], ref["bgr"])
    assert np.array_equal(got["mb_count"], ref["mb_count"]) and np.array_equal(got["mb_map"], ref["mb_map"])
    assert np.array_equal(got["summary"], ref["summary"]) and np.array_equal(got["state"], prev)
    ctx = lv.Context(w, h)
    n = ns * nf
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    d_bgr = torch.empty(n * ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
    ctx.convert_batch(d_in, n, d_bgr)
    assert np.array_equal(d_bgr.cpu().numpy(), ref["bgr"])
    ctx.close()


def test_convert_diff_batch_ex_prezeroed_summary(lv, orc, torch_cuda):
    """lv_convert_diff_batch_ex with LV_STEP_SUMMARY_PREZEROED skips the summary memset: same results when the caller
    zeroed the buffer, and (what the flag means) the blocks ADD into whatever the buffer held."""
    torch = torch_cuda
    w, h, ns, nf = 352, 288, 2, 3
    frames = orc.gen_batch("camera", 13, ns, nf, w, h)
    ref = orc.convert_diff_batch(frames, np.zeros((ns, w * h), np.uint8), ns, nf, w, h, 20, 0, want_mask=False)
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    L = lv.lib()
    st = torch.cuda.current_stream().cuda_stream
    for base in (0, 7):
        ctx = lv.Context(w, h, ns)
        out = ctx.alloc_outputs(ns * nf)
        out["summary"].fill_(base)
        rc = L.lv_convert_diff_batch_ex(ctx._h, d_in.data_ptr(), 0, ns, nf, out["bgr"].data_ptr(), None, None,
                                        out["mb_count"].data_ptr(), out["mb_map"].data_ptr(), out["summary"].data_ptr(), st, 1)
        assert rc == 0
        torch.cuda.synchronize()
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"])
        assert np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"] + base)
        ctx.close()


def test_static_background_reference_mode(lv, orc, torch_cuda):
    """LV_REF_STATIC: every frame of a stream is differenced against the stream's stored plane (the static background
    BkdBuf of MainDlg.cpp:238-246), which the step leaves untouched -- frame by frame the reference's primitive
    (frame_diff_mb) against that plane; conversion unchanged; switching back resumes the previous-frame rule."""
    torch = torch_cuda
    w, h, ns, nf = 352, 288, 2, 4
    frames = orc.gen_batch("camera", 17, ns, nf, w, h)
    px = w * h
    bg = [orc.gen("camera", 17, s, 9, w, h)[:px].copy() for s in range(ns)]
    ctx = lv.Context(w, h, ns)
    for s in range(ns):
        ctx.set_prev_luma(s, bg[s])
    ctx.set_ref_mode(True)
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    out = ctx.alloc_outputs(ns * nf, mask_bytes=True)
    for _ in range(2):                                         # twice: the background must survive the step
        ctx.convert_diff_batch(d_in, ns, nf, out)
        torch.cuda.synchronize()
        got_mask = out["mask_bytes"].cpu().numpy().reshape(ns, nf, px)
        got_cnt = out["mb_count"].cpu().numpy().view(np.uint16).reshape(ns, nf, -1)
        got_sum = out["summary"].cpu().numpy().view(np.uint32).reshape(ns, nf, 2)
        got_bgr = out["bgr"].cpu().numpy().reshape(ns, nf, -1)
        for s in range(ns):
            for t in range(nf):
                mask, cnt, mp, changed = orc.frame_diff_mb(frames[s, t, :px], bg[s], w, h, 20, 0)
                assert np.array_equal(got_mask[s, t], mask) and np.array_equal(got_cnt[s, t], cnt), (s, t)
                assert got_sum[s, t, 0] == changed and got_sum[s, t, 1] == int(mp.sum())
                assert np.array_equal(got_bgr[s, t], orc.convert_frame(frames[s, t], w, h))
            assert np.array_equal(ctx.get_prev_luma(s), bg[s])
    ctx.set_ref_mode(False)
    prev = np.stack(bg)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
    ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    assert np.array_equal(out["mask_bytes"].cpu().numpy(), ref["mask"])
    assert np.array_equal(ctx.get_prev_luma(0), prev[0])
    ctx.close()
