"""Timing of every kernel beside the fused step (SURVEY.md 8(f) rows + the K2/K3 single-purpose entries), device
resident, CUDA events, against the measured HBM copy bandwidth.  Usage: bench_aux.py [W H N]   (default 1920 1080 64)
Each line: entry, us per launch, algorithmic bytes per px, GB/s, fraction of MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import livevideo_b200 as lv  # noqa: E402


def timeit(fn, iters=30, rounds=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


def main():
    w, h, n = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1920, 1080, 64)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        peak = 6650.0
    px = w * h
    ctx = lv.Context(w, h, 1)
    d_in = lv.gen_synthetic("camera", 1, 0, 0, 1, n, w, h)
    out = ctx.alloc_outputs(n, mask_bits=True, mask_bytes=True)
    lean = {k: v for k, v in out.items() if k not in ("mask_bits", "mask_bytes")}
    L = lv.lib()
    rows = []

    def row(name, us, bpp, frames=n, note=""):
        gbs = frames * px * bpp / us / 1e3
        rows.append((name, us, bpp, gbs, gbs / peak, note))
        print(f"{name:34s} {us:9.1f} us  {bpp:6.3f} B/px  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of measured   {note}", flush=True)

    print(f"# {w}x{h} x {n} frames, peak {peak} GB/s")
    row("fused convert+diff (K1)", timeit(lambda: ctx.convert_diff_batch(d_in, 1, n, lean)), 5.5)
    row("fused + bit mask", timeit(lambda: ctx.convert_diff_batch(d_in, 1, n, {**lean, "mask_bits": out["mask_bits"]})), 5.625)
    # the fused step with the object boxes accumulated by the same launch (lv_convert_diff_box_batch): windows of the
    # reference's size (an object of 96 x 64 px + 8 px margin) and the whole picture
    for tag, n_obj, whole in (("1 object window", 1, False), ("4 object windows", 4, False), ("whole picture", 1, True)):
        wins = torch.empty((n, n_obj, 4), dtype=torch.int32)
        for f in range(n):
            for k in range(n_obj):
                x0, y0 = (97 * f + 411 * k) % max(1, w - 112), (53 * f + 177 * k) % max(1, h - 80)
                wins[f, k] = torch.tensor([0, w - 1, 0, h - 1] if whole else [x0, x0 + 111, y0, y0 + 79])
        d_w = wins.cuda()
        boxes = torch.empty((n, n_obj, 4), dtype=torch.int32, device="cuda")
        row(f"fused + object box ({tag})", timeit(lambda: ctx.convert_diff_box_batch(d_in, 1, n, lean, d_w, boxes)), 5.5,
            note="no mask in memory; init kernel in place of the summary memset")
        row(f"plain step + box pass over the windows ({tag})",
            timeit(lambda: ctx.convert_diff_box_batch(d_in, 1, n, lean, d_w, boxes, windowed=True)), 5.5,
            note="LV_BOX_WINDOWED: init kernel + plain fused kernel + k_box_windows")
    row("convert only (K2)", timeit(lambda: ctx.convert_batch(d_in, n, out["bgr"])), 4.5)
    # difference only over explicit planes: cur = luma planes of the batch, ref = the same shifted by one frame
    ys = torch.stack([d_in[f * ctx.frame_bytes: f * ctx.frame_bytes + px] for f in range(n)]).reshape(-1).contiguous()
    ref = torch.roll(ys.reshape(n, px), 1, 0).reshape(-1).contiguous()
    row("diff only (K3)", timeit(lambda: ctx.diff_batch(ys, ref, n, {"mb_count": out["mb_count"], "mb_map": out["mb_map"], "summary": out["summary"]})), 2.0)
    bkd = d_in[: ctx.frame_bytes].clone()
    o420 = torch.empty_like(d_in)
    row("write_bks", timeit(lambda: ctx.write_bks(d_in, bkd, n, o420, tol=25)), 3.0, note="1.5 in + 1.5 out, background frame from L2")
    row("write_bks + bkd_obj bits", timeit(lambda: ctx.write_bks(d_in, bkd, n, o420, tol=25, obj_bits=out["mask_bits"])), 3.125)
    # object box: one frame per call in the C ABI (image.c:4818-4838 runs once per picture)
    ctx.convert_diff_batch(d_in, 1, n, out)
    box = torch.empty(4, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    mb, mbits = out["mask_bytes"], out["mask_bits"]
    row("object_box (byte mask, 1 frame)", timeit(lambda: L.lv_object_box(ctx._h, mb.data_ptr(), 0, 0, w - 1, 0, h - 1, box.data_ptr(), st)), 1.0, frames=1,
        note="launch-bound: 2 launches for one frame")
    row("object_box (bit mask, 1 frame)", timeit(lambda: L.lv_object_box(ctx._h, mbits.data_ptr(), 1, 0, w - 1, 0, h - 1, box.data_ptr(), st)), 0.125, frames=1)
    if hasattr(L, "lv_object_box_batch"):
        boxes = torch.empty(4 * n, dtype=torch.int32, device="cuda")
        row("object_box_batch (byte mask)", timeit(lambda: L.lv_object_box_batch(ctx._h, mb.data_ptr(), 0, n, 0, w - 1, 0, h - 1, boxes.data_ptr(), st)), 1.0)
        row("object_box_batch (bit mask)", timeit(lambda: L.lv_object_box_batch(ctx._h, mbits.data_ptr(), 1, n, 0, w - 1, 0, h - 1, boxes.data_ptr(), st)), 0.125)
    # decoder pictures (16-bit planes, 1920x1088 coded for 1080 lines) -> tight I420
    sx, sy = (w + 15) // 16 * 16, (h + 15) // 16 * 16
    pic = torch.randint(0, 1024, (n * (sx * sy * 3 // 2),), dtype=torch.int16, device="cuda")
    row("narrow_pictures (img2buf on device)", timeit(lambda: ctx.narrow_pictures(pic, n, sx, sy, 0, 0, o420)), 4.5,
        note=f"{sx}x{sy} 16-bit 4:2:0 in (3 B/px) + 1.5 B/px out")
    outp = ctx.alloc_outputs(n)
    row("narrow_pictures + fused step (2 kernels)", timeit(lambda: (ctx.narrow_pictures(pic, n, sx, sy, 0, 0, o420),
                                                                      ctx.convert_diff_batch(o420, 1, n, outp))), 7.0,
        note="algorithmic bytes of the ONE-kernel form: 3 in + 1 previous luma + 3 out")
    row("fused step from 16-bit pictures (1 kernel)", timeit(lambda: ctx.convert_diff_pictures(pic, sx, sy, 0, 0, 1, n, outp)), 7.0,
        note="3 B/px in (16-bit 4:2:0) + 1 B/px previous luma (2 B/px read from L2) + 3 B/px out; no I420 frame in memory")
    del outp
    names = {1: "RGB24_to_YV12", 2: "YVU9_to_YV12", 3: "YUY2_to_YV12", 4: "YV12_to_YVU9", 5: "YV12_to_YUY2"}
    for kind in range(1, 6):
        ib, ob = L.lv_format_bytes(kind, w, h, 0), L.lv_format_bytes(kind, w, h, 1)
        a = torch.randint(0, 256, (n * ib,), dtype=torch.uint8, device="cuda")
        b = torch.empty(n * ob, dtype=torch.uint8, device="cuda")
        row(names[kind], timeit(lambda: ctx.convert_format(kind, a, b, n)), (ib + ob) / px)
    return rows


if __name__ == "__main__":
    main()
