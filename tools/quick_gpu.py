"""Quick GPU sanity: parity of the fused kernel vs the oracle on small cases + raw kernel timing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import livevideo_b200 as lv  # noqa: E402
import lvoracle as orc  # noqa: E402


def parity(w, h, ns, nf, kind="camera", tol=20, thr=0):
    frames = orc.gen_batch(kind, 1, ns, nf, w, h)
    prev = np.zeros((ns, w * h), dtype=np.uint8)
    ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, thr)
    ctx = lv.Context(w, h, ns, tol, thr)
    d_in = torch.from_numpy(frames.reshape(-1)).cuda()
    out = ctx.alloc_outputs(ns * nf, mask_bits=True, mask_bytes=True)
    ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    ok = True
    for k in ("bgr", "mask_bytes", "mb_map"):
        same = np.array_equal(out[k].cpu().numpy(), ref[{"mask_bytes": "mask"}.get(k, k)])
        ok &= same
        print(f"  {k}: {'OK' if same else 'MISMATCH'}")
    same = np.array_equal(out["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"]); ok &= same
    print(f"  mb_count: {'OK' if same else 'MISMATCH'}")
    same = np.array_equal(out["summary"].cpu().numpy().view(np.uint32).reshape(-1, 2), ref["summary"]); ok &= same
    print(f"  summary: {'OK' if same else 'MISMATCH'}", ref["summary"][:3].tolist())
    bits = np.concatenate([orc.pack_mask16(ref["mask"][i * w * h:(i + 1) * w * h], w, h) for i in range(ns * nf)])
    same = np.array_equal(out["mask_bits"].cpu().numpy().view(np.uint16), bits); ok &= same
    print(f"  mask_bits: {'OK' if same else 'MISMATCH'}")
    for s in range(ns):
        same = np.array_equal(ctx.get_prev_luma(s), prev[s]); ok &= same
    print(f"  state: {'OK' if same else 'MISMATCH'}")
    return ok


def bench(w, h, ns, nf, iters=50):
    ctx = lv.Context(w, h, ns)
    n = ns * nf
    d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
    out = ctx.alloc_outputs(n)
    for _ in range(5):
        ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ctx.convert_diff_batch(d_in, ns, nf, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    px = n * w * h
    print(f"{w}x{h} x{ns}x{nf}: {ms*1e3:.1f} us/launch  {px/ms/1e3:.0f} Mpx/s  {px*5.5/ms/1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ok = True
    for cfg in [(352, 288, 1, 4), (1920, 1080, 1, 3), (1280, 720, 3, 2), (64, 48, 2, 3)]:
        print("parity", cfg)
        ok &= parity(*cfg)
    print("parity uniform"); ok &= parity(352, 288, 1, 2, kind="uniform", tol=25, thr=128)
    print("ALL PARITY OK" if ok else "PARITY FAILED")
    bench(1920, 1080, 1, 64)
    bench(3840, 2160, 1, 32)
    bench(1280, 720, 32, 16)
