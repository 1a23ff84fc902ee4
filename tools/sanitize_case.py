"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): all three kernel variants, ragged and partial-band
geometries, mask outputs, conversion-only, difference-only, WriteBks, object box, host step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import livevideo_b200 as lv, lvoracle as orc

for variant in ("tile", "tma"):
    lv.set_kernel_variant(variant)
    for (w, h, ns, nf) in [(352, 288, 2, 2), (96, 20, 1, 2), (1920, 1080, 1, 2)]:
        frames = orc.gen_batch("camera", 1, ns, nf, w, h)
        prev = np.zeros((ns, w * h), np.uint8)
        ref = orc.convert_diff_batch(frames, prev, ns, nf, w, h, 20, 0)
        ctx = lv.Context(w, h, ns)
        out = ctx.alloc_outputs(ns * nf, mask_bits=True, mask_bytes=True)
        d = torch.from_numpy(frames.reshape(-1)).cuda()
        ctx.convert_diff_batch(d, ns, nf, out)
        ctx.convert_batch(d, ns * nf, out["bgr"])
        torch.cuda.synchronize()
        assert np.array_equal(out["bgr"].cpu().numpy(), ref["bgr"]) and np.array_equal(out["mask_bytes"].cpu().numpy(), ref["mask"])
        o2 = ctx.alloc_outputs(ns * nf)
        ctx.reset_prev_luma()
        ctx.convert_diff_batch(d, ns, nf, o2)
        torch.cuda.synchronize()
        assert np.array_equal(o2["mb_count"].cpu().numpy().view(np.uint16), ref["mb_count"])
        ctx.close()
lv.set_kernel_variant("tile")
w, h = 352, 288
ctx = lv.Context(w, h)
fr = orc.gen_batch("camera", 2, 1, 3, w, h)
d = torch.from_numpy(fr.reshape(-1)).cuda()
o = torch.empty_like(d)
bits = torch.zeros(3 * ctx.mask_units, dtype=torch.int16, device="cuda")
ctx.write_bks(d, d[:ctx.frame_bytes].clone(), 3, o, tol=25, obj_bits=bits)
print(ctx.object_box(bits[ctx.mask_units:2 * ctx.mask_units], 0, w - 1, 0, h - 1, mask_is_bits=True))
summ = np.zeros(6, np.uint32)
ctx.reset_prev_luma()
ctx.step_host(fr.reshape(-1).copy(), 1, 3, summary=summ)
print("sanitize case ok", summ.tolist())
