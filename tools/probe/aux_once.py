"""A few launches of every kernel beside the fused step, for one `ncu --set full` capture (profiles/r01_aux_full_summary.txt)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, livevideo_b200 as lv
w, h, n = 1920, 1080, 64
ctx = lv.Context(w, h, 1)
L = lv.lib()
d_in = lv.gen_synthetic("camera", 1, 0, 0, 1, n, w, h)
out = ctx.alloc_outputs(n, mask_bits=True, mask_bytes=True)
ctx.convert_diff_batch(d_in, 1, n, out)
bkd = d_in[: ctx.frame_bytes].clone()
o420 = torch.empty_like(d_in)
sx, sy = 1920, 1088
pic = torch.randint(0, 1024, (n * (sx * sy * 3 // 2),), dtype=torch.int16, device="cuda")
rgb = torch.randint(0, 256, (n * 3 * w * h,), dtype=torch.uint8, device="cuda")
yv = torch.empty(n * ctx.frame_bytes, dtype=torch.uint8, device="cuda")
px = w * h
ys = torch.stack([d_in[f * ctx.frame_bytes: f * ctx.frame_bytes + px] for f in range(n)]).reshape(-1).contiguous()
ref = torch.roll(ys.reshape(n, px), 1, 0).reshape(-1).contiguous()
for _ in range(3):
    ctx.convert_batch(d_in, n, out["bgr"])
    ctx.diff_batch(ys, ref, n, {"mb_count": out["mb_count"], "mb_map": out["mb_map"], "summary": out["summary"]})
    ctx.write_bks(d_in, bkd, n, o420, tol=25)
    ctx.narrow_pictures(pic, n, sx, sy, 0, 0, o420)
    ctx.object_box_batch(out["mask_bytes"], n, 0, w - 1, 0, h - 1)
    ctx.convert_format(1, rgb, yv, n)
torch.cuda.synchronize()
print("ok")
