import os, sys
sys.path.insert(0, os.getcwd())
import torch, livevideo_b200 as lv
w, h, n = 1920, 1080, 64
ctx = lv.Context(w, h, 1)
d_in = lv.gen_synthetic("camera", 1, 0, 0, 1, n, w, h)
out = ctx.alloc_outputs(n, mask_bits=True, mask_bytes=True)
ctx.convert_diff_batch(d_in, 1, n, out)
for _ in range(3):
    b = ctx.object_box_batch(out["mask_bytes"], n, 0, w - 1, 0, h - 1)
    b2 = ctx.object_box_batch(out["mask_bits"], n, 0, w - 1, 0, h - 1, mask_is_bits=True)
torch.cuda.synchronize()
print(b[:3].tolist(), b2[:3].tolist())
