import sys, os, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'oracle')
    import numpy as np, torch
    import livevideo_b200 as lv, lvoracle as orc
    w, h = int(sys.argv[1]), int(sys.argv[2])
    fr = orc.gen("uniform", 1, 0, 0, w, h)
    ctx = lv.Context(w, h)
    d_out = torch.empty(ctx.bgr_bytes, dtype=torch.uint8, device="cuda")
    ctx.convert_batch(torch.from_numpy(fr).cuda(), 1, d_out)
    torch.cuda.synchronize()
    print("ok" if np.array_equal(d_out.cpu().numpy(), orc.convert_frame(fr, w, h)) else "MISMATCH")
else:
    for (w, h) in [(32, 16), (32, 18), (128, 18), (128, 24), (128, 20), (256, 10), (32, 2), (128, 16), (64, 18), (160, 8)]:
        r = subprocess.run([sys.executable, __file__, str(w), str(h)], capture_output=True, text=True)
        print(w, h, (r.stdout.strip() or r.stderr.strip().splitlines()[-1][:100]))
