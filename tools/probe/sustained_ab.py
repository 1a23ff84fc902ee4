"""Sustained (power-capped) A/B of library builds: each library runs the fused step back to back for ~1.5 s in its own
process; the time of the LAST third is reported (clocks have settled under sw_power_cap by then).
Usage: sustained_ab.py lib1.so lib2.so ...   (run under gpurun)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("LV_CHILD"):
    sys.path.insert(0, ROOT)
    import torch, livevideo_b200 as lv
    w, h, ns, nf = 1920, 1080, 1, 64
    ctx = lv.Context(w, h, ns)
    d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
    out = ctx.alloc_outputs(ns * nf)
    for _ in range(20): ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    seg = []
    for s in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2500): ctx.convert_diff_batch(d_in, ns, nf, out)
        b.record(); seg.append((a, b))
    torch.cuda.synchronize()
    t = [x.elapsed_time(y) / 2500 * 1e3 for x, y in seg]
    print("RES", " ".join(f"{v:.1f}" for v in t))
    sys.exit(0)
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, LV_CHILD="1", LV_LIB_OVERRIDE=os.path.abspath(lib))
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
        print(f"{os.path.basename(lib):12s} us/step per 2500-step segment:", line[0][4:] if line else r.stderr[-200:], flush=True)
