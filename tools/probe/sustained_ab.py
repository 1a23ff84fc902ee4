"""Sustained (power-capped) A/B of the fused step: every variant runs the step back to back for ~1.5 s in its own process; the
time of each sixth is printed (clocks have settled under sw_power_cap by the last third) with the SM clock and board power
sampled through NVML meanwhile.
Usage: sustained_ab.py variant ...   variant = path/to/lib.so | order=0 | order=1 | tile     (run under gpurun)
Env LV_GEOM=WxHxSxT (default 1920x1080x1x64), LV_SEG_LAUNCHES (default: ~0.28 s worth)."""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("LV_CHILD"):
    sys.path.insert(0, ROOT)
    import torch, livevideo_b200 as lv
    import pynvml
    w, h, ns, nf = (int(x) for x in os.environ.get("LV_GEOM", "1920x1080x1x64").split("x"))
    ctx = lv.Context(w, h, ns)
    d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
    out = ctx.alloc_outputs(ns * nf)
    for _ in range(20): ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    per = w * h * ns * nf / 1.17e6 * 1e-6                       # ~ seconds per launch
    n_seg = int(os.environ.get("LV_SEG_LAUNCHES", max(50, int(0.28 / per))))
    pynvml.nvmlInit()
    hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], False
    def sampler():
        while not stop:
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(hdl) / 1000.0))
            time.sleep(0.02)
    th = threading.Thread(target=sampler, daemon=True); th.start()
    seg = []
    t_start = time.perf_counter()
    for s in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_seg): ctx.convert_diff_batch(d_in, ns, nf, out)
        b.record(); seg.append((a, b))
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    stop = True; th.join()
    t = [x.elapsed_time(y) / n_seg * 1e3 for x, y in seg]
    last = [s for s in samples if s[0] > t_start + 0.67 * (t_end - t_start)]
    first = [s for s in samples if s[0] < t_start + 0.15 * (t_end - t_start)]
    med = lambda v: sorted(v)[len(v) // 2] if v else 0
    print("RES", " ".join(f"{v:.1f}" for v in t), f"| first sixth {med([s[1] for s in first])} MHz {med([s[2] for s in first]):.0f} W"
          f" | last third {med([s[1] for s in last])} MHz {med([s[2] for s in last]):.0f} W ({n_seg} launches per segment)")
    sys.exit(0)
for rnd in range(2):
    for var in sys.argv[1:]:
        env = dict(os.environ, LV_CHILD="1")
        if var.startswith("order="):
            env["LV_LOAD_ORDER"] = var[6:]
        elif var != "tile":
            env["LV_LIB_OVERRIDE"] = os.path.abspath(var)
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
        print(f"{os.path.basename(var):12s} us/step per segment:", line[0][4:] if line else r.stderr[-300:], flush=True)
