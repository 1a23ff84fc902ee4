"""Host step (lv_step_host) timing probe: wall clock per synchronous call on pinned buffers, 1080p x 64.
Env LV_HOST_CHUNK_MB = input MiB per chunk; argv[1] = "bgr" to ask for the pictures only (no small outputs)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import livevideo_b200 as lv
w, h, ns, nf = 1920, 1080, 1, 64
only_bgr = len(sys.argv) > 1 and sys.argv[1] == "bgr"
ctx = lv.Context(w, h, ns); n = ns * nf
d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
h_in = lv.pinned(n * ctx.frame_bytes); h_in.array[:] = d_in.cpu().numpy()
h_bgr = lv.pinned(n * ctx.bgr_bytes); h_cnt = lv.pinned(2 * n * ctx.mbs, np.uint16); h_map = lv.pinned(n * ctx.mbs); h_sum = lv.pinned(8 * n, np.uint32)
def step():
    if only_bgr:
        ctx.step_host(h_in.array, ns, nf, bgr=h_bgr.array)
    else:
        ctx.step_host(h_in.array, ns, nf, bgr=h_bgr.array, mb_count=h_cnt.array, mb_map=h_map.array, summary=h_sum.array)
for _ in range(3): step()
ts = []
for _ in range(15):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
ts.sort()
print("chunk", os.environ.get("LV_HOST_CHUNK_MB"), "bgr only" if only_bgr else "all outputs", "best %.2f ms  med %.2f ms  -> %.0f Mpx/s  D2H %.1f GB/s" % (ts[0]*1e3, ts[7]*1e3, n*w*h/ts[7]/1e6, n*ctx.bgr_bytes/ts[7]/1e9))
