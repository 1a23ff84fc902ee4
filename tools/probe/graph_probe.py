"""Does a CUDA graph of the steady-state step (summary memset + fused kernel, x2 for the ping-pong state) beat plain
stream launches?  Captured with torch.cuda.graph around the C-ABI calls.  Run under gpurun."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, livevideo_b200 as lv
w, h, ns, nf = 1920, 1080, 1, 64
ctx = lv.Context(w, h, ns)
d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
out = ctx.alloc_outputs(ns * nf)

def timeit(fn, iters, steps_per_call):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / (iters * steps_per_call) * 1e3)
    return best

plain = timeit(lambda: ctx.convert_diff_batch(d_in, ns, nf, out), 200, 1)
res = {"plain": plain}
for k in (2, 8):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k):
            ctx.convert_diff_batch(d_in, ns, nf, out)
    res[f"graph x{k}"] = timeit(g.replay, 200 // k, k)
plain2 = timeit(lambda: ctx.convert_diff_batch(d_in, ns, nf, out), 200, 1)
res["plain again"] = plain2
for k, v in res.items():
    print(f"{k:12s} {v:7.2f} us/step")
