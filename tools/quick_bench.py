"""Raw kernel timing of the fused step (device-resident inputs).  Usage: quick_bench.py [WxHxSxT ...]
Env LV_LIBS="a.so,b.so": A/B mode -- every library is timed in its own subprocess, interleaved over several rounds,
and the per-library min / median are printed (run-to-run noise on the shared boxes is ~10 %)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def bench(w, h, ns, nf, iters=30, rounds=3):
    import torch
    import livevideo_b200 as lv
    ctx = lv.Context(w, h, ns)
    n = ns * nf
    d_in = lv.gen_synthetic("camera", 1, 0, 0, ns, nf, w, h)
    out = ctx.alloc_outputs(n)
    for _ in range(5):
        ctx.convert_diff_batch(d_in, ns, nf, out)
    torch.cuda.synchronize()
    res = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ctx.convert_diff_batch(d_in, ns, nf, out)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / iters * 1e3)
    return res

def parse(argv):
    cfgs = [(1920, 1080, 1, 64), (3840, 2160, 1, 32), (1280, 720, 32, 16)]
    if argv:
        cfgs = [tuple(int(x) for x in a.split("x")) for a in argv]
    return cfgs

if __name__ == "__main__":
    if os.environ.get("LV_BENCH_CHILD"):
        for c in parse(sys.argv[1:]):
            print("RES", "x".join(map(str, c)), " ".join(f"{t:.1f}" for t in bench(*c)), flush=True)
        sys.exit(0)
    libs = [x for x in os.environ.get("LV_LIBS", "").split(",") if x] or [None]
    rounds = int(os.environ.get("LV_ROUNDS", "2" if len(libs) > 1 else "1"))
    acc = {}
    for _ in range(rounds):
        for lib in libs:
            env = dict(os.environ, LV_BENCH_CHILD="1")
            if lib:
                env["LV_LIB_OVERRIDE"] = os.path.abspath(lib)
            out = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print(lib, "FAILED", out.stderr[-300:])
                continue
            for line in out.stdout.splitlines():
                if line.startswith("RES"):
                    _, cfg, *ts = line.split()
                    acc.setdefault((lib, cfg), []).extend(float(t) for t in ts)
    for (lib, cfg), ts in acc.items():
        w, h, ns, nf = (int(x) for x in cfg.split("x"))
        px = w * h * ns * nf
        ts.sort()
        best, med = ts[0], ts[len(ts) // 2]
        name = os.path.basename(lib) if lib else os.environ.get("LV_KERNEL", "default")
        print(f"[{name:14s}] {cfg:>16s}: min {best:7.1f} us  med {med:7.1f} us   {px/best:9.0f} Mpx/s  {px*5.5/best/1e3:6.0f} GB/s alg (best)", flush=True)
