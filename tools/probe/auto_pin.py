"""LV_AUTO_PIN: one YV12_to_RGB24 / FrameDiff_MB call per picture on plain malloc'ed buffers, call by call (run twice: with and
without LV_AUTO_PIN=1 in the environment)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import livevideo_b200 as lv
w, h = (int(x) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (1920, 1080)
px = w * h
rng = np.random.default_rng(1)
y, u, v = (rng.integers(0, 256, n, dtype=np.uint8) for n in (px, px // 4, px // 4))
dst = np.zeros(3 * px, np.uint8)
csc = lv.ColorSpaceConversions()
ts = []
for i in range(12):
    t0 = time.perf_counter(); csc.YV12_to_RGB24(y, u, v, dst, w, h); ts.append((time.perf_counter() - t0) * 1e6)
print(f"LV_AUTO_PIN={os.environ.get('LV_AUTO_PIN')} {w}x{h} YV12_to_RGB24 us per call:", " ".join(f"{t:.0f}" for t in ts), "| mapped:", lv.host_is_mapped(dst))
