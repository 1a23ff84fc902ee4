"""lv_convert_host_multi: fixed cost and per-picture cost (registered host buffers, direct ctypes calls).
Usage (under gpurun): python tools/probe/multi_probe.py [W H]"""
import ctypes as C
import mmap
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import livevideo_b200 as lv  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
px = w * h
L = lv.lib()


def own(nb):
    return np.frombuffer(mmap.mmap(-1, nb), dtype=np.uint8, count=nb)


NMAX = 40
ys, us, vs, ds = ([own(nb) for _ in range(NMAX)] for nb in (px, px // 4, px // 4, 3 * px))
for a in ys + us + vs + ds:
    a[:] = 7
    lv.host_register(a)


def arr(xs, n):
    return (C.c_void_p * n)(*[a.ctypes.data for a in xs[:n]])


for n in (1, 2, 5, 10, 20, 40):
    A = (arr(ys, n), arr(us, n), arr(vs, n), arr(ds, n))
    for _ in range(3):
        assert L.lv_convert_host_multi(*A, n, w, h) == 0
    reps = max(3, 100 // n)
    t0 = time.perf_counter()
    for _ in range(reps):
        L.lv_convert_host_multi(*A, n, w, h)
    dt = (time.perf_counter() - t0) / reps
    print(f"{w}x{h} n={n:3d}: {dt * 1e6:8.1f} us per call, {dt / n * 1e6:7.1f} us per picture, {n * px / dt / 1e9:6.2f} Gpx/s "
          f"({n * 4.5 * px / dt / 1e9:5.1f} GB/s over the link)", flush=True)
t0 = time.perf_counter()
for _ in range(20):
    L.YV12_to_RGB24(ys[0].ctypes.data, us[0].ctypes.data, vs[0].ctypes.data, ds[0].ctypes.data, w, h)
print(f"single synchronous YV12_to_RGB24: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us")
