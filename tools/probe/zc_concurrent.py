"""Is the host link's ceiling shared between the copy engines and SM-issued (zero-copy) traffic, or do they add up?

A: the zero-copy conversion of P 1080p pictures per call (the kernel reads 1.5 B/px from pinned host memory and writes
   3 B/px to it), looped by one thread.
B: the bare copies of one bench step (200 MB H2D + 400 MB D2H, two streams, copy engines), looped by another thread.
Each alone, then both at once: GB/s of A, of B and the sum.  Run under gpurun (one GPU)."""
import mmap
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import livevideo_b200 as lv  # noqa: E402

W, H, P = 1920, 1080, 8
PX = W * H


def own(nb):
    return np.frombuffer(mmap.mmap(-1, nb), dtype=np.uint8, count=nb)


def main():
    torch.cuda.set_device(0)
    ys, us, vs, ds = ([own(nb) for _ in range(P)] for nb in (PX, PX // 4, PX // 4, 3 * PX))
    for a in ys + us + vs:
        a[:] = 128
    for a in ys + us + vs + ds:
        lv.host_register(a)
    csc = lv.ColorSpaceConversions()
    call = csc.bind_multi(ys, us, vs, ds, W, H)
    lv.set_zero_copy("staged")
    for _ in range(3):
        call()
    n_in, n_out = 200 * 10**6, 400 * 10**6
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def copies():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    copies()
    res = {}

    def loop(fn, nbytes, key, secs, stop):
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < secs and not stop.is_set():
            fn()
            n += 1
        res[key] = n * nbytes / (time.perf_counter() - t0) / 1e9

    zc_bytes = P * PX * 4.5
    ce_bytes = n_in + n_out
    stop = threading.Event()
    loop(call, zc_bytes, "A alone (zero-copy kernel)", 1.0, stop)
    loop(copies, ce_bytes, "B alone (copy engines, H2D + D2H)", 1.0, stop)
    ta = threading.Thread(target=loop, args=(call, zc_bytes, "A with B", 2.0, stop))
    tb = threading.Thread(target=loop, args=(copies, ce_bytes, "B with A", 2.0, stop))
    ta.start(); tb.start(); ta.join(); tb.join()
    for k, v in res.items():
        print(f"{k:38s} {v:7.1f} GB/s")
    print(f"{'A + B together':38s} {res['A with B'] + res['B with A']:7.1f} GB/s")
    # D2H only by the copy engine while the kernel runs (the kernel then competes with one direction only)
    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s2.synchronize()
    res.clear()
    loop(d2h, n_out, "C alone (copy engine, D2H only)", 1.0, stop)
    ta = threading.Thread(target=loop, args=(call, zc_bytes, "A with C", 2.0, stop))
    tb = threading.Thread(target=loop, args=(d2h, n_out, "C with A", 2.0, stop))
    ta.start(); tb.start(); ta.join(); tb.join()
    for k, v in res.items():
        print(f"{k:38s} {v:7.1f} GB/s")
    print(f"{'A + C together':38s} {res['A with C'] + res['C with A']:7.1f} GB/s")
    for a in ys + us + vs + ds:
        lv.host_unregister(a)


if __name__ == "__main__":
    main()
