"""What the application sees: one call of the drop-in entry YV12_to_RGB24 (Convert.h:26) / FrameDiff_MB per picture,
host pointers, synchronous -- with the app's buffers as they are (pageable) and after lv_host_register -- next to the
CPU path (the oracle port, one thread, and the reference's original object code).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import livevideo_b200 as lv
import lvoracle as orc


def t_call(fn, reps):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]


def reg(a):
    try:
        lv.host_register(a)
        return True
    except lv.LiveVideoError:      # small arrays from the malloc heap can share a page with a registered neighbour
        return False


def big(n, dtype=np.uint8):
    """Page-aligned array of its own (what a large malloc / VirtualAlloc gives the application)."""
    import mmap
    nbytes = n * np.dtype(dtype).itemsize
    return np.frombuffer(mmap.mmap(-1, max(nbytes, 1)), dtype=dtype, count=n)


csc = lv.ColorSpaceConversions()
print("registered = pinned + mapped in place (lv_host_register); staged = H2D copies -> kernel -> D2H copy (zero copy off); "
      "zc plain / zc = the kernel addresses the host buffers itself (tile-kernel stores / stores re-dealt to 512 B runs)")
print(f"{'size':>10s} {'entry':>14s} {'pageable us':>12s} {'staged us':>10s} {'zc plain us':>12s} {'zc us':>8s} {'cpu 1 thread us':>16s} {'obj code us':>12s}   Mpx/s (zc) / speed-up")
for (w, h) in [(352, 288), (1280, 720), (1920, 1080), (3840, 2160)]:
    px = w * h
    fr = orc.gen("camera", 1, 0, 1, w, h)
    y, u, v, ref, dst = big(px), big(px // 4), big(px // 4), big(px), big(3 * px)
    y[:], u[:], v[:] = fr[:px], fr[px:px + px // 4], fr[px + px // 4:]
    ref[:] = orc.gen("camera", 1, 0, 0, w, h)[:px]
    reps = 200 if px < 1e6 else 60
    pg = t_call(lambda: csc.YV12_to_RGB24(y, u, v, dst, w, h), reps)
    want = dst.copy()
    for a in (y, u, v, dst):
        assert reg(a)
    rgs = {}
    for mode in ("off", "plain", "staged"):
        lv.set_zero_copy(mode)
        dst[:] = 0
        rgs[mode] = t_call(lambda: csc.YV12_to_RGB24(y, u, v, dst, w, h), reps)
        assert np.array_equal(dst, want), mode
    rg = rgs["staged"]
    prev = np.zeros((1, px), np.uint8)
    cpu = t_call(lambda: orc.convert_diff_batch(fr.reshape(1, 1, -1), prev, 1, 1, w, h, 20, 0, want_mask=False, n_threads=1), max(5, reps // 10))
    o32 = None
    if orc.have_oracle32():
        sec, _ = orc.oracle32_time(fr, w, h, 5)
        o32 = sec / 5
    o32s = f"{o32 * 1e6:12.0f}" if o32 else " " * 12
    print(f"{w:5d}x{h:<4d} {'YV12_to_RGB24':>14s} {pg * 1e6:12.0f} {rgs['off'] * 1e6:10.0f} {rgs['plain'] * 1e6:12.0f} {rg * 1e6:8.0f} {cpu * 1e6:16.0f} {o32s}   {px / rg / 1e6:8.0f} / {cpu / rg:5.1f}x")
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    mask, cnt, mp = big(px), big(mbs, np.uint16), big(mbs)
    lv.set_zero_copy("off")
    for a in (y, u, v, dst):
        lv.host_unregister(a)
    pgd = t_call(lambda: lv.FrameDiff_MB(y, ref, w, h, 20, 0, mask=mask, mb_count=cnt, mb_map=mp), reps)
    wantd = (mask.copy(), cnt.copy(), mp.copy())
    assert reg(y) and reg(ref) and reg(mask) and reg(cnt) and reg(mp)
    rgd = {}
    for mode in ("off", "staged"):
        lv.set_zero_copy(mode)
        mask[:] = 9
        rgd[mode] = t_call(lambda: lv.FrameDiff_MB(y, ref, w, h, 20, 0, mask=mask, mb_count=cnt, mb_map=mp), reps)
        assert all(np.array_equal(a, b) for a, b in zip((mask, cnt, mp), wantd)), mode
    print(f"{w:5d}x{h:<4d} {'FrameDiff_MB':>14s} {pgd * 1e6:12.0f} {rgd['off'] * 1e6:10.0f} {'':>12s} {rgd['staged'] * 1e6:8.0f}")
    for a in (y, ref, mask, cnt, mp):
        lv.host_unregister(a)
