#!/usr/bin/env python3
"""BASELINE.json configs[4]: JM-decoded H.264 Annex-B stream -> pinned double-buffered H2D -> fused GPU step.

    python tools/config5_e2e.py [--width 1920 --height 1080 --frames 64 --batch 8] [--check]

The reference's bundled JM 10.2 decoder (oracle/_ref/jm_ldecod, compiled unmodified from /root/reference/lib/JM) is
the host-side producer: it entropy-decodes a synthetic I_PCM Annex-B stream and emits planar I420 frames through
write_out_picture (lib/JM/output.c:362-453) into a pipe.  A reader thread copies them into pinned batch buffers
(two, alternating); every full batch is handed to lv_step_host, which overlaps H2D / kernel / D2H on three CUDA
streams while the decoder keeps filling the other buffer.  JM decodes a few 1080p frames per second, so the chain is
producer-bound; the report separates decoder rate, end-to-end rate, GPU-stage rate and how much of the GPU stage is
hidden behind decoding.  The decoder is test infrastructure (untimed producer in the north-star's sense)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
JM = os.path.join(ROOT, "oracle", "_ref", "jm_ldecod")
JM_INPROC = os.path.join(ROOT, "oracle", "_ref", "jm_inproc")


def run(width, height, n_frames, batch, check=True, seed=1):
    import ipcm_stream
    import lvoracle as orc

    import livevideo_b200 as lv
    assert os.path.exists(JM), "oracle/_ref/jm_ldecod missing: python oracle/jm/build_jm.py (needs /root/reference)"
    fb = width * height * 3 // 2
    tmp = tempfile.mkdtemp(prefix="lv_c5_")
    stream = os.path.join(tmp, "synthetic_ipcm.264")
    t0 = time.perf_counter()
    ipcm_stream.write_stream(stream, (orc.gen("camera", seed, 0, t, width, height) for t in range(n_frames)), width, height)
    t_write = time.perf_counter() - t0

    ctx = lv.Context(width, height, 1, 20, 0)
    n_batches = (n_frames + batch - 1) // batch
    pins = [lv.pinned(batch * fb) for _ in range(2)]
    h_bgr = [lv.pinned(batch * ctx.bgr_bytes) for _ in range(2)]
    cnt = np.zeros((n_frames, ctx.mbs), np.uint16)
    mp = np.zeros((n_frames, ctx.mbs), np.uint8)
    summ = np.zeros((n_frames, 2), np.uint32)
    bgr_crc = []
    gpu_busy = [0.0]
    last_step = [0.0]
    ready = [threading.Semaphore(0), threading.Semaphore(0)]     # batch k%2 filled
    free = [threading.Semaphore(1), threading.Semaphore(1)]      # batch k%2 consumed
    sizes = [0] * n_batches
    err = []

    def producer():
        try:
            p = subprocess.Popen([JM, stream, "-", str(width), str(height)], stdout=subprocess.PIPE, bufsize=0)
            for k in range(n_batches):
                free[k & 1].acquire()
                buf = pins[k & 1].array
                nt = min(batch, n_frames - k * batch)
                view = memoryview(buf)[: nt * fb]
                got = 0
                while got < nt * fb:
                    r = p.stdout.readinto(view[got:])
                    if not r:
                        raise RuntimeError(f"decoder ended early: batch {k}, {got} of {nt * fb} bytes")
                    got += r
                sizes[k] = nt
                ready[k & 1].release()
            p.stdout.close()
            p.wait()
        except Exception as e:  # noqa: BLE001
            err.append(e)
            for s in ready:
                s.release()

    th = threading.Thread(target=producer)
    t_start = time.perf_counter()
    th.start()
    first_batch_in = None
    for k in range(n_batches):
        ready[k & 1].acquire()
        if err:
            raise err[0]
        nt = sizes[k]
        if first_batch_in is None:
            first_batch_in = pins[k & 1].array[: nt * fb].copy() if check else None
        g0 = time.perf_counter()
        f0 = k * batch
        ctx.step_host(pins[k & 1].array, 1, nt, bgr=h_bgr[k & 1].array, mb_count=cnt[f0:f0 + nt].reshape(-1),
                      mb_map=mp[f0:f0 + nt].reshape(-1), summary=summ[f0:f0 + nt].reshape(-1))
        last_step[0] = time.perf_counter() - g0
        gpu_busy[0] += last_step[0]
        if check:
            bgr_crc.append(orc.crc32(h_bgr[k & 1].array[: nt * ctx.bgr_bytes]))
        free[k & 1].release()
    th.join()
    t_total = time.perf_counter() - t_start

    # decoder alone (same stream, output discarded) for the producer rate
    t0 = time.perf_counter()
    subprocess.run([JM, stream, "-", str(width), str(height)], stdout=subprocess.DEVNULL, check=True)
    t_dec = time.perf_counter() - t0

    ok = None
    ref = None
    if check:
        frames = np.stack([orc.gen("camera", seed, 0, t, width, height) for t in range(n_frames)])[None]
        assert np.array_equal(first_batch_in, frames[0, :sizes[0]].reshape(-1)), "JM output != generator frames"
        prev = np.zeros((1, width * height), np.uint8)
        ref = orc.convert_diff_batch(frames, prev, 1, n_frames, width, height, 20, 0, want_mask=False,
                                     n_threads=min(8, os.cpu_count() or 1))
        ok = bool(np.array_equal(cnt.reshape(-1), ref["mb_count"]) and np.array_equal(mp.reshape(-1), ref["mb_map"]) and
                  np.array_equal(summ, ref["summary"]))
        for k in range(n_batches):
            a, b = k * batch, min(n_frames, (k + 1) * batch)
            ok = ok and bgr_crc[k] == orc.crc32(ref["bgr"][a * ctx.bgr_bytes:b * ctx.bgr_bytes])
    px = n_frames * width * height
    rep = {
        "config": f"JM-decoded {width}x{height} Annex-B (synthetic I_PCM, {n_frames} frames) -> pinned x2 -> lv_step_host, batch {batch}",
        "decoder_fps": n_frames / t_dec, "e2e_fps": n_frames / t_total, "e2e_mpx_s": px / t_total / 1e6,
        "gpu_stage_s": gpu_busy[0], "gpu_stage_mpx_s": px / gpu_busy[0] / 1e6,
        # batch k's GPU step runs while the decoder fills the other pinned buffer with batch k+1: only the last
        # batch's step is exposed
        "gpu_stage_fraction_of_wall": gpu_busy[0] / t_total, "gpu_stage_exposed_s": last_step[0],
        "gpu_stage_hidden_behind_decode": 1.0 - last_step[0] / gpu_busy[0] if gpu_busy[0] > 0 else None,
        "decode_s": t_dec, "total_s": t_total, "stream_write_s": t_write, "bit_exact_vs_oracle": ok,
        "gpu_launches": ctx.kernel_launches,
    }
    ctx.close()
    # second feed: the decoder and the GPU library in ONE process (oracle/_ref/jm_inproc: JM unmodified, its write_out_picture
    # replaced at link time; the real unsigned-short planes go to lv_step_host_pictures, narrowed and cropped on the device)
    if os.path.exists(JM_INPROC):
        prefix = os.path.join(tmp, "inproc")
        t0 = time.perf_counter()
        r = subprocess.run([JM_INPROC, stream, str(width), str(height), str(batch), prefix], capture_output=True, timeout=900)
        t_in = time.perf_counter() - t0
        if r.returncode == 0:
            ip = json.loads(r.stdout.decode().strip().splitlines()[-1])
            ok2 = None
            if check:
                ok2 = bool(np.array_equal(np.fromfile(prefix + ".bgr", np.uint8), ref["bgr"]) and
                           np.array_equal(np.fromfile(prefix + ".cnt", np.uint16), ref["mb_count"]) and
                           np.array_equal(np.fromfile(prefix + ".map", np.uint8), ref["mb_map"]) and
                           np.array_equal(np.fromfile(prefix + ".sum", np.uint32).reshape(-1, 2), ref["summary"]))
            rep["in_process"] = {
                "what": "jm_inproc: decoder + liblivevideo.so in one process, 16-bit decoder pictures -> lv_step_host_pictures "
                        "(synchronous per batch: decode and GPU stage alternate)",
                "e2e_fps": ip["frames"] / ip["total_s"], "gpu_stage_s": ip["gpu_stage_s"], "plane_copy_s": ip["plane_copy_s"],
                "gpu_stage_mpx_s": px / ip["gpu_stage_s"] / 1e6 if ip["gpu_stage_s"] > 0 else None,
                "gpu_stage_fraction_of_wall": ip["gpu_stage_s"] / ip["total_s"], "picture": ip["picture"],
                "imgpel_bytes": ip["imgpel_bytes"], "gpu_launches": ip["gpu_launches"], "process_wall_s": t_in,
                "bit_exact_vs_oracle": ok2}
        else:
            rep["in_process"] = {"error": r.stderr.decode()[-300:]}
        for ext in ("bgr", "cnt", "map", "sum"):
            try:
                os.remove(prefix + "." + ext)
            except OSError:
                pass
    try:
        os.remove(stream)
        os.rmdir(tmp)
    except OSError:
        pass
    return rep


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.width, a.height, a.frames, a.batch, check=not a.no_check)))
