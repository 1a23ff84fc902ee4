#!/usr/bin/env python3
"""bench.py -- Mpixel/s of the fused YUV420 -> BGR24 + frame-difference step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic frames: a single fused launch
(lv_convert_diff_batch through the C ABI) over this rank's streams, plus -- at N > 1 -- the all-gather of the
tiny per-(stream, frame) change summaries (the only exchange the path has).  Streams are sharded by stream ID,
one process per GPU, no pixel ever crosses GPUs; per-GPU work is fixed as N grows ("weak").

Printed JSON line (rank 0): value = whole-job Mpx/s with inputs resident in HBM; e2e = the same metric through
lv_step_host with pinned HOST buffers (H2D of the frames and D2H of every output inside the timed region);
roofline = algorithmic bytes (5.5 B/px) / mean kernel time vs the measured HBM copy bandwidth; cpu_baseline =
the CPU oracle port (all host threads) on a bounded sample of the same workload, timed in the same run.
`--impl reference` times that CPU path alone (the reference has no other implementation of this path).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_PX = 5.5          # 1.5 in (Y+U+V) + 3 out (BGR) + 1 previous luma   
WORKLOADS = {
    # name: (width, height, streams per GPU, frames per step)      BASELINE.json configs[1..3]
    "1080p_x64": (1920, 1080, 1, 64),
    "4k_x32": (3840, 2160, 1, 32),
    "720p_32streams_x16": (1280, 720, 32, 16),      # one GPU's share of the 256-stream config at 8 GPUs
    "cif_x300": (352, 288, 1, 300),
    # BASELINE.json configs[3]: 256 concurrent 720p streams, 16 frames per step, sharded by stream ID over the ranks
    # (strong scaling: the 256 streams are fixed; `streams per GPU` = 256 / N)
    "720p_256streams_x16": (1280, 720, 256, 16),
}
STRONG = {"720p_256streams_x16"}
DEFAULT_WORKLOAD = "1080p_x64"
METRIC = "Mpixel/s fused YUV420->BGR24 + frame-diff"


def config_of(workload, world):
    w, h, ns, nf = WORKLOADS[workload]
    total = ns if workload in STRONG else ns * world
    return {"workload": workload, "width": w, "height": h, "streams_per_gpu": total / world, "frames_per_step": nf,
            "streams_total": total, "generator": "camera (seed 1)", "tol": 20, "mb_thresh": 0,
            "outputs": "bgr + mb_count + mb_map + summary (bit mask off)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="host-buffer steps to time (default: min(steps, 20))")
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="wall-clock budget of the cpu_baseline leg")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the per-step summaries reach the other ranks (peer-memory pushes or NCCL all-gather)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the `also` object (the other BASELINE.json configs)")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


# ---- clocks during the timed region (NVML; same fields as the nvidia-smi recipe line) -----------------------
class ClockSampler:
    """Samples SM clock + throttle reasons every millisecond on its own thread.  Built and started BEFORE the barrier
    in front of the timed region (nvmlInit with eight processes contending takes milliseconds: it must not sit between
    the barrier and the first event); summary(t0, t1) reports the samples whose host time fell inside the region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period_s=0.001):
        self.samples, self.max_mhz, self.ok, self.period = [], None, False, period_s   # samples: (t, mhz, reason bits)
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        try:
            mhz = int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.samples.append((time.perf_counter(), mhz, r))
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self, t0=None, t1=None):
        sel = [x for x in self.samples if t0 is None or t0 <= x[0] <= t1]
        if not sel and self.samples and t0 is not None:       # a region shorter than the sampling period: nearest samples
            mid = 0.5 * (t0 + t1)
            sel = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:3]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mhz = sorted(x[1] for x in sel)
        bits = 0
        for x in sel:
            bits |= x[2]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(sel)}


# ---- CPU legs (the only place bench.py executes oracle/) ------------------------------------------------------
def cpu_leg(w, h, ns, nf, budget_s, threads, min_reps=1, max_reps=50):
    """Times the CPU oracle port of the path (oracle/lv_oracle.c, multi-threaded, one frame per task)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import lvoracle as orc
    # bounded sample: at most 64 frames of the workload's geometry
    s_ns = min(ns, 4)
    s_nf = min(nf, max(1, 64 // s_ns))
    frames = orc.gen_batch("camera", 1, s_ns, s_nf, w, h)
    prev = np.zeros((s_ns, w * h), dtype=np.uint8)
    px = s_ns * s_nf * w * h
    orc.convert_diff_batch(frames, prev, s_ns, s_nf, w, h, 20, 0, want_mask=False, n_threads=threads)   # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_reps or (time.perf_counter() < t_end and len(times) < max_reps):
        prev[:] = 0
        t0 = time.perf_counter()
        orc.convert_diff_batch(frames, prev, s_ns, s_nf, w, h, 20, 0, want_mask=False, n_threads=threads)
        times.append(time.perf_counter() - t0)
    best, med = min(times), sorted(times)[len(times) // 2]
    return {"mpx_s_median": px / med / 1e6, "mpx_s_best": px / best / 1e6, "reps": len(times), "px": px,
            "sample": f"{s_ns} stream(s) x {s_nf} frames of {w}x{h} ('camera' generator), {len(times)} repetitions",
            "times": times}


def oracle32_leg(w, h):
    """Single-thread timing of the reference's ORIGINAL object code (conversion only), when it can execute."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lvoracle as orc
        if not orc.have_oracle32():
            return None
        fr = orc.gen("uniform", 1, 0, 0, w, h)
        reps = max(1, int(2e8 / (w * h)))
        sec, _ = orc.oracle32_time(fr, w, h, reps)
        return {"mpx_s": reps * w * h / sec / 1e6, "threads": 1, "what": "cscc.lib YV12_to_RGB24 only (no difference)"}
    except Exception:  # noqa: BLE001
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, h, ns, nf = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import lvoracle as orc
    s_ns, s_nf = min(ns, 4), min(nf, max(1, 64 // min(ns, 4)))
    frames = orc.gen_batch("camera", 1, s_ns, s_nf, w, h)
    prev = np.zeros((s_ns, w * h), dtype=np.uint8)
    px = s_ns * s_nf * w * h
    for _ in range(max(args.warmup, 1) if args.warmup < 4 else 3):
        orc.convert_diff_batch(frames, prev, s_ns, s_nf, w, h, 20, 0, want_mask=False, n_threads=threads)
    steps = max(1, min(args.steps, 200))
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        orc.convert_diff_batch(frames, prev, s_ns, s_nf, w, h, 20, 0, want_mask=False, n_threads=threads)
        done += 1
        if time.perf_counter() - t0 > 120:
            break
    dt = time.perf_counter() - t0
    val = px * done / dt / 1e6
    sample = f"each step = {s_ns} stream(s) x {s_nf} frames of {w}x{h}; {done} steps timed"
    cfg = config_of(args.workload, max(args.gpus, 1))
    notes = {"reference_arm": ("CPU port of the reference path (oracle/lv_oracle.c: cscc.lib tables + IsNearlyZero loops), "
                               f"{threads} host threads, one frame per task")}
    line = {"impl": "reference", "metric": METRIC,
            "value": val, "unit": "Mpx/s", "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
            "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg, "notes": notes,
            "cpu_baseline": {"value": val, "unit": "Mpx/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---- the GPU arm ----------------------------------------------------------------------------------------------
class Dist:
    """The few collectives bench.py itself needs (none of them on a timed step)."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def gather(self, x):
        """x (float) of every rank, as a list, on every rank."""
        if self.world == 1:
            return [float(x)]
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        out = self.torch.empty(self.world, device="cuda", dtype=self.torch.float64)
        self.dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.cpu()]

    def max(self, x):
        return max(self.gather(x))


def timed_steps(torch, D, sc, d_in, nf, out, steps, warmup, sampler=None, exchange=True):
    """W warm-up steps of the TIMED path, then exactly K steps between two events on the launching stream.

    N > 1: the timed path is step_async (fused kernel + peer pushes of the summaries); the last step's table is asked
    for inside the region, so every rank's region ends only when the slowest rank's last block has arrived.  After the
    host barrier the GPUS are aligned with a device-side rendezvous in the compute stream (sc.gate) right in front of
    the first event: the region then starts within microseconds of the same instant on every GPU, whatever the skew of
    the host processes.  Returns this rank's region in ms, the host time stamps around it and the launches counted."""
    ctx, world = sc.ctx, D.world
    multi = world > 1 and exchange

    def one():
        if multi:
            return sc.step_async(d_in, nf, out)
        ctx.convert_diff_batch(d_in, sc.local_streams, nf, out)
        return None
    # at least one full turn of the exchange ring plus its collect lag, so ring slots, events and peers are all warm
    w = max(warmup, 3, (sc.RING + sc.RING // 2 + 1) if multi else 0)
    pending = None
    for _ in range(w):
        pending = one()
    if pending is not None:
        pending.result(copy=False)
    # the events exist (and have been recorded once) before the region: at N > 1 the closing record sits on the critical
    # path behind the last table, and a torch event is only created at its first record
    ev0, ev1, ev_last = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for e in (ev0, ev1, ev_last):
        e.record()
    D.barrier()
    launches0 = ctx.kernel_launches
    if multi:
        # held gate: a rank counts as arrived only once its host has the first step enqueued (gate_open below), so no GPU
        # starts its K steps late merely because its host process was the last one out of the barrier
        sc.gate(hold=True)
    # The timed region holds nothing but the K steps: two events on the launching stream, no per-step events (an
    # event pair between back-to-back launches costs ~5 us per step on this kernel: 119 us vs 114 us).
    t0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        pending = one()
        if multi and i == 0:
            sc.gate_open()
    tail_ms = host_tail_ms = 0.0
    if pending is not None:
        ev_last.record()                             # behind the last fused kernel: what follows is the exchange's tail
        # The last exchange is inside the timed region, timed ON THE DEVICE like everything else: the compute stream
        # waits until the table of the last step is complete (every rank's last block has arrived and been collected) and
        # the closing event is recorded behind that wait.  The host then picks the table up (result()); how long that
        # takes after the device is done is a property of the host (reported as tail_host_us), not of the GPUs.
        pending.wait_table()
    ev1.record()
    if pending is not None:
        gathered = pending.result(copy=False)
        th = time.perf_counter()
        assert gathered.shape == (sc.total_streams, nf, 2)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if pending is not None:
        tail_ms = ev_last.elapsed_time(ev1)
        host_tail_ms = (th - t0) * 1e3              # host clock: first launch -> table in hand
    if sampler is not None:
        sampler.sample()
    D.barrier()
    return {"ms": ev0.elapsed_time(ev1), "t0": t0, "t1": t1, "launches": ctx.kernel_launches - launches0, "warmup": w,
            "tail_ms": tail_ms, "host_ms": host_tail_ms}


def single_launch_us(torch, ctx, d_in, ns, nf, out, reps):
    """The fused launch bracketed by its own event pair (median), outside any timed region."""
    iso = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.convert_diff_batch(d_in, ns, nf, out)
        b.record()
        iso.append((a, b))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in iso)[len(iso) // 2] * 1e3


def also_workload(torch, lv, D, name, rank, local_rank, steps, peak):
    """One more BASELINE.json config under the same clock: value, us/step, roofline fraction (device-resident)."""
    from livevideo_b200.shard import ShardedContext
    w, h, ns, nf = WORKLOADS[name]
    strong = name in STRONG
    world = D.world
    total = ns if strong else ns * world
    res = {"workload": name, "scaling": "strong" if strong else "weak", "steps": steps}
    sc = ShardedContext(w, h, total, 20, 0, rank=rank, world=world, device=local_rank,
                        exchange="p2p" if world > 1 else None)
    n = sc.local_streams * nf
    d_in = lv.gen_synthetic("camera", 1, sc.first_stream, 0, sc.local_streams, nf, w, h, device=local_rank)
    out = sc.ctx.alloc_outputs(n)
    r = timed_steps(torch, D, sc, d_in, nf, out, steps, 5)
    ms = D.max(r["ms"])
    px_job = total * nf * w * h
    res.update({"value": px_job * steps / (ms * 1e-3) / 1e6, "unit": "Mpx/s", "us_per_step": ms / steps * 1e3,
                "streams_per_gpu": sc.local_streams,
                "frac": n * w * h * BYTES_PER_PX / (ms / steps * 1e-3) / 1e9 / peak})
    del d_in, out, sc
    torch.cuda.empty_cache()
    if strong and world > 1:
        # the strong-scaling base IN THE SAME RUN: every GPU steps the WHOLE 256-stream job alone (it fits: 17 GB), all
        # GPUs at once (same power and thermal situation as the sharded steps); efficiency = T1 / (N * TN)
        sc1 = ShardedContext(w, h, total, 20, 0, rank=0, world=1, device=local_rank)
        d_in = lv.gen_synthetic("camera", 1, 0, 0, total, nf, w, h, device=local_rank)
        out = sc1.ctx.alloc_outputs(total * nf)
        r1 = timed_steps(torch, D, sc1, d_in, nf, out, max(3, steps // 4), 3, exchange=False)
        us1 = D.max(r1["ms"]) / max(3, steps // 4) * 1e3
        res["single_gpu_us_per_step_same_run"] = us1
        res["efficiency_vs_single_gpu_same_run"] = us1 / (world * res["us_per_step"])
        del d_in, out, sc1
        torch.cuda.empty_cache()
    return res


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import livevideo_b200 as lv
    from livevideo_b200.shard import ShardedContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- livevideo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = Dist(torch, dist, world)
    sampler = ClockSampler(local_rank).start()         # NVML comes up now, far from the timed region
    w, h, ns, nf = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    total_streams = ns if strong else ns * world
    sc = ShardedContext(w, h, total_streams, 20, 0, rank=rank, world=world, device=local_rank,
                        exchange="p2p" if (world > 1 and args.exchange == "p2p") else None)
    ctx = sc.ctx
    ns = sc.local_streams                      # this rank's block of streams
    n = ns * nf
    px_step = n * w * h
    d_in = lv.gen_synthetic("camera", 1, sc.first_stream, 0, ns, nf, w, h, device=local_rank)
    out = ctx.alloc_outputs(n)
    working_set = n * (ctx.frame_bytes + ctx.bgr_bytes)

    if world > 1:
        # every rank must see every stream's summary: check the gathered table once against the local block
        full = sc.step(d_in, nf, out, gather=True)
        torch.cuda.synchronize()
        mine = out["summary"][: n * 2].reshape(ns, nf, 2)
        assert torch.equal(full[sc.first_stream:sc.first_stream + ns], mine), "summary all-gather mismatch"
        # the exchange used in the timed steps (peer-memory pushes by default) must deliver the same table as the
        # blocking NCCL all-gather above: step twice from the same state and compare
        sc.ctx.reset_prev_luma()
        full = sc.step(d_in, nf, out, gather=True).clone()
        sc.ctx.reset_prev_luma()
        got = sc.step_async(d_in, nf, out).result()
        assert torch.equal(got, full), "summary exchange disagrees with the NCCL all-gather"

    r = timed_steps(torch, D, sc, d_in, nf, out, args.steps, args.warmup, sampler=sampler)
    sampler.stop()
    clocks = sampler.summary(r["t0"], r["t1"])
    launches = r["launches"]
    per_rank_ms = D.gather(r["ms"])
    total_ms = max(per_rank_ms)
    per_rank_tail_us = [1e3 * x for x in D.gather(r["tail_ms"])]
    per_rank_host_ms = D.gather(r["host_ms"])
    # N > 1: the same K steps WITHOUT the exchange on every GPU at once -- what each GPU does on its own in this run, so
    # that the cost of the exchange and the spread between the GPUs can be told apart
    per_rank_plain_us = None
    if world > 1:
        rp = timed_steps(torch, D, sc, d_in, nf, out, args.steps, 3, exchange=False)
        per_rank_plain_us = [1e3 * x / args.steps for x in D.gather(rp["ms"])]
    # average duration of one step over the timed region (one fused launch per step; the 2-word-per-frame memset of the
    # summary and, at N > 1, the pushes and the final collect are inside it, so this is an upper bound of the kernel's
    # own time) and the same launch bracketed by its own event pair, outside the timed region
    step_us = total_ms / args.steps * 1e3
    iso_us = D.max(single_launch_us(torch, ctx, d_in, ns, nf, out, min(args.steps, 50)))
    px_job = total_streams * nf * w * h        # pixels of one step over all ranks
    value = px_job * args.steps / (total_ms * 1e-3) / 1e6

    # end to end: pinned host buffers in, every output back in host memory
    e2e = None
    if not args.no_e2e:
        e_steps = args.e2e_steps or max(3, min(args.steps, 20))
        # pinned host buffers for the whole step; only steps beyond 24 GB of host buffers are sampled by streams
        e_ns = ns
        while e_ns > 1 and e_ns * nf * (ctx.frame_bytes + ctx.bgr_bytes) > (24 << 30):
            e_ns //= 2
        e_n = e_ns * nf
        h_in = lv.pinned(e_n * ctx.frame_bytes)
        h_in.array[:] = d_in[: e_n * ctx.frame_bytes].cpu().numpy()
        h_bgr = lv.pinned(e_n * ctx.bgr_bytes)
        h_cnt = lv.pinned(2 * e_n * ctx.mbs, np.uint16)
        h_map = lv.pinned(e_n * ctx.mbs)
        h_sum = lv.pinned(8 * e_n, np.uint32)

        def host_step():
            ctx.step_host(h_in.array, e_ns, nf, bgr=h_bgr.array, mb_count=h_cnt.array, mb_map=h_map.array,
                          summary=h_sum.array)
        for _ in range(3):
            host_step()
        D.barrier()
        l0 = ctx.kernel_launches
        t0 = time.perf_counter()
        for _ in range(e_steps):
            host_step()
        dt_own = time.perf_counter() - t0
        D.barrier()
        e2e_launches = ctx.kernel_launches - l0
        dt = D.max(dt_own)
        h2d = e_n * ctx.frame_bytes
        d2h = e_n * (ctx.bgr_bytes + 3 * ctx.mbs + 8)
        e2e = {"value": e_n * w * h * world * e_steps / dt / 1e6, "unit": "Mpx/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps, "ms_per_step": dt / e_steps * 1e3,
               "gpu_launches": e2e_launches, "api": "lv_step_host (pinned host buffers, 3-slot H2D/kernel/D2H pipeline)",
               "pcie_gbs": (h2d + d2h) * e_steps / dt / 1e9}
        if e_ns != ns:
            e2e["sample"] = f"{e_ns} of this rank's {ns} streams per host step (pinned-memory bound)"
        # The ceiling of that number, MEASURED in the same run: the bare copies of one step (the same pinned buffers, the
        # same bytes, H2D and D2H on two streams, no kernel), on all N GPUs at the same time.
        try:
            t_in, t_out = torch.from_numpy(h_in.array), torch.from_numpy(h_bgr.array)
            d_stage = out["bgr"][: e_n * ctx.bgr_bytes]
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

            def bare():
                with torch.cuda.stream(s_up):
                    d_in[: h2d].copy_(t_in, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    t_out.copy_(d_stage, non_blocking=True)
            bare()
            D.barrier()
            t0 = time.perf_counter()
            for _ in range(max(3, e_steps // 2)):
                bare()
            torch.cuda.synchronize()
            dtb = D.max(time.perf_counter() - t0) / max(3, e_steps // 2)
            D.barrier()
            ceiling = (h2d + e_n * ctx.bgr_bytes) / dtb / 1e9
            e2e["link_ceiling_gbs"] = ceiling
            e2e["link_ceiling"] = {"what": "bare pinned copies of one step (H2D + D2H on two streams, no kernel), all "
                                           f"{world} GPU(s) at once, max over ranks", "ms_per_step": dtb * 1e3,
                                   "pinned_seen_by_torch": bool(t_in.is_pinned()), "aggregate_gbs": ceiling * world,
                                   "mpx_s_ceiling": e_n * w * h * world / dtb / 1e6}
            e2e["frac_of_link_ceiling"] = e2e["pcie_gbs"] / ceiling
            e2e["bound"] = "platform (host link)" if e2e["frac_of_link_ceiling"] > 0.9 else "pipeline"
        except Exception as ex:  # noqa: BLE001
            e2e["link_ceiling"] = {"error": repr(ex)}
        # The literal drop-in entries with host pointers (what the reference's dialogs do, MainDlg.cpp:937-977), on the
        # application's own buffers pinned in place with lv_host_register: one synchronous YV12_to_RGB24 / FrameDiff_MB
        # per picture, and DisplayPic's five conversions through ONE lv_convert_host_multi call.  Rank 0 only.
        if rank == 0:
            try:
                import mmap
                px1 = w * h

                def own(nb):
                    return np.frombuffer(mmap.mmap(-1, nb), dtype=np.uint8, count=nb)
                P = 5
                by, bu, bv, bdst = ([own(nb) for _ in range(P)] for nb in (px1, px1 // 4, px1 // 4, 3 * px1))
                bref, bmask = own(px1), own(px1)
                cnt1 = np.frombuffer(mmap.mmap(-1, 2 * ctx.mbs), dtype=np.uint16, count=ctx.mbs)
                map1 = own(ctx.mbs)
                for i in range(P):
                    fr = h_in.array[(i % e_n) * ctx.frame_bytes:(i % e_n + 1) * ctx.frame_bytes]
                    by[i][:], bu[i][:], bv[i][:] = fr[:px1], fr[px1:px1 + px1 // 4], fr[px1 + px1 // 4:]
                bref[:] = 0
                regs = by + bu + bv + bdst + [bref, bmask, cnt1, map1]
                for a in regs:
                    lv.host_register(a)
                csc = lv.ColorSpaceConversions()
                five = csc.bind_multi(by, bu, bv, bdst, w, h)      # pointer arrays built once, like the dialog's members
                calls = 32

                def drop_in(mode):
                    lv.set_zero_copy(mode)
                    for _ in range(3):
                        csc.YV12_to_RGB24(by[0], bu[0], bv[0], bdst[0], w, h)
                        lv.FrameDiff_MB(by[0], bref, w, h, 20, 0, mask=bmask, mb_count=cnt1, mb_map=map1)
                        five()
                    t0 = time.perf_counter()
                    for _ in range(calls):
                        csc.YV12_to_RGB24(by[0], bu[0], bv[0], bdst[0], w, h)
                    t1 = time.perf_counter()
                    for _ in range(calls):
                        lv.FrameDiff_MB(by[0], bref, w, h, 20, 0, mask=bmask, mb_count=cnt1, mb_map=map1)
                    t2 = time.perf_counter()
                    for _ in range(calls):
                        five()
                    t3 = time.perf_counter()
                    return {"YV12_to_RGB24_us": (t1 - t0) / calls * 1e6, "FrameDiff_MB_us": (t2 - t1) / calls * 1e6,
                            "both_mpx_s": px1 * calls / (t2 - t0) / 1e6,
                            "YV12_to_RGB24_mpx_s": px1 * calls / (t1 - t0) / 1e6,
                            "five_pictures_one_call_us": (t3 - t2) / calls * 1e6,
                            "five_pictures_one_call_mpx_s": P * px1 * calls / (t3 - t2) / 1e6}
                zc_before = lv.get_zero_copy()
                staged = drop_in("off")
                e2e["drop_in_calls"] = {
                    "what": "host pointers pinned + mapped in place (lv_host_register); single caller thread; the kernel "
                            "reads and writes the host buffers itself (zero copy, the default for such buffers)",
                    **drop_in("staged"),
                    "staged_through_device_buffers": staged}
                lv.set_zero_copy(zc_before)
                for a in regs:
                    lv.host_unregister(a)
            except Exception as ex:  # noqa: BLE001
                e2e["drop_in_calls"] = {"error": repr(ex)}
        # sanity: the host path returned the same summaries as the device path
        a = h_sum.array.view(np.uint32).reshape(e_ns, nf, 2)[:, 1:]
        ctx.convert_diff_batch(d_in, ns, nf, out)
        b = out["summary"].cpu().numpy().view(np.uint32).reshape(ns, nf, 2)[:e_ns, 1:]
        assert np.array_equal(a, b), "host step and device step disagree"     # t = 0 depends on the carried state
        del h_in, h_bgr, h_cnt, h_map, h_sum

    # the other BASELINE.json configs under the same clock (device-resident; a few ms each)
    peak, peak_src = measured_peak()
    also = None
    if not args.no_also and args.workload == DEFAULT_WORKLOAD:
        del d_in, out
        torch.cuda.empty_cache()
        also = {}
        names = ["4k_x32", "720p_32streams_x16", "720p_256streams_x16"] if world == 1 else ["720p_256streams_x16"]
        for name in names:
            try:
                also[name] = also_workload(torch, lv, D, name, rank, local_rank, 50 if name != "720p_256streams_x16" else 20, peak)
            except Exception as ex:  # noqa: BLE001
                also[name] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    achieved = px_step * BYTES_PER_PX / (step_us * 1e-6) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": r["warmup"], "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": config_of(args.workload, world),
        "notes": {"l2": f"no flush: the step streams {working_set / 1e6:.0f} MB per GPU (> 126 MB L2) between two uses "
                        "of any byte"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(args.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_px": BYTES_PER_PX, "px_per_launch": px_step,
                     "step_us": step_us, "kernel_us": step_us if world == 1 else iso_us,
                     "kernel_us_single_launch_events": iso_us,
                     "what": "achieved = algorithmic bytes per launch / step_us (timed region / steps: one fused launch "
                             "per step" + ("" if world == 1 else ", plus the summary pushes and the final collect") + ")",
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "per_rank_ms": {"min": min(per_rank_ms), "max": max(per_rank_ms), "all": per_rank_ms},
    }
    if world > 1:
        line["exchange_cost"] = {
            "plain_step_us_per_rank": per_rank_plain_us, "tail_us_per_rank": per_rank_tail_us,
            "host_first_launch_to_table_in_hand_ms_per_rank": per_rank_host_ms,
            "what": "plain = the same K steps without the exchange, all GPUs at once (no gate, no pushes, no final collect); "
                    "tail = last fused kernel's end -> table of the last step complete on the device (inside the timed "
                    "region, once); host_* = host wall clock from the first launch to result() returning the table",
            "us_per_step_over_slowest_gpu_alone": step_us - max(per_rank_plain_us)}
    if world > 1:
        line["notes"]["summary_exchange"] = ("peer-memory pushes over NVLink (lv_exchange_*), no per-step collective; GPUs "
                                             "aligned by lv_exchange_gate in front of the first event"
                                             if sc.exchange is not None else "NCCL all_gather_into_tensor per step")
        if sc.exchange is None and args.exchange == "p2p":
            line["notes"]["summary_exchange"] += f" (peer exchange unavailable: {sc.exchange_error})"
    if e2e is not None:
        line["e2e"] = e2e
    if also is not None:
        line["also"] = also
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        c = cpu_leg(w, h, ns, nf, args.cpu_seconds, threads)
        c1 = cpu_leg(w, h, 1, min(nf, 8), min(args.cpu_seconds, 2.0), 1)
        line["cpu_baseline"] = {"value": c["mpx_s_median"], "unit": "Mpx/s", "cores": threads, "kind": "port",
                                "sample": c["sample"], "best": c["mpx_s_best"],
                                "single_thread_mpx_s": c1["mpx_s_median"],
                                "reference_object_code_single_thread": oracle32_leg(w, h)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
