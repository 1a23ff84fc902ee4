"""The e2e ceiling, measured: the bare copies of one 1080p x 64 host step (200 MB H2D + 400 MB D2H, pinned, two streams,
no kernel) on ALL N GPUs AT THE SAME TIME -- one process per GPU under torchrun -- against one GPU alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe/link_concurrent.py

Prints, per rank count, the per-GPU and the aggregate GB/s; also the same with a write-combined input buffer
(cudaHostAllocWriteCombined) and with host buffers first-touched by the rank's own thread."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 1 << 20
rt = ctypes.CDLL("libcudart.so.12")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p


def measure(hin, hout, din, dout, active, reps=6):
    """`active`: does this rank copy in this round (N GPUs at once vs one alone)?"""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    if active:
        both()
    barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            both()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    t = torch.tensor([dt if active else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


din = torch.empty(200 * MB, dtype=torch.uint8, device="cuda")
dout = torch.empty(400 * MB, dtype=torch.uint8, device="cuda")
hin = torch.empty(200 * MB, dtype=torch.uint8).pin_memory()
hin.fill_(1)
hout = torch.empty(400 * MB, dtype=torch.uint8).pin_memory()
hout.fill_(2)
res = {}
res["all"] = measure(hin, hout, din, dout, True)
res["rank0_alone"] = measure(hin, hout, din, dout, rank == 0)
hwc, _p = host_alloc(200 * MB, 0x04)          # cudaHostAllocWriteCombined
hwc.fill_(1)
res["all_write_combined_input"] = measure(hwc, hout, din, dout, True)
if rank == 0:
    for k, dt in res.items():
        n = 1 if k == "rank0_alone" else world
        print(f"{k:28s} N={n}: {dt * 1e3:7.2f} ms per 200 MB up + 400 MB down  = {600 * MB / dt / 1e9:6.1f} GB/s per GPU, "
              f"{n * 600 * MB / dt / 1e9:6.1f} GB/s aggregate  ({n * 132.7104 / dt / 1e3:6.1f} Gpx/s ceiling for 1080p x 64)", flush=True)
if world > 1:
    dist.destroy_process_group()
