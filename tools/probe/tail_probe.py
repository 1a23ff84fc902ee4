"""Where does the exchange's tail go?  torchrun -N 2: 20 sharded steps, then the table of the last step; GPU time stamps of
the push / collect kernels of the last step (lv_exchange_trace) against the end of the last fused kernel, and host time
stamps around lv_exchange_result."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import livevideo_b200 as lv  # noqa: E402
from livevideo_b200.shard import ShardedContext  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h, nf = 1920, 1080, 64
sc = ShardedContext(w, h, world, 20, 0, rank=rank, world=world, device=local, exchange="p2p")
d_in = lv.gen_synthetic("camera", 1, rank, 0, 1, nf, w, h, device=local)
out = sc.ctx.alloc_outputs(nf)
for _ in range(14):
    p = sc.step_async(d_in, nf, out)
p.result(copy=False)
tr = lv.lib().lv_exchange_trace(sc.exchange._h)
for trial in range(6):
    dist.barrier()
    torch.cuda.synchronize()
    sc.gate(hold=True)
    ev0, evl, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ev0.record()
    for i in range(20):
        p = sc.step_async(d_in, nf, out)
        if i == 0:
            sc.gate_open()
    evl.record()
    ta = time.perf_counter()
    g = p.result(copy=False)
    tb = time.perf_counter()
    ev1.record()
    torch.cuda.synchronize()
    tc = time.perf_counter()
    t = [tr[i] for i in range(5)]
    print(f"rank {rank} trial {trial}: region {ev0.elapsed_time(ev1) * 1e3:8.1f} us  tail(evl->ev1) {evl.elapsed_time(ev1) * 1e3:6.1f} us | "
          f"fused end -> push entry {(t[0] - t[4]) / 1e3:6.1f}  push {(t[1] - t[0]) / 1e3:5.1f}  push exit -> collect entry {(t[2] - t[1]) / 1e3:5.1f}  "
          f"collect {(t[3] - t[2]) / 1e3:5.1f}  => fused end -> collect exit {(t[3] - t[4]) / 1e3:6.1f} us | host: result() {1e6 * (tb - ta):7.1f} us "
          f"(includes waiting for the GPU), after result -> synchronized {1e6 * (tc - tb):5.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
