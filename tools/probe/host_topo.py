"""Host topology + PCIe probe for the host-step pipeline (run under gpurun).  Prints what the e2e number is bounded by:
NUMA layout, the CPUs this container may use, GPU <-> NUMA affinity, and pinned-copy rates alone and concurrent."""
import os, subprocess, sys, time
import torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"

print("cpus allowed:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...", "os.cpu_count", os.cpu_count())
print(sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'"))
print(sh("cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective /sys/fs/cgroup/cpu.max 2>/dev/null"))
print(sh("nvidia-smi topo -m"))
print(sh("grep -E 'MemTotal|MemFree' /sys/devices/system/node/node*/meminfo"))
print(sh("grep -i Mems_allowed_list /proc/self/status; grep -i Cpus_allowed_list /proc/self/status"))
n = torch.cuda.device_count()
print("gpus:", n)
import pynvml
pynvml.nvmlInit()
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        aff = pynvml.nvmlDeviceGetCpuAffinity(h, 8)
    except Exception as e:  # noqa: BLE001
        aff = str(e)
    try:
        mem = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, 0)
    except Exception as e:  # noqa: BLE001
        mem = str(e)
    print(i, "cpu affinity", [hex(x) for x in aff] if isinstance(aff, list) else aff, "mem affinity", mem,
          "pcie gen", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width", pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))

def rate(nbytes, fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9

MB = 1 << 20
for dev in range(min(n, 8)):
    torch.cuda.set_device(dev)
    hin = torch.empty(200 * MB, dtype=torch.uint8).pin_memory(); hin.fill_(1)
    hout = torch.empty(400 * MB, dtype=torch.uint8).pin_memory(); hout.fill_(2)
    din = torch.empty(200 * MB, dtype=torch.uint8, device="cuda")
    dout = torch.empty(400 * MB, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    def both():
        h2d(); d2h()
    t = rate(600 * MB, both)
    print(f"gpu{dev}: H2D alone {rate(200*MB, h2d):.1f} GB/s  D2H alone {rate(400*MB, d2h):.1f} GB/s  "
          f"200MB H2D + 400MB D2H concurrent: {t:.1f} GB/s aggregate, {600*MB/ (t*1e9) *1e3:.2f} ms")
    del hin, hout, din, dout
# host memory bandwidth seen by one thread / all threads (numpy copy)
import numpy as np, threading
a = np.ones(512 * MB, np.uint8); b = np.empty_like(a)
t0 = time.perf_counter(); np.copyto(b, a); dt = time.perf_counter() - t0
print(f"host memcpy 1 thread: {2*a.nbytes/dt/1e9:.1f} GB/s (read+write)")
nt = len(os.sched_getaffinity(0))
parts = np.array_split(np.arange(a.nbytes), nt)
def work(i):
    lo, hi = parts[i][0], parts[i][-1] + 1
    np.copyto(b[lo:hi], a[lo:hi])
ths = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
print(f"host memcpy {nt} threads: {2*a.nbytes/dt/1e9:.1f} GB/s (read+write)")
