"""Stream-ID sharding across the GPUs of one box (one process per GPU, torch.distributed).

Camera streams are independent units: a stream's only state is its own previous luma
(lib/JM/image.c:1553-1561), and no pixel ever crosses streams.  So the path shards with NO data-path
collective: rank g owns the contiguous block of streams [g*S/G, (g+1)*S/G) and runs the
same fused kernel on its block.  The only exchange is the tiny per-(stream, frame) change summary
{changed_px, changed_mbs}.  On GPUs every rank pushes its block into every peer's table with NVLink stores from a tiny
kernel behind the fused kernel (PeerExchange over the C ABI's lv_exchange_*: no per-step rendezvous, no library
collective on the step); the library all-gather (NCCL on GPUs, gloo in the CPU tests) is the blocking variant and the
cross-check.
"""
from typing import Optional, Tuple


def shard_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """(first_stream, count) of rank `rank`: contiguous blocks, sizes differ by at most one stream."""
    if world < 1 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard arguments")
    first = rank * n_streams // world
    last = (rank + 1) * n_streams // world
    return first, last - first


def owner_of(stream: int, n_streams: int, world: int) -> int:
    """Rank that owns `stream` under shard_range."""
    if not (0 <= stream < n_streams):
        raise ValueError("stream out of range")
    # smallest g with (g+1)*S//G > stream
    g = (stream * world) // n_streams
    while (g + 1) * n_streams // world <= stream:
        g += 1
    while g > 0 and g * n_streams // world > stream:
        g -= 1
    return g


def max_shard(n_streams: int, world: int) -> int:
    return max(shard_range(n_streams, world, r)[1] for r in range(world))


class PendingGather:
    """Handle of an in-flight summary all-gather; result() waits and returns [n_streams, n_frames, 2]."""

    def __init__(self, work, out, n_streams, n_frames, world, cap):
        self.work, self.out = work, out
        self.n_streams, self.n_frames, self.world, self.cap = n_streams, n_frames, world, cap

    def flush(self):
        """Nothing to do: the collective was started when the handle was made (same interface as PendingPeerGather)."""

    def wait_summary(self, stream=None):
        """Nothing to do: on this path out["summary"] is written in-stream by the fused kernel."""

    def wait_table(self, stream=None):
        """Device-side wait for the gathered table (the library collective's own stream semantics)."""
        if self.work is not None:
            self.work.wait()          # NCCL: makes the current stream wait, the host does not block
            self.work = None
        return self.out.reshape(self.world, self.cap, self.n_frames, 2)

    def result(self, copy=True):
        import torch
        if self.work is not None:
            self.work.wait()
            self.work = None
        out = self.out.reshape(self.world, self.cap, self.n_frames, 2)
        parts = [out[r, :shard_range(self.n_streams, self.world, r)[1]] for r in range(self.world)]
        return torch.cat(parts, dim=0)


def gather_summaries_async(local_summary, n_streams: int, n_frames: int, group=None, staging=None, out=None):
    """Start the all-gather and return a PendingGather without making the compute stream wait for it.

    The local summaries are first copied into `staging` (so the caller may overwrite `local_summary` with the
    next step right away); the collective then runs on NCCL's own stream behind that copy, overlapped with
    the next fused kernel.  `staging`/`out` let the caller reuse buffers (ping-pong across steps)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = shard_range(n_streams, world, rank)
    cap = max_shard(n_streams, world)
    local = local_summary.reshape(-1)
    if local.numel() != count * n_frames * 2:
        raise ValueError(f"rank {rank}: expected {count * n_frames * 2} summary words, got {local.numel()}")
    if staging is None:
        staging = torch.zeros(cap * n_frames * 2, dtype=local.dtype, device=local.device)
    if out is None:
        out = torch.empty(world * cap * n_frames * 2, dtype=local.dtype, device=local.device)
    staging[: local.numel()].copy_(local)
    work = dist.all_gather_into_tensor(out, staging, group=group, async_op=True)
    return PendingGather(work, out, n_streams, n_frames, world, cap)


def gather_summaries(local_summary, n_streams: int, n_frames: int, group=None):
    """All-gather the per-(stream, frame) summaries.

    local_summary: int32 tensor [local_streams * n_frames, 2] (or flat) of this rank's block.
    Returns an int32 tensor [n_streams, n_frames, 2] on every rank, on the same device.
    Single process (no initialised process group): returns the local tensor reshaped, no collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_summary.reshape(n_streams, n_frames, 2)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = shard_range(n_streams, world, rank)
    cap = max_shard(n_streams, world)
    local = local_summary.reshape(-1)
    if local.numel() != count * n_frames * 2:
        raise ValueError(f"rank {rank}: expected {count * n_frames * 2} summary words, got {local.numel()}")
    if count < cap:   # pad to the largest block: all_gather needs equal sizes
        pad = torch.zeros((cap - count) * n_frames * 2, dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad])
    out = torch.empty(world * cap * n_frames * 2, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    out = out.reshape(world, cap, n_frames, 2)
    parts = [out[r, :shard_range(n_streams, world, r)[1]] for r in range(world)]
    return torch.cat(parts, dim=0)


class PeerExchange:
    """lv_exchange_*: this rank's ring of summary tables, written into by every peer over NVLink.

    One process per GPU: the 64-byte CUDA IPC handles travel once through torch.distributed (all_gather_object);
    one process driving several ranks (tests): connect_local() wires the raw pointers."""

    def __init__(self, device, rank, world, slot_words, depth=8):
        import ctypes as C

        from . import capi
        self._C, self._capi = C, capi
        self._h = C.c_void_p()
        capi.check(capi.lib().lv_exchange_create(device, rank, world, depth, slot_words, C.byref(self._h)),
                   "lv_exchange_create")
        self.device, self.rank, self.world, self.slot_words, self.depth = device, rank, world, slot_words, depth
        self._views = {}

    def close(self):
        if self._h:
            self._views = {}
            self._capi.lib().lv_exchange_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def handle(self) -> bytes:
        buf = self._C.create_string_buffer(64)
        self._capi.check(self._capi.lib().lv_exchange_handle(self._h, buf), "lv_exchange_handle")
        return buf.raw

    def buffer_ptr(self) -> int:
        return int(self._capi.lib().lv_exchange_buffer(self._h) or 0)

    def connect(self, handles):
        """handles[r] = rank r's handle() (bytes, 64 each)."""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != 64 * self.world:
            raise ValueError("need one 64-byte handle per rank")
        self._capi.check(self._capi.lib().lv_exchange_connect(self._h, blob), "lv_exchange_connect")

    def connect_group(self, group=None):
        """Exchange the IPC handles over torch.distributed and map every peer's buffer."""
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.handle(), group=group)
        self.connect(handles)
        dist.barrier(group=group)        # nobody publishes into a peer that has not finished mapping

    @staticmethod
    def connect_local(exchanges):
        """All ranks live in this process: wire the raw device pointers (no IPC)."""
        import ctypes as C
        ptrs = (C.c_void_p * len(exchanges))(*[x.buffer_ptr() for x in exchanges])
        for x in exchanges:
            x._capi.check(x._capi.lib().lv_exchange_connect_ptrs(x._h, ptrs), "lv_exchange_connect_ptrs")

    def publish(self, step, block, n_words, stream):
        """Push `n_words` int32 words of `block` (a CUDA tensor) as this rank's block of `step` (manual half)."""
        self._capi.check(self._capi.lib().lv_exchange_publish(self._h, step, block.data_ptr(), n_words, stream.cuda_stream),
                         "lv_exchange_publish")

    def collect(self, step, table, stream):
        tp = None if table is None else table.data_ptr()
        self._capi.check(self._capi.lib().lv_exchange_collect(self._h, step, tp, stream.cuda_stream),
                         "lv_exchange_collect")

    def error(self) -> int:
        return int(self._capi.lib().lv_exchange_error(self._h))

    def step(self, ctx, i420, first_stream, n_streams, n_frames, out, stream) -> int:
        """lv_exchange_step: fused kernel + push + lagged collect in one C call; returns the step number."""
        import torch

        from .api import _dev_ptr
        n, d = n_streams * n_frames, ctx.device
        k = self._C.c_uint64()
        self._capi.check(self._capi.lib().lv_exchange_step(
            self._h, ctx._h, _dev_ptr(i420, torch.uint8, n * ctx.frame_bytes, "i420", d, 16), first_stream, n_streams, n_frames,
            _dev_ptr(out.get("bgr"), torch.uint8, n * ctx.bgr_bytes, "bgr", d, 16),
            _dev_ptr(out.get("mask_bits"), torch.int16, n * ctx.mask_units, "mask_bits", d),
            _dev_ptr(out.get("mask_bytes"), torch.uint8, n * ctx.px, "mask_bytes", d, 16),
            _dev_ptr(out.get("mb_count"), torch.int16, n * ctx.mbs, "mb_count", d),
            _dev_ptr(out.get("mb_map"), torch.uint8, n * ctx.mbs, "mb_map", d),
            _dev_ptr(out.get("summary"), torch.int32, n * 2, "summary", d),
            stream.cuda_stream, self._C.byref(k)), "lv_exchange_step")
        return k.value

    def flush(self):
        """Push the newest step's block now (it otherwise travels in front of the next step); asynchronous."""
        self._capi.check(self._capi.lib().lv_exchange_flush(self._h), "lv_exchange_flush")

    def gate(self, stream, hold=False):
        """lv_exchange_gate: device-side rendezvous of all ranks in `stream` (collective).  hold=True: this rank counts as
        arrived only after gate_open() -- call it once the work behind the gate is enqueued."""
        self._capi.check(self._capi.lib().lv_exchange_gate(self._h, stream.cuda_stream, 1 if hold else 0), "lv_exchange_gate")

    def gate_open(self):
        self._capi.check(self._capi.lib().lv_exchange_gate_open(self._h), "lv_exchange_gate_open")

    def wait_summary(self, step, stream):
        """lv_exchange_wait_summary: `stream` waits until the caller's d_summary of `step` has been written."""
        self._capi.check(self._capi.lib().lv_exchange_wait_summary(self._h, step, stream.cuda_stream),
                         "lv_exchange_wait_summary")

    def _view(self, ptr, step):
        """Tensor view of the ring slot at `ptr` (the table of `step`).  Wrapping device memory costs tens of
        microseconds, so the views of ALL ring slots are made at the first call."""
        t = self._views.get(ptr)
        if t is None:
            import torch

            from .api import _DevView
            nbytes = 4 * self.world * self.slot_words
            base = ptr - (step % self.depth) * nbytes
            for s in range(self.depth):
                self._views[base + s * nbytes] = _DevView(base + s * nbytes, nbytes, self.device).tensor().view(
                    torch.int32).reshape(self.world, self.slot_words)
            t = self._views[ptr]
        return t

    def result(self, step, copy=True):
        """int32 CUDA tensor [world, slot_words]: the table of `step` (blocks until every block has arrived).
        copy=False returns a view of the exchange's own ring slot, valid until `depth` further steps are issued."""
        p = self._C.c_void_p()
        self._capi.check(self._capi.lib().lv_exchange_result(self._h, step, self._C.byref(p)), "lv_exchange_result")
        t = self._view(p.value, step)
        return t.clone() if copy else t

    def wait_table(self, step, stream):
        """lv_exchange_wait_table: `stream` waits on the device until the table of `step` is complete; returns the view of
        the ring slot (valid for `depth` further steps) without blocking the host."""
        p = self._C.c_void_p()
        self._capi.check(self._capi.lib().lv_exchange_wait_table(self._h, step, stream.cuda_stream, self._C.byref(p)),
                         "lv_exchange_wait_table")
        return self._view(p.value, step)


class PendingPeerGather:
    """Handle of a step whose summary table arrives through the peer exchange; result() must be asked for before
    RING further steps are issued (its table buffer is a ring slot)."""

    def __init__(self, owner, step, n_frames):
        self.owner, self.step, self.n_frames = owner, step, n_frames

    def flush(self):
        """Make this rank's block of the newest step travel now (one host thread driving several ranks calls this on
        every rank before asking any of them for the result)."""
        self.owner.exchange.flush()

    def wait_summary(self, stream=None):
        """Make `stream` (default: the current one) wait until out["summary"] of this step holds this rank's block: with
        the peer exchange it is written by the push kernel on the exchange's own stream, not in-stream."""
        import torch
        o = self.owner
        o.exchange.wait_summary(self.step, stream or torch.cuda.current_stream(o.device))

    def wait_table(self, stream=None):
        """The device-side result(): `stream` (default: the current one) waits until this step's table is complete and the
        table is returned at once -- its consumer is enqueued behind this call, the host never blocks.  Returns
        [world, cap, n_frames, 2] (rank-padded) as a view of the exchange's ring slot."""
        import torch
        o = self.owner
        cap = max_shard(o.total_streams, o.world)
        tab = o.exchange.wait_table(self.step, stream or torch.cuda.current_stream(o.device))
        return tab[:, : cap * self.n_frames * 2].reshape(o.world, cap, self.n_frames, 2)

    def result(self, copy=True):
        """[total_streams, n_frames, 2].  copy=False: when every rank owns the same number of streams the table is
        returned as a view of the exchange's ring slot (no kernel at all; valid for RING further steps)."""
        import torch
        o = self.owner
        cap = max_shard(o.total_streams, o.world)
        even = cap * o.world == o.total_streams and cap * self.n_frames * 2 == o.exchange.slot_words
        tab = o.exchange.result(self.step, copy=copy or not even)
        if even:
            return tab.reshape(o.total_streams, self.n_frames, 2)
        tab = tab[:, : cap * self.n_frames * 2].reshape(o.world, cap, self.n_frames, 2)
        parts = [tab[r, :shard_range(o.total_streams, o.world, r)[1]] for r in range(o.world)]
        return torch.cat(parts, dim=0)


class ShardedContext:
    """This rank's share of `total_streams` camera streams on its own GPU."""

    RING = 8

    def __init__(self, width: int, height: int, total_streams: int, tol: int = 20, mb_thresh: int = 0,
                 rank: int = 0, world: int = 1, device: Optional[int] = None, exchange=None):
        """exchange: None = library all-gather (torch.distributed) in step_async; a connected PeerExchange, or the
        string "p2p" (create one and connect it over the default process group) = peer-memory exchange."""
        from .api import Context
        self.total_streams, self.rank, self.world = total_streams, rank, world
        self.first_stream, self.local_streams = shard_range(total_streams, world, rank)
        self.device = rank if device is None else device
        self.ctx = Context(width, height, max(self.local_streams, 1), tol, mb_thresh, device=self.device)
        # ring of in-flight gathers: the gather of step k overlaps the kernels of steps k+1 .. k+RING-1 (an 8-rank
        # all-gather of a few hundred bytes is pure latency, ~2 kernel times at 1080p x 64)
        self._pending = [None] * self.RING
        self._staging = [None] * self.RING
        self._gathered = [None] * self.RING
        self._k = 0
        self.exchange = None
        self.exchange_error = None
        if exchange == "p2p":
            self._want_p2p = True
        else:
            self._want_p2p = False
            if exchange is not None:
                self._attach_exchange(exchange)

    # ---- peer-memory exchange -------------------------------------------------------------------------------------
    def _attach_exchange(self, x):
        if x.world != self.world or x.rank != self.rank or x.depth != self.RING:
            raise ValueError("exchange does not match this shard (world, rank, depth)")
        self.exchange = x

    def enable_peer_exchange(self, n_frames, group=None) -> bool:
        """Create this rank's exchange sized for `n_frames`, swap CUDA IPC handles over `group` and map the peers.
        Collective: every rank must call it.  Returns False on EVERY rank -- and leaves the all-gather path in place --
        when any rank could not allocate, export or map (e.g. CUDA IPC not permitted between the processes)."""
        import torch
        import torch.distributed as dist

        def all_ok(ok):
            dev = torch.device("cuda", self.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        cap = max_shard(self.total_streams, self.world)
        words = (cap * n_frames * 2 + 3) & ~3
        x, handle = None, b""
        try:
            x = PeerExchange(self.device, self.rank, self.world, words, depth=self.RING)
            handle = x.handle()
        except Exception as e:  # noqa: BLE001
            self.exchange_error = repr(e)
            x = None
        if not all_ok(x is not None):
            if x is not None:
                x.close()
            return False
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        ok = True
        try:
            x.connect(handles)
        except Exception as e:  # noqa: BLE001
            self.exchange_error = repr(e)
            ok = False
        if not all_ok(ok):          # also the barrier: nobody publishes into a peer that has not finished mapping
            x.close()
            return False
        self._attach_exchange(x)
        return True

    def _make_exchange(self, n_frames, group=None):
        if not self.enable_peer_exchange(n_frames, group):
            self._want_p2p = False      # stay on the library all-gather

    def _step_async_p2p(self, i420_local, n_frames, out):
        import torch
        k = self.exchange.step(self.ctx, i420_local, 0, self.local_streams, n_frames, out,
                               torch.cuda.current_stream(self.device))
        self._k = k + 1
        return PendingPeerGather(self, k, n_frames)

    def step(self, i420_local, n_frames, out, cuda_stream=None, gather=True, group=None):
        """Fused kernel over this rank's streams, then (optionally) the summary all-gather.
        Returns the [total_streams, n_frames, 2] summary tensor when gather=True."""
        self.ctx.convert_diff_batch(i420_local, self.local_streams, n_frames, out, cuda_stream=cuda_stream)
        if gather:
            return gather_summaries(out["summary"][: self.local_streams * n_frames * 2], self.total_streams,
                                    n_frames, group=group)
        return None

    def gate(self, group=None, hold=False):
        """Align the GPUs of all ranks: what is enqueued on the current stream after this call starts within microseconds
        of the same instant on every rank, however far apart the host processes are.  Peer exchange: one tiny kernel
        that signals and waits over NVLink (lv_exchange_gate); otherwise an all-reduce of one word on the stream.
        hold=True (peer exchange only): the gate opens only after every rank has called gate_open(), i.e. when every GPU
        already has the work behind the gate in its queue."""
        import torch
        import torch.distributed as dist
        self._held = False
        if self.exchange is not None:
            self.exchange.gate(torch.cuda.current_stream(self.device), hold=hold)
            self._held = hold
        elif self.world > 1 and dist.is_available() and dist.is_initialized():
            t = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", self.device))
            dist.all_reduce(t, group=group)

    def gate_open(self):
        """Open a gate taken with hold=True (no-op otherwise)."""
        if getattr(self, "_held", False):
            self.exchange.gate_open()
            self._held = False

    def step_async(self, i420_local, n_frames, out, group=None):
        """Fused kernel, then the summary all-gather started WITHOUT blocking the compute stream: the returned
        PendingGather completes while the next steps' kernels run.  At most RING gathers are in flight; the one
        issued RING steps ago is waited for before its buffers are reused.

        NOTE on out["summary"]: with the library all-gather it is valid in-stream right behind the fused kernel (as in
        step()).  With the peer exchange (exchange="p2p") the fused kernel accumulates into the exchange's own ring row
        and out["summary"] is filled by the push kernel on the exchange's side stream: a consumer on the compute
        stream must call handle.wait_summary() first (or use handle.result())."""
        import torch
        import torch.distributed as dist
        if self.exchange is None and self._want_p2p and self.world > 1:
            self._make_exchange(n_frames, group)
        if self.exchange is not None and self.local_streams * n_frames * 2 > self.exchange.slot_words:
            raise ValueError("n_frames larger than the peer exchange was sized for")
        if self.exchange is not None:
            return self._step_async_p2p(i420_local, n_frames, out)
        self.ctx.convert_diff_batch(i420_local, self.local_streams, n_frames, out)
        local = out["summary"][: self.local_streams * n_frames * 2]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return PendingGather(None, local.clone(), self.total_streams, n_frames, 1, self.total_streams)
        i = self._k % self.RING
        self._k += 1
        if self._pending[i] is not None and self._pending[i].work is not None:
            self._pending[i].work.wait()
        cap = max_shard(self.total_streams, self.world)
        need = cap * n_frames * 2
        if self._staging[i] is None or self._staging[i].numel() != need:
            self._staging[i] = torch.zeros(need, dtype=local.dtype, device=local.device)
            self._gathered[i] = torch.empty(self.world * need, dtype=local.dtype, device=local.device)
        self._pending[i] = gather_summaries_async(local, self.total_streams, n_frames, group=group,
                                                  staging=self._staging[i], out=self._gathered[i])
        return self._pending[i]


class SingleProcessShard:
    """lv_shard_*: all GPUs of the box driven from ONE process through the C ABI (NCCL all-gather of the summaries).
    The torchrun deployment (ShardedContext, one process per GPU) is the same partition."""

    def __init__(self, width, height, total_streams, devices, tol=20, mb_thresh=0):
        import ctypes as C

        from . import capi
        self._C, self._capi = C, capi
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        capi.check(capi.lib().lv_shard_create(len(devices), devs, width, height, total_streams, tol, mb_thresh,
                                              C.byref(self._h)), "lv_shard_create")
        self.devices, self.total_streams = list(devices), total_streams
        self.ranges = []
        for r in range(len(devices)):
            d, f, n = C.c_int(), C.c_int(), C.c_int()
            capi.check(capi.lib().lv_shard_range(self._h, r, C.byref(d), C.byref(f), C.byref(n)), "lv_shard_range")
            self.ranges.append((d.value, f.value, n.value))

    def close(self):
        if self._h:
            self._capi.lib().lv_shard_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _ptrs(self, tensors):
        C = self._C
        if tensors is None:
            return None
        return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

    def step(self, i420, n_frames, bgr=None, mb_count=None, mb_map=None):
        """i420 etc.: one CUDA tensor per device (on that device) holding that device's block of streams."""
        self._capi.check(self._capi.lib().lv_shard_step(self._h, self._ptrs(i420), n_frames, self._ptrs(bgr),
                                                        self._ptrs(mb_count), self._ptrs(mb_map)), "lv_shard_step")

    def gather(self, n_frames):
        """[total_streams, n_frames, 2] uint32 on the host (NCCL all-gather across the devices, then one D2H)."""
        import numpy as np
        out = np.zeros((self.total_streams, n_frames, 2), dtype=np.uint32)
        self._capi.check(self._capi.lib().lv_gather_summaries(self._h, out.ctypes.data), "lv_gather_summaries")
        return out

    def sync(self):
        self._capi.check(self._capi.lib().lv_shard_sync(self._h), "lv_shard_sync")
