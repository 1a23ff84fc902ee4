"""Stream-ID sharding across the GPUs of one box (one process per GPU, torch.distributed).

Camera streams are independent units: a stream's only state is its own previous luma
(lib/JM/image.c:1553-1561), and no pixel ever crosses streams.  So the path shards with NO data-path
collective: rank g owns the contiguous block of streams [g*S/G, (g+1)*S/G) and runs the
same fused kernel on its block.  The only exchange is the tiny per-(stream, frame) change summary
{changed_px, changed_mbs}, all-gathered after the step (NCCL on GPUs, gloo in the CPU tests).
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


class ShardedContext:
    """This rank's share of `total_streams` camera streams on its own GPU."""

    def __init__(self, width: int, height: int, total_streams: int, tol: int = 20, mb_thresh: int = 0,
                 rank: int = 0, world: int = 1, device: Optional[int] = None):
        from .api import Context
        self.total_streams, self.rank, self.world = total_streams, rank, world
        self.first_stream, self.local_streams = shard_range(total_streams, world, rank)
        self.device = rank if device is None else device
        self.ctx = Context(width, height, max(self.local_streams, 1), tol, mb_thresh, device=self.device)

    def step(self, i420_local, n_frames, out, cuda_stream=None, gather=True, group=None):
        """Fused kernel over this rank's streams, then (optionally) the summary all-gather.
        Returns the [total_streams, n_frames, 2] summary tensor when gather=True."""
        self.ctx.convert_diff_batch(i420_local, self.local_streams, n_frames, out, cuda_stream=cuda_stream)
        if gather:
            return gather_summaries(out["summary"][: self.local_streams * n_frames * 2], self.total_streams,
                                    n_frames, group=group)
        return None
