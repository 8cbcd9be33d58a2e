"""Multi-GPU plumbing for the KLT path: one process per GPU, sequences sharded across ranks, no data-path collective.

The path shards at frame-batch / sequence granularity only: every fb_tracking!/detect/pyramid call depends
on its own one or two images, and the previous pyramid of a sequence stays on the GPU that owns the sequence.  The only
exchange is a small gather of tracked-keypoint results (counts, or the (N x 2 + N) result arrays when one consumer
needs them all), done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


def shard_sequences(n_sequences: int, world_size: int, rank: int) -> List[int]:
    """Sequence s lives on rank s mod world_size.  Returns this rank's sequence ids, in order."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return [s for s in range(n_sequences) if s % world_size == rank]


def shard_frame_chunks(n_frames: int, world_size: int, rank: int):
    """Contiguous chunk [lo, hi) of frame pairs of ONE stream for this rank (configs 2/5 batched over GPUs).
    Chunk boundaries re-build one extra pyramid: rank r > 0 also needs frame lo-1 as its slot 0."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def max_over_ranks(values: Sequence[float], dist=None, device=None) -> List[float]:
    """Element-wise MAX over ranks of a few timings (device times are taken per rank with CUDA events)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_counts(count: int, dist=None, device=None) -> List[int]:
    """all_gather of one integer per rank (tracked-keypoint counts): the only collective of the path."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(count)]
    import torch
    c = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(c) for _ in range(dist.get_world_size())]
    dist.all_gather(out, c)
    return [int(x.item()) for x in out]


def gather_tracks(points, status, dist=None, device=None):
    """Gather (n x 2 float64 positions, n uint8 status) of every rank to all ranks (fixed n per rank): 34 KB per
    2000-point frame, latency-bound; used only when a single consumer needs every sequence's tracks."""
    import torch
    p = torch.as_tensor(points, dtype=torch.float64, device=device).contiguous()
    s = torch.as_tensor(status, dtype=torch.uint8, device=device).contiguous()
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [p], [s]
    ps = [torch.empty_like(p) for _ in range(dist.get_world_size())]
    ss = [torch.empty_like(s) for _ in range(dist.get_world_size())]
    dist.all_gather(ps, p)
    dist.all_gather(ss, s)
    return ps, ss
