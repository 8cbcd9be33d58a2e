"""world_size-2 gloo tests (CPU) of the N > 1 host logic: sharding, max-over-ranks timing, result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import slamklt  # noqa: F401  (registers slam_jl_b200 in sys.modules)
from slam_jl_b200 import dist as skd


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seqs = skd.shard_sequences(7, world, rank)
        t = skd.max_over_ranks([1.0 + rank, 5.0 - rank], dist)
        counts = skd.gather_counts(100 * (rank + 1) + len(seqs), dist)
        pts = np.full((4, 2), float(rank)); st = np.full(4, rank, dtype=np.uint8)
        ps, ss = skd.gather_tracks(pts, st, dist)
        q.put((rank, seqs, t, counts, [float(p[0, 0]) for p in ps], [int(s[0]) for s in ss]))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(60)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]        # sequence s -> rank s mod G
    for r in res:
        assert r[2] == [2.0, 5.0]                                       # max over ranks
        assert r[3] == [104, 203]                                       # gathered counts, rank order
        assert r[4] == [0.0, 1.0] and r[5] == [0, 1]                     # gathered tracks


def test_sharding_covers_everything():
    for n, g in [(64, 8), (65, 8), (5, 8), (64, 1), (7, 2)]:
        allseq = sorted(s for r in range(g) for s in skd.shard_sequences(n, g, r))
        assert allseq == list(range(n))
        spans = [skd.shard_frame_chunks(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        skd.shard_sequences(4, 2, 2)
    assert skd.max_over_ranks([3.0]) == [3.0] and skd.gather_counts(7) == [7]
