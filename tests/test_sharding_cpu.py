"""CPU tests of the multi-GPU host logic: nchunks semantics, pair-range sharding, and the world_size-2 gather over
gloo (the collective path of SURVEY.md §8e with a stub in place of the GPU pipeline)."""
import os
import socket

import numpy as np
import pytest


def test_chunk_bounds_follow_reference():
    """split = total // nchunks; remainder silently dropped (calculate_optical_flow.py:266-269)"""
    from tee_optical_flow_b200.sharding import chunk_bounds
    assert [chunk_bounds(23, 4, c) for c in range(4)] == [(0, 5), (5, 10), (10, 15), (15, 20)]
    assert chunk_bounds(256, 8, 7) == (224, 256)
    assert chunk_bounds(3, 10, 0) == (0, 0)


@pytest.mark.parametrize("n_pairs,world", [(63, 1), (63, 2), (63, 4), (63, 8), (5, 8), (0, 2), (299, 8)])
def test_pair_ranges_partition(n_pairs, world):
    from tee_optical_flow_b200.sharding import frame_range_for_pairs, pair_range
    covered = []
    for r in range(world):
        lo, hi = pair_range(n_pairs, r, world)
        assert 0 <= lo <= hi <= n_pairs
        covered += list(range(lo, hi))
        f0, f1 = frame_range_for_pairs(lo, hi)
        assert (f1 - f0) == ((hi - lo) + 1 if hi > lo else 0)      # one-frame overlap
    assert covered == list(range(n_pairs))
    sizes = [pair_range(n_pairs, r, world)[1] - pair_range(n_pairs, r, world)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from tee_optical_flow_b200.sharding import WAVEFORM_COLUMNS, allreduce_minmax, process_clip_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.arange(n_frames, dtype=np.uint8)[:, None, None] * np.ones((1, 4, 4), np.uint8)

    def stub(shard, lo, hi):                      # stands in for the GPU pipeline: row p depends on frames p, p+1
        assert shard.shape[0] == hi - lo + 1 and shard[0, 0, 0] == lo
        rows = np.zeros((hi - lo, len(WAVEFORM_COLUMNS)))
        for p in range(lo, hi):
            rows[p - lo] = [shard[p - lo, 0, 0] * 10 + shard[p - lo + 1, 0, 0] + c / 10 for c in range(rows.shape[1])]
        return rows

    table = process_clip_sharded(frames, rank, world, stub)
    lo, hi = allreduce_minmax([float(rank), -rank], [float(rank), 5.0 + rank], world)
    q.put((rank, table, lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [12, 2])
def test_world2_gloo_gather(n_frames):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[p * 10 + (p + 1) + c / 10 for c in range(8)] for p in range(n_frames - 1)])
    for rank, table, lo, hi in got:
        assert table.shape == (n_frames - 1, 8)
        assert np.array_equal(table, want)          # sharded == unsharded, on every rank
        assert lo.tolist() == [0.0, -1.0] and hi.tolist() == [1.0, 6.0]
