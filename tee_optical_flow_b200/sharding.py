"""Multi-GPU sharding of the TV-L1 path: one process per GPU, frame pairs are independent units.

* clips -> ranks: the reference's `nchunks` rule (calculate_optical_flow.py:266-269): split = total // nchunks,
  chunk c owns [c*split, (c+1)*split), remainder dropped; rank r == chunk r.
* one clip -> ranks: contiguous pair ranges [floor(rP/R), floor((r+1)P/R)) with a one-frame overlap of the
  input; flow fields never leave the GPU that produced them.
* the only exchange: the per-frame waveform rows (a few floats per frame) are all-gathered at the end
  (NCCL over NVLink on GPUs, gloo in the CPU tests) -- plus a min/max all-reduce when histograms with
  clip-global edges are wanted (analysis.py:181-182).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

WAVEFORM_COLUMNS = ("mag_hi", "ang_mode", "rad_hi", "rad_lo", "long_hi", "long_lo", "background", "inner_iterations")


def chunk_bounds(total: int, nchunks: int, chunk_index: int) -> Tuple[int, int]:
    split = total // nchunks
    return chunk_index * split, (chunk_index + 1) * split


def pair_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """pairs [lo, hi) of rank `rank`; ranges are contiguous, disjoint and cover [0, n_pairs)."""
    return (rank * n_pairs) // world, ((rank + 1) * n_pairs) // world


def frame_range_for_pairs(lo: int, hi: int) -> Tuple[int, int]:
    """frames needed for pairs [lo, hi): pair p uses frames p and p+1 -> one-frame overlap between ranks."""
    return (lo, hi + 1) if hi > lo else (lo, lo)


def gather_rows(local_rows: np.ndarray, n_total: int, rank: int, world: int, group=None, device=None) -> np.ndarray:
    """All-gather of per-frame rows.  local_rows: (n_local, C) float64 for this rank's frames
    [pair_range(n_total, rank, world)); returns the (n_total, C) table on every rank.  Shards may be ragged (or
    empty): they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    C = local_rows.shape[1] if local_rows.ndim == 2 else len(WAVEFORM_COLUMNS)
    sizes = [pair_range(n_total, r, world)[1] - pair_range(n_total, r, world)[0] for r in range(world)]
    if local_rows.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {sizes[rank]}")
    if world == 1:
        return np.array(local_rows, dtype=np.float64, copy=True)
    cap = max(max(sizes), 1)
    buf = torch.zeros((cap, C), dtype=torch.float64, device=device)
    if sizes[rank]:
        buf[:sizes[rank]] = torch.from_numpy(np.ascontiguousarray(local_rows, np.float64)).to(buf.device)
    out = torch.empty((world, cap, C), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf, group=group) if hasattr(dist, "all_gather_into_tensor") and buf.is_cuda \
        else dist.all_gather(list(out.unbind(0)), buf, group=group)
    host = out.cpu().numpy()
    return np.concatenate([host[r, :sizes[r]] for r in range(world)], axis=0)


def allreduce_minmax(mins: Sequence[float], maxs: Sequence[float], world: int, group=None, device=None):
    """clip-global histogram ranges when a clip is split across ranks (analysis.py:181-182)."""
    import torch
    import torch.distributed as dist
    lo = torch.tensor(list(mins), dtype=torch.float64, device=device)
    hi = torch.tensor(list(maxs), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return lo.cpu().numpy(), hi.cpu().numpy()


def process_clip_sharded(frames_u8: np.ndarray, rank: int, world: int,
                         compute_shard: Callable[[np.ndarray, int, int], np.ndarray], group=None, device=None):
    """Splits one clip by pair range, runs `compute_shard(frames[f0:f1], lo, hi) -> (hi-lo, C) rows` on this rank's
    shard and all-gathers the rows.  `compute_shard` is the GPU pipeline (flow + reductions) in production and a
    stub in the CPU tests."""
    n_pairs = frames_u8.shape[0] - 1
    lo, hi = pair_range(n_pairs, rank, world)
    f0, f1 = frame_range_for_pairs(lo, hi)
    if hi > lo:
        rows = np.asarray(compute_shard(frames_u8[f0:f1], lo, hi), np.float64).reshape(hi - lo, -1)
    else:
        rows = np.zeros((0, len(WAVEFORM_COLUMNS)))
    return gather_rows(rows, n_pairs, rank, world, group, device)


def gpu_shard_pipeline(engine, masks: Optional[dict], centroids: Optional[np.ndarray], label: str = "rv",
                       out_scale: float = 1.0):
    """compute_shard for process_clip_sharded: TV-L1 on the shard's pairs + per-frame reductions of `label`.
    Returns a closure; the fp16 flow of the shard stays on this GPU in closure.flow16."""
    import torch

    def run(shard_frames: np.ndarray, lo: int, hi: int) -> np.ndarray:
        dev = torch.device("cuda", engine.device)
        d = torch.from_numpy(np.ascontiguousarray(shard_frames)).to(dev)
        _, f16 = engine.calc_clip(d, out_scale=out_scale, duplicate_last=False, want_f32=False, want_f16=True)
        run.flow16 = f16
        counters, _ = engine.last_counters()
        n = hi - lo
        rows = np.full((n, len(WAVEFORM_COLUMNS)), np.nan)
        rows[:, 7] = counters[:, :, 0].sum(axis=1)
        rows[:, 6] = engine.last_backgrounds()
        if masks is not None and centroids is not None:
            m = torch.from_numpy(np.ascontiguousarray(masks[label][lo:hi])).to(dev)
            res = engine.analyze_clip(f16, m, np.asarray(centroids)[lo:hi], n)
            for j, k in enumerate(WAVEFORM_COLUMNS[:6]):
                rows[:, j] = res[k]
        return rows

    run.flow16 = None
    return run
