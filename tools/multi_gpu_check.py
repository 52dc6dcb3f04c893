"""torchrun --nproc-per-node N tools/multi_gpu_check.py : one clip sharded by pair range across N GPUs must give the
same flow bits and the same per-frame waveform table as the unsharded run."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.sharding import gpu_shard_pipeline, pair_range, process_clip_sharded
from tee_optical_flow_b200.synth import make_clip, make_masks
from oracle import downstream_ref as R

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, H, W = 18, 120, 160
frames = make_clip(seed=2, n_frames=N, H=H, W=W, peak_disp=4.0, period=9.0)
masks = make_masks(2, N, H, W, period=9.0)
cent = R.calc_av_centroid(masks["av"], N - 1)
eng = TVL1Engine(device=local)
pipe = gpu_shard_pipeline(eng, masks, cent, "rv", out_scale=1.5)
table = process_clip_sharded(frames, rank, world, pipe, device=dev)
lo, hi = pair_range(N - 1, rank, world)
# unsharded reference on this rank
whole = gpu_shard_pipeline(eng, masks, cent, "rv", out_scale=1.5)
ref_table = whole(frames, 0, N - 1)
ok_flow = bool(torch.equal(pipe.flow16, whole.flow16[lo:hi])) if hi > lo else True
ok_tab = bool(np.array_equal(np.nan_to_num(table, nan=-1), np.nan_to_num(ref_table, nan=-1)))
flag = torch.tensor([int(ok_flow), int(ok_tab)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded flow == unsharded: {bool(flag[0])}; gathered waveform table == unsharded: {bool(flag[1])}")
    print("table head:", np.round(table[:2], 4).tolist())
dist.destroy_process_group()
sys.exit(0 if flag.min().item() == 1 else 1)
