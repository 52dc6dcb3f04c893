"""Per-phase launch times (CUDA events per launch, teeflow_time_launches): 63 pairs x 600x800 px in lockstep on one
pyramid level -- launch 0 level-init, 1 warp, 2 median, 3.. inner iterations."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tee_optical_flow_b200.synth import make_clip

fr = torch.from_numpy(make_clip(seed=0, n_frames=64, H=600, W=800)).cuda()
res = bench.phase_probe(fr, 0)
print(json.dumps(res, indent=1))
