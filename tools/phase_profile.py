"""Phase-pure launches for profiling: a 64-frame 600x800 clip solved on ONE pyramid level with one slot group
(TEEFLOW_GROUPS=1), so the first launches are, for all 63 slots in lockstep: level-init, warp, median, inner, inner...
Run under `ncu --set full -k regex:tvl1_step -c 8` to get per-phase metrics of the strip ops at full resolution."""
import os
import sys
from pathlib import Path
os.environ.setdefault("TEEFLOW_GROUPS", "1")
os.environ.setdefault("TEEFLOW_STEPPED", "1")   # one launch per phase step (the default scheduler is one launch per run)
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fr = torch.from_numpy(make_clip(seed=0, n_frames=n, H=600, W=800)).cuda()
eng = TVL1Engine(device=0, nscales=1, warps=2)
f32, f16 = eng.calc_clip(fr, want_f32=False, want_f16=True)
torch.cuda.synchronize()
c, info = eng.last_counters()
print(info, "inner iterations per pair", c[:, 0, 0].mean())
