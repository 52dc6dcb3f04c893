"""Raw durations of the first solver launches (one level, one group, lockstep): init, warp, median, inner ..."""
import os, sys
from pathlib import Path
os.environ.setdefault("TEEFLOW_GROUPS", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip
fr = torch.from_numpy(make_clip(seed=0, n_frames=64, H=600, W=800)).cuda()
eng = TVL1Engine(device=0, nscales=1, warps=1, max_slots=63)
eng.time_launches(16)
best = None
for _ in range(3):
    eng._calc_clip_device(fr, 1.0, True, False, True)
    t = eng.launch_times_ms()
    best = t if best is None else np.minimum(best, t)
c, info = eng.last_counters()
print("launch us:", " ".join(f"{x*1e3:.0f}" for x in best), info)
