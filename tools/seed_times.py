"""Device time of one 64-frame 600x800 clip for several synthetic seeds (data-dependent iteration counts)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip
eng = TVL1Engine(device=0)
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    fr = torch.from_numpy(make_clip(seed=seed, n_frames=64, H=600, W=800)).cuda()
    for _ in range(2):
        eng.calc_clip(fr, want_f32=False, want_f16=True)
    c, info = eng.last_counters()
    print(seed, f"{info['device_ms']:.1f} ms launches {info['solver_launches']} meanK {c[:, :, 0].sum() / 63:.1f} maxK {c[:, :, 0].sum(1).max()}", flush=True)
