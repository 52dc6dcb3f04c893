"""A/B driver: the 64-frame 600x800 clip through the device path `reps` times back to back (sustained load, so the
power cap is in effect like in bench.py), prints the median / best device time.   python tools/ab_clip.py [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 14
fr = torch.from_numpy(make_clip(seed=0, n_frames=64, H=600, W=800)).cuda()
out16 = torch.empty((64, 600, 800, 2), dtype=torch.float16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
eng = TVL1Engine(device=0)
ms = []
for r in range(reps):
    flush.fill_(1)
    eng._calc_clip_device(fr, 1.0, True, False, True, out_f16=out16)
    _, info = eng.last_counters()
    ms.append(info["solver_ms"])
ms = np.array(ms[4:])
print(f"solver ms median {np.median(ms):.3f} best {ms.min():.3f} worst {ms.max():.3f}  -> {63 / (np.median(ms) + 0.75) * 1e3:.1f} pairs/s "
      f"(median + 0.75 ms pyramid)  checksum {float(out16.float().abs().sum()):.6e} spec {info['double_steps']}/{info['double_steps_discarded']}")
