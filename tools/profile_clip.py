"""Profiling driver: one 64-frame 600x800 clip through the device path, twice (first = warm-up)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fr = torch.from_numpy(make_clip(seed=0, n_frames=n, H=600, W=800)).cuda()
eng = TVL1Engine(device=0)
for r in range(reps):
    f32, f16 = eng.calc_clip(fr, want_f32=False, want_f16=True)
    torch.cuda.synchronize()
    c, info = eng.last_counters()
    print(r, info, "pairs/s", (n - 1) / info["device_ms"] * 1e3, flush=True)
on, cyc, cnt = eng.flow_stats()
if on:
    names = {1: "level_init", 2: "warp", 3: "median", 4: "inner", 5: "final", 6: "wase", 7: "inner2", 9: "wait", 10: "sched"}
    tot = cyc.sum()
    print("flow stats (share of warp cycles, intervals, mean cycles):")
    for i, nme in names.items():
        if cnt[i]:
            print(f"  {nme:10s} {100 * cyc[i] / tot:6.2f} %  n={int(cnt[i]):9d}  mean={cyc[i] / cnt[i]:10.0f}")
