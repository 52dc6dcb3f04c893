"""Quick GPU sanity run (development aid): engine vs oracle on a few cases + rough timing."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from oracle import tvl1_oracle as O
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip

def cmp(name, a, b):
    d = np.abs(a - b)
    epe = np.sqrt(((a - b) ** 2).sum(-1))
    print(f"{name}: bit-equal={bool(np.all(a == b))} n_neq={(a != b).sum()} max|d|={d.max():.3e} meanEPE={epe.mean():.3e}", flush=True)

eng = TVL1Engine()
g = np.load(ROOT / "tests/golden/tvl1_pairs.npz")
for case in ["u8_default", "u8_fast", "f32_default", "u8_tiny_pyramid_stop"]:
    if case == "u8_tiny_pyramid_stop":
        eng.setScalesNumber(6)
    f = eng.calc(g[f"{case}__I0"], g[f"{case}__I1"])
    cnt, info = eng.last_counters()
    cmp(case + " vs golden em1", f, g[f"{case}__flow_em1"])
    print("   counters", cnt[0, :, 0].tolist(), "golden", g[f"{case}__counters_em1"][:, 0].tolist(), info)
eng.setScalesNumber(5)

H, W = 600, 800
N = int(sys.argv[1]) if len(sys.argv) > 1 else 9
fr = make_clip(seed=0, n_frames=N, H=H, W=W)
t = time.time(); f32, _ = eng.calc_clip(fr, duplicate_last=True); dt = time.time() - t
cnt, info = eng.last_counters()
print(f"clip {N} frames host path: {dt*1e3:.1f} ms, {(N-1)/dt:.1f} pairs/s; info {info}")
print("   K per level pair0:", cnt[0, :, 0].tolist(), "sumK all pairs", cnt[:, :, 0].sum())
om = O.OracleDualTVL1(err_mode=1)
for i in [0, N - 2]:
    ref = om.calc(fr[i], fr[i + 1])
    cmp(f"600x800 pair {i} vs oracle em1", f32[i], ref)
    print("   counters gpu", cnt[i, :, 0].tolist(), "oracle", om.last_counters[:, 0].tolist())
assert np.array_equal(f32[-1], f32[-2])
frd = torch.from_numpy(fr).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    d32, d16 = eng.calc_clip(frd, want_f16=True)
    torch.cuda.synchronize(); dt = time.time() - t
    cnt, info = eng.last_counters()
    print(f"device path rep{rep}: {dt*1e3:.1f} ms wall, device_ms {info['device_ms']:.1f}, launches {info['solver_launches']}, {(N-1)/dt:.1f} pairs/s")
cmp("device vs host path", d32.cpu().numpy(), f32)
print("f16 equal astype:", bool(np.array_equal(d16.cpu().numpy(), f32.astype(np.float16))))
