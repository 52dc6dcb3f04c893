"""Data points for the other BASELINE configs (parity-test cases, not bench lines):
   config 4 share of one rank: a batch of 64-frame 600x800 clips through one scheduler run (continuous refill);
   config 5: long high-res clip, 1024x1024, 7 scales, 10 warps, WASE on."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from tee_optical_flow_b200.engine import TVL1Engine
from tee_optical_flow_b200.synth import make_clip, make_masks

which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("3", "both"):
    # config 3: saliency on + TV-L1 on the saliency maps + masked radial / longitudinal decomposition of the stored flow
    from tee_optical_flow_b200.sharding import WAVEFORM_COLUMNS
    clip = make_clip(seed=0, n_frames=64, H=600, W=800)
    rgb = torch.from_numpy(np.repeat(clip[..., None], 3, -1)).cuda()
    rv = torch.from_numpy(make_masks(0, 64, 600, 800)["rv"]).cuda()
    cent = np.tile(np.array([[0.77 * 600, 0.5 * 800]]), (62, 1))
    eng = TVL1Engine(device=0)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        sal = eng.compute_saliency(rgb)
        torch.cuda.synchronize(); t1 = time.time()
        _, f16 = eng.calc_clip(sal, want_f32=False, want_f16=True)
        torch.cuda.synchronize(); t2 = time.time()
        res = eng.analyze_clip(f16, rv, cent, 62)
        torch.cuda.synchronize(); t3 = time.time()
    c, info = eng.last_counters()
    print(f"config3: saliency {1e3*(t1-t):.1f} ms + TV-L1 on the saliency maps {1e3*(t2-t1):.1f} ms ({63/(t2-t1):.0f} pairs/s, "
          f"{c[:, :, 0].sum() / 63:.0f} inner iterations/pair) + decomposition {1e3*(t3-t2):.1f} ms", flush=True)
    eng.close()
if which in ("4", "both"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    clips = torch.from_numpy(np.stack([make_clip(seed=s, n_frames=64, H=600, W=800) for s in range(B)])).cuda()
    for slots in (64, 128):
        eng = TVL1Engine(device=0, max_slots=slots)
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            _, f16 = eng.calc_batch(clips)
            torch.cuda.synchronize(); dt = time.time() - t
        c, info = eng.last_counters()
        print(f"config4 share: {B} clips x 63 pairs, slots {slots}: {dt*1e3:.1f} ms -> {B*63/dt:.1f} pairs/s, launches {info['solver_launches']}", flush=True)
        eng.close()
if which in ("5", "both"):
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    fr = make_clip(seed=0, n_frames=N, H=1024, W=1024, peak_disp=4.0, period=40.0)
    masks = make_masks(0, N, 1024, 1024, period=40.0)
    eng = TVL1Engine(device=0, nscales=7, warps=10, max_slots=64)
    eng.set_wase_masks(masks["bkgd"])
    d = torch.from_numpy(fr).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        _, f16 = eng.calc_clip(d, want_f32=False, want_f16=True)
        torch.cuda.synchronize(); dt = time.time() - t
    c, info = eng.last_counters()
    print(f"config5: {N} frames 1024x1024, 7 scales, 10 warps, WASE: {dt*1e3:.1f} ms -> {(N-1)/dt:.1f} pairs/s, "
          f"mean inner iterations/pair {c[:, :, 0].sum() / (N - 1):.0f}, launches {info['solver_launches']}, finite {bool(torch.isfinite(f16).all())}", flush=True)
    print("backgrounds[:3]", eng.last_backgrounds()[:3])
