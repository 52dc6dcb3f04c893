"""Generates the golden fixtures in this directory.  Run HERE (the build container), never on the GPU box:

    python tests/golden/make_golden.py

* primitives_cv2.npz -- inputs and outputs of the GENUINE cv2.resize / cv2.remap / cv2.medianBlur
  (OpenCV 4.13, IPP dispatch off so that OpenCV's own open-source resize runs).  Pins oracle/tvl1_oracle.c's
  primitives and the CUDA kernels bit for bit.
* tvl1_pairs.npz -- small frame pairs and the flow computed by oracle/tvl1_cv2ref.py (numpy pointwise steps
  composed with the genuine cv2 primitives; IPP off), with per-level iteration counters.  The real
  cv2.optflow is not installable here, so these are restatement outputs, not OpenCV's:
  the solver as a whole stays "parity unpinned".
"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.tvl1_cv2ref import Cv2ComposedDualTVL1  # noqa: E402
from tee_optical_flow_b200.synth import make_clip  # noqa: E402

OUT = Path(__file__).resolve().parent
f32 = np.float32


def primitives():
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(7)
    d = {}
    for tag, (H, W) in {"a": (41, 53), "b": (64, 80)}.items():
        img = (rng.random((H, W)) * 255).astype(f32)
        d[f"img_{tag}"] = img
        down = cv2.resize(img, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
        d[f"down_{tag}"] = down
        d[f"up_{tag}"] = cv2.resize(down, (W, H), interpolation=cv2.INTER_LINEAR)
        yy, xx = np.mgrid[0:H, 0:W].astype(f32)
        mx = xx + (rng.standard_normal((H, W)) * 2.5).astype(f32)
        my = yy + (rng.standard_normal((H, W)) * 2.5).astype(f32)
        d[f"mapx_{tag}"], d[f"mapy_{tag}"] = mx, my
        d[f"remap_{tag}"] = cv2.remap(img, mx, my, cv2.INTER_CUBIC)
        d[f"median5_{tag}"] = cv2.medianBlur(img, 5)
        d[f"median3_{tag}"] = cv2.medianBlur(img, 3)
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(OUT / "primitives_cv2.npz", **d)


def pairs():
    d = {}
    cases = {
        # name: (H, W, seed, peak_disp, dtype, params)
        "u8_default": (96, 128, 0, 3.0, "u8", {}),
        "u8_fast": (80, 112, 1, 12.0, "u8", {}),
        "f32_default": (72, 96, 2, 3.0, "f32", {}),
        "u8_params": (90, 100, 3, 6.0, "u8", dict(lambda_=0.1, tau=0.2, theta=0.25, nscales=3, warps=3,
                                                  epsilon=0.02, inner=10, outer=3)),
        "u8_tiny_pyramid_stop": (40, 44, 4, 2.0, "u8", dict(nscales=6)),
    }
    for name, (H, W, seed, peak, dt, params) in cases.items():
        fr = make_clip(seed=seed, n_frames=2, H=H, W=W, peak_disp=peak, period=8.0)
        if dt == "f32":
            fr = (fr.astype(f32) / f32(255)).astype(f32)
        for em in (0, 1):
            model = Cv2ComposedDualTVL1(err_mode=em, ipp=False, **params)
            flow = model.calc(fr[0], fr[1])
            d[f"{name}__flow_em{em}"] = flow
            d[f"{name}__counters_em{em}"] = model.last_counters
        d[f"{name}__I0"], d[f"{name}__I1"] = fr[0], fr[1]
        d[f"{name}__params"] = np.array(repr(params))
    np.savez_compressed(OUT / "tvl1_pairs.npz", **d)


if __name__ == "__main__":
    primitives()
    pairs()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
