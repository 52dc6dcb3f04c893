"""Deterministic synthetic TEE-like clips and masks: the reference's sample DICOM
(test_data/dcm/stanford_RVIO_49_2.dcm) is a missing blob, so every measurement and parity test runs on these.

Only numpy + scipy.ndimage; no GPU, no oracle.  Same seed -> same bytes on every box.
"""
from __future__ import annotations

import numpy as np


def sector_mask(H: int, W: int) -> np.ndarray:
    """Ultrasound sector (black outside, as in TEE): apex above the image centre, +-42 degrees."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    ay, ax = -0.04 * H, 0.5 * W
    r = np.hypot(yy - ay, xx - ax)
    ang = np.arctan2(xx - ax, yy - ay)
    return (np.abs(ang) < np.deg2rad(42.0)) & (r > 0.09 * H) & (r < 1.0 * H)


def displacement_field(H: int, W: int, t: float, rng_state: dict):
    """Smooth periodic displacement (sum of 3 low-frequency sinusoids), peak ~= rng_state['peak'] px."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    dx = np.zeros((H, W))
    dy = np.zeros((H, W))
    for k in range(3):
        fx, fy, ph, ps, ax_, ay_ = rng_state["modes"][k]
        spatial = np.sin(2 * np.pi * (fx * xx / W + fy * yy / H) + ph)
        temporal = np.sin(2 * np.pi * t / rng_state["period"] + ps)
        dx += ax_ * spatial * temporal
        dy += ay_ * spatial * temporal
    return dx, dy


def make_clip(seed: int = 0, n_frames: int = 64, H: int = 600, W: int = 800, peak_disp: float = 3.0,
              period: float = 32.0, return_truth: bool = False):
    """(n_frames, H, W) uint8 speckle-textured sector clip advected by a smooth periodic field.

    peak_disp=3, period=32 gives <= ~0.6 px/frame ("slow" clip S of SURVEY.md §8d); peak_disp=15 gives the
    "fast" clip (~3 px/frame) that needs more inner iterations.
    """
    from scipy.ndimage import gaussian_filter, map_coordinates

    rng = np.random.default_rng(seed)
    tex = gaussian_filter(rng.standard_normal((H, W)), 2.0)
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    speckle = gaussian_filter(rng.rayleigh(1.0, (H, W)), 0.7)
    img = tex * speckle
    lo, hi = np.percentile(img, [0.5, 99.5])
    img = np.clip((img - lo) / (hi - lo), 0.0, 1.0) * 255.0

    modes = []
    amp = peak_disp / 3.0
    for _ in range(3):
        modes.append((rng.uniform(0.5, 2.0), rng.uniform(0.5, 2.0), rng.uniform(0, 2 * np.pi),
                      rng.uniform(0, 2 * np.pi), amp * rng.uniform(0.6, 1.0) * rng.choice([-1, 1]),
                      amp * rng.uniform(0.6, 1.0) * rng.choice([-1, 1])))
    state = {"modes": modes, "period": period, "peak": peak_disp}

    sector = sector_mask(H, W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    frames = np.empty((n_frames, H, W), np.uint8)
    truth = []
    for t in range(n_frames):
        dx, dy = displacement_field(H, W, float(t), state)
        warped = map_coordinates(img, [yy + dy, xx + dx], order=3, mode="reflect")
        frames[t] = np.clip(np.rint(warped * sector), 0, 255).astype(np.uint8)
        if return_truth:
            truth.append((dx, dy))
    if return_truth:
        return frames, truth
    return frames


def make_masks(seed: int, n_frames: int, H: int, W: int, period: float = 32.0) -> dict:
    """Synthetic stand-ins for the SAM masks the reference feeds the analysis (RVIO_2class):
    'rv' = moving ellipse annulus, 'av' = 40x40 blob at the annulus base, 'bkgd' = complement of the union.
    Every value is (N, H, W, 2) bool, like clean_mask's output (calculate_optical_flow.py:113-182)."""
    rng = np.random.default_rng(seed + 10_000)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    cy0, cx0 = 0.55 * H + rng.uniform(-5, 5), 0.5 * W + rng.uniform(-5, 5)
    ry, rx = 0.22 * H, 0.18 * W
    rv = np.zeros((n_frames, H, W), bool)
    av = np.zeros((n_frames, H, W), bool)
    for t in range(n_frames):
        ph = 2 * np.pi * t / period
        cy = cy0 + 3.0 * np.sin(ph)
        cx = cx0 + 2.0 * np.cos(ph)
        e = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        rv[t] = (e < 1.0) & (e > 0.45)
        by, bx = int(round(cy + ry)), int(round(cx))
        half = max(2, min(20, H // 8, W // 8))
        y0, y1 = max(0, by - half), min(H, by + half)
        x0, x1 = max(0, bx - half), min(W, bx + half)
        av[t, y0:y1, x0:x1] = True
    bkgd = ~(rv | av)
    rep = lambda m: np.repeat(m[..., None], 2, axis=-1)
    return {"rv": rep(rv), "av": rep(av), "bkgd": rep(bkgd)}
