"""GPU versions of the reference's per-frame reductions, under the reference's function names and return
conventions (optical_flow/analysis.py).  The heavy part -- cartToPolar, the radial unit grid, the radial /
longitudinal projections, the exact per-frame percentiles and the 1000-bin histograms over (N, H, W) arrays --
runs in libteeflow.so; only the tiny per-frame bookkeeping the reference does in Python (empty-frame
carry-forward, the +1 on the histogram counts) stays on the host.

calc_AV_centroid (connected components + Savitzky-Golay, analysis.py:39-86) lives in masks.py (GPU labelling); its
result is passed in as `centroid_list`.  The downstream waveform pipeline (smoothing, systole / diastole runs, peak
picking) is waveforms.py.
"""
from __future__ import annotations

import numpy as np

from .engine import TVL1Engine


def _carry_forward(vals, counts, first_default):
    """reference rule for frames without non-zero entries (analysis.py:192-202, 249-258): repeat the previous
    frame's value, or use the array extreme when the first frame is empty."""
    out = np.array(vals, copy=True)
    for i in range(len(out)):
        if counts[i] == 0:
            out[i] = out[i - 1] if i > 0 else first_default
    return out


def _carry_forward_hist(freq, counts, first_default):
    out = freq + 1                      # "prevent frequency of 0 for lognorm" (analysis.py:207)
    for i in range(len(out)):
        if counts[i] == 0:
            out[i] = out[i - 1] if i > 0 else first_default
    return out


def calculate_3dhist(engine: TVL1Engine, flow_f16, mask, nframes: int, nbins: int = 1000, percentile: int = 99,
                     centroid_list=None, _res=None):
    """analysis.py:215-286 -> (mag_freq, ang_freq, mag_edges, ang_edges, perc_hi) for masked_arr = flow * mask."""
    H, W = flow_f16.shape[1:3]
    if centroid_list is None:
        centroid_list = np.tile(np.array([[H / 2, W / 2]], np.float64), (nframes, 1))
    res = _res or engine.analyze_clip(flow_f16, mask, np.asarray(centroid_list)[:nframes], nframes, 1, percentile)
    c = res["counts"]
    hi = _carry_forward(res["mag_hi"], c[:, 0], res["mag_max"])
    mag_f, mag_e = engine.analysis_histogram("mag", nframes, res["mag_min"], res["mag_max"], nbins)
    ang_f, ang_e = engine.analysis_histogram("ang", nframes, res["ang_min"], res["ang_max"], nbins)
    one_hot = lambda e, v: np.histogram([v], bins=nbins, range=(e[0], e[-1]))[0] + 1   # first-frame-empty default
    mag = _carry_forward_hist(mag_f, c[:, 0], one_hot(mag_e, res["mag_max"]))
    ang = _carry_forward_hist(ang_f, c[:, 1], one_hot(ang_e, res["ang_max"]))
    return mag, ang, mag_e, ang_e, hi


def calc_bidirectional_hist(engine: TVL1Engine, res: dict, which: str, nframes: int, nbins: int = 1000):
    """analysis.py:166-212 on an analysed quantity ('rad' | 'long') -> (freq, edges, hi_arr, low_arr)."""
    col = {"rad": 2, "long": 3}[which]
    c = res["counts"][:, col]
    hi = _carry_forward(res[f"{which}_hi"], c, res[f"{which}_max"])
    lo = _carry_forward(res[f"{which}_lo"], c, res[f"{which}_min"])
    f, e = engine.analysis_histogram(which, nframes, res[f"{which}_min"], res[f"{which}_max"], nbins)
    freq = _carry_forward_hist(f, c, np.ones(nbins, dtype=np.int64))
    return freq, e, hi, lo


def calculate_3dhist_radlong(engine: TVL1Engine, flow_f16, mask, centroid_list, nframes: int, nbins: int = 1000,
                             perc_lo: int = 1, perc_hi: int = 99) -> dict:
    """analysis.py:289-327 (centroids supplied by the caller): {'radial': (freq, edges[:-1], hi, lo),
    'longitudinal': (...)}."""
    res = engine.analyze_clip(flow_f16, mask, np.asarray(centroid_list)[:nframes], nframes, perc_lo, perc_hi)
    rf, re, rh, rl = calc_bidirectional_hist(engine, res, "rad", nframes, nbins)
    lf, le, lh, ll = calc_bidirectional_hist(engine, res, "long", nframes, nbins)
    return {"radial": (rf, re[:-1], rh, rl), "longitudinal": (lf, le[:-1], lh, ll), "_raw": res}


def angle_mode_per_frame(engine: TVL1Engine, flow_f16, mask, nframes: int) -> np.ndarray:
    """the flow-dependent reduction of AngleDetector.detect (cardiac_cycle_detection.py:100-116)."""
    H, W = flow_f16.shape[1:3]
    cent = np.tile(np.array([[H / 2, W / 2]], np.float64), (nframes, 1))
    return engine.analyze_clip(flow_f16, mask, cent, nframes)["ang_mode"]
