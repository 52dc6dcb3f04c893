"""Downstream waveform pipeline of the reference, in the product: per-frame waveforms (from the GPU reductions) ->
smoothing -> systole / diastole intervals -> systolic / e' / l' / a' peak frame indices.

Reference code this follows (same function names, arguments, return layouts):
    AngleDetector.detect             optical_flow/cardiac_cycle_detection.py:87-143   (after the per-frame angle mode,
                                     which teeflow_analyze_clip computes on the GPU)
    find_start_stop                  optical_flow/optical_flow_utils.py:40-49
    PeakDetector, calculate_radlong_peaks, calculate_single_peaks      optical_flow/peak_detection.py:16-375
and the two third-party helpers those call, neither of which is installable in this image -- written from their
published behaviour, independently of the test harness' restatement in
oracle/downstream_ref.py, against which tests/test_waveforms.py checks them case by case:
    tsmoothie.smoother.SpectralSmoother(smooth_fraction, pad_len).smooth(x) -> .smooth_data[0]
    peakutils.peak.indexes(y, thres, min_dist)

Everything here is O(frames) work on 64-300 samples and stays on the host like in the reference; the heavy
(N, H, W) reductions that produce the waveforms are the CUDA kernels behind analysis.py / TVL1Engine.analyze_clip.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .config import CardiacCycleConfig, PeakDetectionConfig


# ------------------------------------------------------------------------------------------ tsmoothie subset
class SpectralSmoother:
    """Low-pass by zeroing the upper rfft bins of the symmetrically padded series (tsmoothie's SpectralSmoother).
    `smooth(x)` stores the result in `smooth_data` with tsmoothie's (n_series, n_samples) layout."""

    def __init__(self, smooth_fraction: float, pad_len: int):
        if not 0 < smooth_fraction < 1:
            raise ValueError("smooth_fraction must be in the range (0,1)")
        if pad_len < 1:
            raise ValueError("pad_len must be >= 1")
        self.smooth_fraction = float(smooth_fraction)
        self.pad_len = int(pad_len)
        self.smooth_data: Optional[np.ndarray] = None
        self.data: Optional[np.ndarray] = None

    def smooth(self, data) -> "SpectralSmoother":
        series = np.atleast_2d(np.asarray(data, dtype=np.float64))          # (n_series, n_samples)
        n = series.shape[1]
        if self.pad_len >= n:
            # np.pad(mode='symmetric') would silently wrap more than once; tsmoothie's output for that case is not
            # part of the contract the reference relies on (pad_len 20 with >= 21 frames)
            raise ValueError(f"pad_len ({self.pad_len}) must be smaller than the series length ({n})")
        k = self.pad_len
        # symmetric padding mirrors INCLUDING the edge sample: x[k-1] .. x[0] | x | x[n-1] .. x[n-k]
        padded = np.concatenate([series[:, k - 1::-1] if k > 0 else series[:, :0], series, series[:, :n - k - 1:-1]], axis=1)
        spectrum = np.fft.rfft(padded, axis=1)
        cut = int(spectrum.shape[1] * self.smooth_fraction)
        spectrum[:, cut:] = 0
        smooth = np.fft.irfft(spectrum, n=padded.shape[1], axis=1)
        self.data = series
        self.smooth_data = smooth[:, k:k + n]
        return self


def spectral_smooth(x, smooth_fraction: float, pad_len: int) -> np.ndarray:
    """`SpectralSmoother(...).smooth(x); smooth_data[0]` as one call."""
    return SpectralSmoother(smooth_fraction, pad_len).smooth(x).smooth_data[0]


# ------------------------------------------------------------------------------------------ peakutils subset
def _slopes_with_plateaus_resolved(y: np.ndarray) -> Optional[np.ndarray]:
    """First differences of y with every run of zeros replaced by a neighbouring slope (peakutils' plateau rule):
    a leading run copies the slope after it, a trailing run the slope before it, an inner run takes the slope
    before it for the positions left of the run's middle and the slope after it from the middle on.  None when y
    is flat."""
    dy = np.diff(y)
    m = dy.size
    if m == 0 or not np.any(dy != 0):
        return None
    out = dy.copy()
    i = 0
    while i < m:
        if dy[i] != 0:
            i += 1
            continue
        j = i
        while j + 1 < m and dy[j + 1] == 0:
            j += 1                                                     # zero run dy[i..j]
        if i == 0:
            out[i:j + 1] = dy[j + 1]
        elif j == m - 1:
            out[i:j + 1] = dy[i - 1]
        else:
            mid = 0.5 * (i + j)                                        # np.median of the run's positions
            for k in range(i, j + 1):
                out[k] = dy[i - 1] if k < mid else dy[j + 1]
        i = j + 1
    return out


def peak_indexes(y, thres: float = 0.3, min_dist: int = 1) -> np.ndarray:
    """peakutils.peak.indexes(y, thres, min_dist) with the default thres_abs=False: local maxima above
    thres * (max - min) + min, then greedy suppression from the highest peak down within +-min_dist samples.
    Ascending int64 indices."""
    y = np.asarray(y, dtype=np.float64)
    if y.size == 0:
        raise ValueError("peak_indexes of an empty array")           # np.max in peakutils raises the same way
    level = thres * (np.max(y) - np.min(y)) + np.min(y)
    dy = _slopes_with_plateaus_resolved(y)
    if dy is None:
        return np.array([], dtype=np.int64)
    rising = np.concatenate([[0.0], dy]) > 0                          # slope into the sample
    falling = np.concatenate([dy, [0.0]]) < 0                         # slope out of it
    peaks = np.flatnonzero(rising & falling & (y > level))
    min_dist = int(min_dist)
    if peaks.size > 1 and min_dist > 1:
        keep = np.zeros(y.size, dtype=bool)
        blocked = np.zeros(y.size, dtype=bool)
        for p in peaks[np.argsort(y[peaks])][::-1]:                   # highest first
            if not blocked[p]:
                keep[p] = True
                blocked[max(0, p - min_dist):p + min_dist + 1] = True
        peaks = np.flatnonzero(keep)
    return peaks.astype(np.int64)


# ------------------------------------------------------------------------------------------ intervals
def find_start_stop(arr) -> List[List[int]]:
    """optical_flow_utils.py:40-49: runs of consecutive integers -> [[first, last], ...]."""
    a = np.atleast_1d(np.asarray(arr))
    if a.size == 0:
        raise IndexError("find_start_stop of an empty index list")   # the reference indexes arr[0] here
    cuts = np.flatnonzero(np.diff(a) != 1) + 1
    firsts = np.concatenate([[0], cuts])
    lasts = np.concatenate([cuts - 1, [a.size - 1]])
    return [[a[i], a[j]] for i, j in zip(firsts, lasts)]


def angle_cycle_intervals(ang_mode_arr, cc_config: Optional[CardiacCycleConfig] = None
                          ) -> Tuple[List[List[int]], List[List[int]]]:
    """The tail of AngleDetector.detect (cardiac_cycle_detection.py:117-127): smooth the per-frame angle mode,
    frames below pi are systole, the others diastole, both as [start, stop] runs."""
    cfg = cc_config or CardiacCycleConfig()
    filt = spectral_smooth(ang_mode_arr, cfg.smooth_fraction, cfg.pad_len)
    up_frames = np.squeeze(np.argwhere(filt < np.pi))
    down_frames = np.squeeze(np.argwhere(filt >= np.pi))
    return find_start_stop(up_frames), find_start_stop(down_frames)


def _angle_diastole(true_sys: Sequence[Sequence[int]], nframes: int) -> List[List[int]]:
    """cc_method == 'angle' (peak_detection.py:177-187, 289-299): diastole = the gaps around the systole runs, in the
    reference's order (head, tail, then the inner gaps)."""
    gaps: List[List[int]] = []
    if len(true_sys) > 0:
        if true_sys[0][0] > 1:
            gaps.append([0, true_sys[0][0] - 1])
        if true_sys[-1][1] < nframes - 2:
            gaps.append([true_sys[-1][1], nframes - 1])
        for left, right in zip(true_sys[:-1], true_sys[1:]):
            gaps.append([left[1], right[0]])
    return gaps


# ------------------------------------------------------------------------------------------ peak picking
def _pick(curve: np.ndarray, lo: int, hi: int, all_peaks: np.ndarray, cfg: PeakDetectionConfig, sign: float,
          what: str) -> Tuple[int, bool]:
    """One peak inside the window [lo, hi] of `curve`: the extreme candidate (largest for sign=+1, smallest for -1)
    among the peaks found inside the window (`pick_peak_by_subset`) or among `all_peaks`; without a candidate the
    extreme sample of curve[lo:hi] (sic: the fallback window excludes hi, like the reference).  -> (index, found)"""
    if cfg.pick_peak_by_subset:
        cand = peak_indexes(sign * curve[lo:hi + 1], thres=cfg.peak_thres, min_dist=cfg.min_dist) + lo
    else:
        cand = np.array([k for k in all_peaks if lo <= k <= hi], dtype=np.int64)
    if len(cand) > 0:
        vals = sign * curve[cand]
        return int(cand[int(np.argmax(vals))]), True
    print(f"Warning no {what} peak found! Using max value")
    return int(np.argmax(sign * curve[lo:hi])) + lo, False


class PeakDetector:
    """peak_detection.py:16-136."""

    def __init__(self, peak_config: Optional[PeakDetectionConfig] = None,
                 cc_config: Optional[CardiacCycleConfig] = None):
        self.peak_config = peak_config or PeakDetectionConfig()
        self.cc_config = cc_config or CardiacCycleConfig()

    def detect_systolic_peaks(self, filt_lo, sys_frames, lo_peaks_i):
        """most negative peak of the LOW percentile curve inside every systole run -> (sys_i, true_sys); runs
        without a candidate still contribute an index but are dropped from true_sys (:48-57)"""
        sys_i, true_sys = [], []
        for start, stop in sys_frames:
            idx, found = _pick(np.asarray(filt_lo), int(start), int(stop), lo_peaks_i, self.peak_config, -1.0, "systolic")
            sys_i.append(idx)
            if found:
                true_sys.append([start, stop])
        return sys_i, true_sys

    def detect_diastolic_peaks(self, filt_hi, dia_frames, hi_peaks_i, nframes):
        """e', l', a': the highest peak of the HIGH percentile curve in each third of every diastole run (:77-136)"""
        out: Tuple[List[int], List[int], List[int]] = ([], [], [])
        curve = np.asarray(filt_hi)
        for start, stop in dia_frames:
            third = np.floor((stop - start) / 3)
            e0, e1 = int(start), int(start + third)
            l0 = int(e1 + 1); l1 = int(l0 + third)
            a0, a1 = int(l1 + 1), int(stop + 1)
            for dst, (w0, w1), name in zip(out, ((e0, e1), (l0, l1), (a0, a1)), ("e'", "l'", "a'")):
                dst.append(_pick(curve, w0, w1, hi_peaks_i, self.peak_config, +1.0, name)[0])
        return out


def calculate_radlong_peaks(hi_arr, lo_arr, frame_times, sys_frames, dia_frames, nframes: int, cc_method: str = 'angle',
                            smooth_fraction: float = 0.3, pad_len: int = 20, peak_thres: float = 0.5, min_dist: int = 5,
                            pick_peak_by_subset: bool = False) -> Dict[str, object]:
    """peak_detection.py:139-226, same keys in the returned dict plus the frame INDICES ('sys_i', 'e_i', 'l_i', 'a_i')."""
    filt_lo = spectral_smooth(lo_arr, smooth_fraction, pad_len)
    filt_hi = spectral_smooth(hi_arr, smooth_fraction, pad_len)
    hi_peaks = peak_indexes(filt_hi, thres=peak_thres, min_dist=min_dist)
    lo_peaks = peak_indexes(-filt_lo, thres=peak_thres, min_dist=min_dist)
    if cc_method == 'angle':
        true_sys, true_dia = sys_frames, _angle_diastole(sys_frames, nframes)
    else:
        true_sys, true_dia = sys_frames, dia_frames
    det = PeakDetector(PeakDetectionConfig(peak_thres=peak_thres, min_dist=min_dist, pick_peak_by_subset=pick_peak_by_subset))
    sys_i, true_sys_kept = det.detect_systolic_peaks(filt_lo, true_sys, lo_peaks)
    e_i, l_i, a_i = det.detect_diastolic_peaks(filt_hi, true_dia, hi_peaks, nframes)
    ft = np.asarray(frame_times)
    return {'filt_hi': filt_hi, 'filt_lo': filt_lo, 'true_sys': true_sys_kept, 'true_dia': true_dia,
            'sys_px': ft[sys_i], 'sys_py': filt_lo[sys_i], 'e_px': ft[e_i], 'e_py': filt_hi[e_i],
            'l_px': ft[l_i], 'l_py': filt_hi[l_i], 'a_px': ft[a_i], 'a_py': filt_hi[a_i],
            'sys_i': sys_i, 'e_i': e_i, 'l_i': l_i, 'a_i': a_i}


def calculate_single_peaks(filt_arr, frame_times, sys_frames, dia_frames, nframes: int, cc_method: str = 'angle',
                           peak_thres: float = 0.2, min_dist: int = 5, pick_peak_by_subset: bool = False,
                           show_all_peaks: bool = False) -> Dict[str, object]:
    """peak_detection.py:229-375: one curve (e.g. the smoothed 99th percentile of |v|); systolic peaks are MAXIMA
    here, and with cc_method='angle' the diastole runs are derived from the systole runs that had a peak."""
    curve = np.asarray(filt_arr)
    cfg = PeakDetectionConfig(peak_thres=peak_thres, min_dist=min_dist, pick_peak_by_subset=pick_peak_by_subset)
    peaks = peak_indexes(curve, thres=peak_thres, min_dist=min_dist)
    sys_i, true_sys = [], []
    for start, stop in sys_frames:
        idx, found = _pick(curve, int(start), int(stop), peaks, cfg, +1.0, "sys")
        sys_i.append(idx)
        if found:
            true_sys.append([start, stop])
    if cc_method == 'angle':
        true_dia = _angle_diastole(true_sys, nframes)
    else:
        true_dia, true_sys = dia_frames, sys_frames
    e_i, l_i, a_i = PeakDetector(cfg).detect_diastolic_peaks(curve, true_dia, peaks, nframes)
    ft = np.asarray(frame_times)
    res = {'filt_arr': curve, 'true_sys': true_sys, 'true_dia': true_dia,
           'sys_px': ft[sys_i], 'sys_py': curve[sys_i], 'e_px': ft[e_i], 'e_py': curve[e_i],
           'l_px': ft[l_i], 'l_py': curve[l_i], 'a_px': ft[a_i], 'a_py': curve[a_i],
           'sys_i': sys_i, 'e_i': e_i, 'l_i': l_i, 'a_i': a_i}
    if show_all_peaks:
        res['all_px'] = ft[peaks]
        res['all_py'] = curve[peaks]
    return res


# ------------------------------------------------------------------------------------------ whole clip
def clip_waveform_indices(analysis: Dict[str, np.ndarray], nframes: int, frame_rate: float = 1.0,
                          cc_config: Optional[CardiacCycleConfig] = None,
                          peak_config: Optional[PeakDetectionConfig] = None,
                          single_smooth_fraction: float = 0.5, strict: bool = True) -> Dict[str, object]:
    """The reference's downstream order of operations (example_peak_plots.py:124-267) on the per-frame waveforms of
    one label, as TVL1Engine.analyze_clip returns them ('ang_mode', 'mag_hi', 'rad_hi', 'rad_lo', 'long_hi',
    'long_lo'): angle-based systole / diastole runs, then systolic / e' / l' / a' frame indices of the magnitude
    curve and of the radial and longitudinal curve pairs.  Defaults = the reference's configs (config.py:13-16, 75-82).
    The reference's pickers raise ValueError when a fallback window is empty (np.argmax of an empty slice,
    peak_detection.py:56, 115-133); strict=True propagates that like the reference, strict=False records
    {'raises': 'ValueError'} for that curve and goes on with the others."""
    cc = cc_config or CardiacCycleConfig()
    pk = peak_config or PeakDetectionConfig()
    frame_times = np.arange(nframes) * (1000.0 / frame_rate)
    sys_frames, dia_frames = angle_cycle_intervals(np.asarray(analysis['ang_mode'], dtype=np.float64)[:nframes], cc)
    out: Dict[str, object] = {'sys_frames': [[int(a), int(b)] for a, b in sys_frames],
                              'dia_frames': [[int(a), int(b)] for a, b in dia_frames]}

    def stage(fn):
        if strict:
            return fn()
        try:
            return fn()
        except ValueError:
            return {'raises': 'ValueError'}

    out['single'] = stage(lambda: calculate_single_peaks(
        spectral_smooth(np.asarray(analysis['mag_hi'], dtype=np.float64)[:nframes], single_smooth_fraction, pk.pad_len),
        frame_times, sys_frames, dia_frames, nframes, 'angle', pk.peak_thres, pk.min_dist, pk.pick_peak_by_subset))
    for name, hi, lo in (('radial', 'rad_hi', 'rad_lo'), ('longitudinal', 'long_hi', 'long_lo')):
        out[name] = stage(lambda hi=hi, lo=lo: calculate_radlong_peaks(
            np.asarray(analysis[hi])[:nframes], np.asarray(analysis[lo])[:nframes], frame_times, sys_frames, dia_frames,
            nframes, 'angle', pk.smooth_fraction, pk.pad_len, pk.peak_thres, pk.min_dist, pk.pick_peak_by_subset))
    return out


def indices_of(result: Dict[str, object]) -> Dict[str, object]:
    """Only the integer outcome of clip_waveform_indices (what the north star wants bit-exact)."""
    def pick(d):
        if 'raises' in d:
            return dict(d)
        return {k: [int(i) for i in d[k]] for k in ('sys_i', 'e_i', 'l_i', 'a_i')}
    return {'sys_frames': result['sys_frames'], 'dia_frames': result['dia_frames'],
            'single': pick(result['single']), 'radial': pick(result['radial']), 'longitudinal': pick(result['longitudinal'])}
