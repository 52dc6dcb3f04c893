"""numpy restatement of the reference's host-side code AROUND the TV-L1 solver: WASE background compensation,
the radial / longitudinal decomposition with its per-frame reductions, the angle-mode detector and the peak
pickers.  TEST INFRASTRUCTURE ONLY (never imported by the product).

Each function follows the reference lines it cites.  Third-party helpers that are missing in this image are
restated from their published behaviour: tsmoothie.SpectralSmoother, peakutils.indexes,
skimage.measure.label/regionprops (scipy.ndimage.label with full connectivity is equivalent).
cv2.cartToPolar, np.percentile, np.histogram, scipy.stats.mode and scipy.signal.savgol_filter are the genuine
functions the reference calls.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ calculate_optical_flow.py:649-660
def wase_background(flow: np.ndarray, bkgd_mask: np.ndarray) -> np.float32:
    """`mask = mask_dict['bkgd']` is (N, H, W, 2) for ALL frames; `flow * mask` broadcasts; the background is
    ONE scalar: the mean of every non-zero entry (both channels, all N masks)."""
    masked = flow * bkgd_mask
    return np.mean(masked[masked != 0])


def wase_weight_map(bkgd_mask: np.ndarray) -> np.ndarray:
    """Equivalent form: w[y,x,c] = sum_n bkgd[n,y,x,c]; mean = sum(w f [f!=0]) / sum(w [f!=0])."""
    return bkgd_mask.sum(axis=0).astype(np.float32)


# ------------------------------------------------------------------ analysis.py:39-86 (skimage restated)
def calc_av_centroid(mask_arr: np.ndarray, nframes: int, do_filter: bool = True, savgol_window: int = 10,
                     savgol_poly: int = 4):
    from scipy import ndimage
    from scipy.signal import savgol_filter
    cents = []
    for i in range(nframes):
        frame = np.squeeze(mask_arr[i, :, :, 0])
        lab, n = ndimage.label(frame, structure=np.ones((3, 3)))   # skimage.measure.label default: 8-connectivity
        if n >= 1:
            areas = ndimage.sum_labels(np.ones_like(lab), lab, index=np.arange(1, n + 1))
            k = int(np.argmax(areas)) + 1                           # first largest, like np.argmax over props
            rr, cc = np.nonzero(lab == k)
            cents.append((rr.mean(), cc.mean()))                   # regionprops.centroid = mean (row, col)
        elif cents:
            cents.append(cents[i - 1])
        else:
            cents.append((mask_arr.shape[1] / 2, mask_arr.shape[2] / 2))
    if do_filter and len(cents) >= savgol_window:
        cents = savgol_filter(cents, savgol_window, savgol_poly, axis=0)
    return np.asarray(cents, dtype=np.float64)


# ------------------------------------------------------------------ analysis.py:89-163
def radial_vecgrid(H: int, W: int, centroids, nframes: int) -> np.ndarray:
    """unit vector from every pixel to the AV centroid, (N, H, W, 2) float64: ch0 = row part, ch1 = col part;
    0 at the centroid itself (nan_to_num)."""
    rows = np.arange(H, dtype=np.float64)[:, None]
    cols = np.arange(W, dtype=np.float64)[None, :]
    out = np.empty((nframes, H, W, 2), np.float64)
    for i in range(nframes):
        v0 = np.broadcast_to(centroids[i][0] - rows, (H, W))
        v1 = np.broadcast_to(centroids[i][1] - cols, (H, W))
        norm = np.sqrt(np.abs(v0) ** 2 + np.abs(v1) ** 2)           # np.linalg.norm(vec, axis=2)
        with np.errstate(invalid="ignore", divide="ignore"):
            out[i, ..., 0] = np.nan_to_num(v0 / norm, nan=0)
            out[i, ..., 1] = np.nan_to_num(v1 / norm, nan=0)
    return out


def comp_magnitude(of_arr: np.ndarray, centroids):
    """rad = OF . unit ; long = OF . (unit[1], -unit[0])  -- flow channel 0 (dx) is paired with the ROW unit
    component, exactly as the reference does (analysis.py:134,156)."""
    n = len(centroids)
    of_arr = of_arr[:n]
    H, W = of_arr.shape[1:3]
    unit = radial_vecgrid(H, W, centroids, n)
    ortho = np.stack([unit[..., 1], -1 * unit[..., 0]], axis=-1)
    return np.sum(of_arr * unit, axis=3), np.sum(of_arr * ortho, axis=3)


# ------------------------------------------------------------------ analysis.py:166-212
def bidirectional_hist(mag_arr: np.ndarray, nframes: int, perc_lo=1, perc_hi=99, nbins=1000):
    mmax, mmin = np.max(mag_arr), np.min(mag_arr)
    edges = []
    hi, lo, freq = [], [], []
    for i in range(nframes):
        flat = np.ravel(mag_arr[i])
        nz = flat[flat != 0]
        if len(nz) == 0:
            if hi:
                hi.append(hi[-1]); lo.append(lo[-1]); freq.append(freq[-1])
            else:
                hi.append(mmax); lo.append(mmin); freq.append(np.ones(nbins))
        else:
            hi.append(np.percentile(nz, perc_hi)); lo.append(np.percentile(nz, perc_lo))
            f, edges = np.histogram(nz, bins=nbins, range=(mmin, mmax))
            freq.append(f + 1)
    return np.stack(freq), edges, np.asarray(hi), np.asarray(lo)


# ------------------------------------------------------------------ analysis.py:215-286
def hist3d(masked_arr: np.ndarray, nframes: int, nbins=1000, percentile=99):
    import cv2
    mags, angs = [], []
    for i in range(nframes):
        flow = np.squeeze(masked_arr[i])
        m, a = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        mags.append(m); angs.append(a)
    mag_arr, ang_arr = np.stack(mags), np.stack(angs)

    def per_frame(arr, want_pct):
        amax, amin = np.max(arr), np.min(arr)
        freqs, pct, edges = [], [], []
        for i in range(nframes):
            flat = np.ravel(arr[i]); nz = flat[flat != 0]
            if len(nz) == 0:
                if freqs:
                    freqs.append(freqs[-1])
                    if want_pct: pct.append(pct[-1])
                else:
                    f, edges = np.histogram([amax], bins=nbins, range=(amin, amax))
                    freqs.append(f + 1)
                    if want_pct: pct.append(amax)
            else:
                if want_pct: pct.append(np.percentile(nz, percentile))
                f, edges = np.histogram(nz, bins=nbins, range=(amin, amax))
                freqs.append(f + 1)
        return np.stack(freqs), edges, np.asarray(pct)

    mag_f, mag_e, hi = per_frame(mag_arr, True)
    ang_f, ang_e, _ = per_frame(ang_arr, False)
    return mag_f, ang_f, mag_e, ang_e, hi


# ------------------------------------------------------------------ cardiac_cycle_detection.py:100-116
def angle_mode(masked_arr: np.ndarray, nframes: int) -> np.ndarray:
    import cv2
    from scipy.stats import mode
    out = []
    for i in range(nframes):
        flow = np.squeeze(masked_arr[i])
        _, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        flat = np.ravel(np.round(ang, decimals=2))
        nz = flat[flat != 0]
        out.append(mode(nz).mode)
    return np.asarray(out)


# ------------------------------------------------------------------ tsmoothie.SpectralSmoother (restated)
def spectral_smooth(x, smooth_fraction: float, pad_len: int) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    if not (0 < smooth_fraction < 1) or pad_len >= len(x):
        raise ValueError("SpectralSmoother: need 0 < smooth_fraction < 1 and pad_len < len(x)")
    pad = np.pad(x, pad_len, mode="symmetric")
    F = np.fft.rfft(pad)
    F[int(len(F) * smooth_fraction):] = 0
    y = np.fft.irfft(F, n=len(pad))
    return y[pad_len:-pad_len]


# ------------------------------------------------------------------ peakutils.peak.indexes (restated)
def peak_indexes(y, thres=0.3, min_dist=1) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    if len(y) < 3 or np.max(y) == np.min(y):
        return np.array([], dtype=int)
    thres_abs = thres * (np.max(y) - np.min(y)) + np.min(y)
    min_dist = int(min_dist)
    dy = np.diff(y)
    zeros, = np.where(dy == 0)
    if len(zeros) == len(y) - 1:
        return np.array([], dtype=int)
    if len(zeros):
        # plateaus: propagate the neighbouring slope into the flat run (left half from the left, rest from the right)
        zeros_diff = np.diff(zeros)
        zeros_diff_not_one, = np.add(np.where(zeros_diff != 1), 1)
        zero_plateaus = np.split(zeros, zeros_diff_not_one)
        if zero_plateaus[0][0] == 0:
            dy[zero_plateaus[0]] = dy[zero_plateaus[0][-1] + 1]
            zero_plateaus.pop(0)
        if len(zero_plateaus) and zero_plateaus[-1][-1] == len(dy) - 1:
            dy[zero_plateaus[-1]] = dy[zero_plateaus[-1][0] - 1]
            zero_plateaus.pop(-1)
        for plateau in zero_plateaus:
            median = np.median(plateau)
            dy[plateau[plateau < median]] = dy[plateau[0] - 1]
            dy[plateau[plateau >= median]] = dy[plateau[-1] + 1]
    peaks = np.where((np.hstack([dy, 0.0]) < 0.0) & (np.hstack([0.0, dy]) > 0.0) & (np.greater(y, thres_abs)))[0]
    if peaks.size > 1 and min_dist > 1:
        highest = peaks[np.argsort(y[peaks])][::-1]
        rem = np.ones(y.size, dtype=bool)
        rem[peaks] = False
        for peak in highest:
            if not rem[peak]:
                sl = slice(max(0, peak - min_dist), peak + min_dist + 1)
                rem[sl] = True
                rem[peak] = False
        peaks = np.arange(y.size)[~rem]
    return peaks


# ------------------------------------------------------------------ optical_flow_utils.py:40-49
def find_start_stop(arr):
    arr = np.atleast_1d(arr)
    breaks = np.where(np.diff(arr) != 1)[0] + 1
    clusters, start = [], 0
    for end in breaks:
        clusters.append([arr[start], arr[end - 1]])
        start = end
    clusters.append([arr[start], arr[-1]])
    return clusters


# ------------------------------------------------------------------ cardiac_cycle_detection.py:117-127
def angle_detector_intervals(ang_mode_arr, smooth_fraction=0.2, pad_len=20):
    filt = spectral_smooth(ang_mode_arr, smooth_fraction, pad_len)
    up = np.squeeze(np.argwhere(filt < np.pi))
    down = np.squeeze(np.argwhere(filt >= np.pi))
    sys_frames = find_start_stop(up) if np.size(up) else []
    dia_frames = find_start_stop(down) if np.size(down) else []
    return sys_frames, dia_frames


# ------------------------------------------------------------------ peak_detection.py:16-226
def radlong_peak_indices(hi_arr, lo_arr, sys_frames, nframes, smooth_fraction=0.3, pad_len=20, peak_thres=0.5,
                         min_dist=5, pick_peak_by_subset=False):
    """calculate_radlong_peaks with cc_method='angle': returns the FRAME INDICES (sys, e', l', a')."""
    filt_lo = spectral_smooth(lo_arr, smooth_fraction, pad_len)
    filt_hi = spectral_smooth(hi_arr, smooth_fraction, pad_len)
    hi_peaks = peak_indexes(filt_hi, peak_thres, min_dist)
    lo_peaks = peak_indexes(filt_lo * -1, peak_thres, min_dist)
    true_sys, true_dia = sys_frames, []
    if len(true_sys) > 0:
        if true_sys[0][0] > 1:
            true_dia.append([0, true_sys[0][0] - 1])
        if true_sys[-1][1] < (nframes - 2):
            true_dia.append([true_sys[-1][1], nframes - 1])
        for i in range(len(true_sys) - 1):
            true_dia.append([true_sys[i][1], true_sys[i + 1][0]])
    sys_i, kept_sys = [], []
    for start, stop in true_sys:
        if pick_peak_by_subset:
            cand = peak_indexes(filt_lo[start:stop + 1] * -1, peak_thres, min_dist) + start
        else:
            cand = [k for k in lo_peaks if start <= k <= stop]
        if len(cand) > 0:
            sys_i.append(int(cand[int(np.argmin([filt_lo[i] for i in cand]))]))
            kept_sys.append([start, stop])                     # runs without a candidate leave true_sys (:52-56)
        else:
            sys_i.append(int(np.argmin(filt_lo[start:stop]) + start))
    e_i, l_i, a_i = [], [], []
    for start, stop in true_dia:
        third = np.floor((stop - start) / 3)
        e0, e1 = int(start), int(start + third)
        l0 = int(e1 + 1); l1 = int(l0 + third)
        a0, a1 = int(l1 + 1), int(stop + 1)
        for (s0, s1, dst) in ((e0, e1, e_i), (l0, l1, l_i), (a0, a1, a_i)):
            if pick_peak_by_subset:
                cand = peak_indexes(filt_hi[s0:s1 + 1], peak_thres, min_dist) + s0
            else:
                cand = [k for k in hi_peaks if s0 <= k <= s1]
            if len(cand) > 0:
                dst.append(int(cand[int(np.argmax([filt_hi[i] for i in cand]))]))
            else:
                dst.append(int(np.argmax(filt_hi[s0:s1]) + s0))
    return dict(sys=sys_i, e=e_i, l=l_i, a=a_i, true_sys=[list(map(int, s)) for s in kept_sys],
                true_dia=[list(map(int, d)) for d in true_dia])


# ------------------------------------------------------------------ peak_detection.py:229-375
def single_peak_indices(filt_arr, sys_frames, nframes, peak_thres=0.2, min_dist=5, pick_peak_by_subset=False):
    """calculate_single_peaks with cc_method='angle': FRAME INDICES (sys, e', l', a') of one smoothed curve; systolic
    peaks are maxima, and the diastole runs come from the systole runs that had a peak."""
    peaks = peak_indexes(filt_arr, peak_thres, min_dist)

    def best(lo, hi):
        cand = (peak_indexes(filt_arr[lo:hi + 1], peak_thres, min_dist) + lo) if pick_peak_by_subset else \
            [k for k in peaks if lo <= k <= hi]
        if len(cand) > 0:
            return int(cand[int(np.argmax([filt_arr[i] for i in cand]))]), True
        return int(np.argmax(filt_arr[lo:hi]) + lo), False

    sys_i, true_sys = [], []
    for start, stop in sys_frames:
        i, found = best(int(start), int(stop))
        sys_i.append(i)
        if found:
            true_sys.append([int(start), int(stop)])
    true_dia = []
    if len(true_sys) > 0:
        if true_sys[0][0] > 1:
            true_dia.append([0, true_sys[0][0] - 1])
        if true_sys[-1][1] < nframes - 2:
            true_dia.append([true_sys[-1][1], nframes - 1])
        for i in range(len(true_sys) - 1):
            true_dia.append([true_sys[i][1], true_sys[i + 1][0]])
    e_i, l_i, a_i = [], [], []
    for start, stop in true_dia:
        third = np.floor((stop - start) / 3)
        e0, e1 = int(start), int(start + third)
        l0 = int(e1 + 1); l1 = int(l0 + third)
        a0, a1 = int(l1 + 1), int(stop + 1)
        e_i.append(best(e0, e1)[0]); l_i.append(best(l0, l1)[0]); a_i.append(best(a0, a1)[0])
    return dict(sys=sys_i, e=e_i, l=l_i, a=a_i, true_sys=true_sys, true_dia=true_dia)


# ------------------------------------------------------------------ example_peak_plots.py:124-267
def clip_indices(flow_f16, label_mask, av_mask, nframes, cc_smooth=0.2, cc_pad=20, single_smooth=0.5, smooth_fraction=0.3,
                 pad_len=20, peak_thres=0.2, min_dist=5, pick_peak_by_subset=True, av_filter=True):
    """The reference's downstream chain on a stored clip with its default configs (config.py:13-16, 75-82, 85-95):
    OpticalFlowDataset.vel_array * mask -> AngleDetector -> calculate_3dhist -> SpectralSmoother ->
    calculate_single_peaks; calc_AV_centroid -> calculate_3dhist_radlong -> calculate_radlong_peaks (x2).
    Returns only the integer outcomes plus the waveforms they were computed from."""
    masked = flow_f16.astype(np.float32) * label_mask
    ang = angle_mode(masked, nframes)
    sys_frames, dia_frames = angle_detector_intervals(ang, cc_smooth, cc_pad)
    _, _, _, _, mag_hi = hist3d(masked, nframes)
    cent = calc_av_centroid(av_mask, nframes, do_filter=av_filter)
    rad, lng = comp_magnitude(masked, cent)
    _, _, rhi, rlo = bidirectional_hist(rad, nframes)
    _, _, lhi, llo = bidirectional_hist(lng, nframes)
    kw = dict(smooth_fraction=smooth_fraction, pad_len=pad_len, peak_thres=peak_thres, min_dist=min_dist,
              pick_peak_by_subset=pick_peak_by_subset)
    pick = lambda d: {k + '_i': [int(i) for i in d[k]] for k in ('sys', 'e', 'l', 'a')}

    def stage(fn):
        # the reference's pickers raise ValueError when a fallback window is empty (np.argmax of an empty slice,
        # peak_detection.py:56, 115-133); that outcome is part of the behaviour to reproduce, per stage
        try:
            return pick(fn())
        except ValueError:
            return {'raises': 'ValueError'}

    return {
        'sys_frames': [[int(a), int(b)] for a, b in sys_frames], 'dia_frames': [[int(a), int(b)] for a, b in dia_frames],
        'single': stage(lambda: single_peak_indices(spectral_smooth(mag_hi, single_smooth, pad_len), sys_frames, nframes,
                                                    peak_thres, min_dist, pick_peak_by_subset)),
        'radial': stage(lambda: radlong_peak_indices(rhi, rlo, sys_frames, nframes, **kw)),
        'longitudinal': stage(lambda: radlong_peak_indices(lhi, llo, sys_frames, nframes, **kw)),
        'waveforms': dict(ang_mode=ang, mag_hi=mag_hi, rad_hi=rhi, rad_lo=rlo, long_hi=lhi, long_lo=llo), 'centroids': cent,
    }


# ------------------------------------------------------------------ calculate_optical_flow.py:91-182
def moving_avg_mask(arr, n=4, threshold=0.49):
    arr2 = np.vstack((arr[0:1], arr, arr[-1:], arr[-1:]))
    s = np.cumsum(arr2.astype(float), axis=0)
    s[n:] = s[n:] - s[:-n]
    return (s[n - 1:] / n) > threshold


def remove_small_objects(mask, min_size):
    """skimage.morphology.remove_small_objects on a bool image: connectivity 1 (4-connected) components with
    fewer than min_size pixels are removed"""
    from scipy import ndimage
    lab, n = ndimage.label(mask)                       # default structure: 4-connectivity
    if n == 0:
        return mask.copy()
    sizes = np.bincount(lab.ravel())
    too_small = sizes < min_size
    too_small[0] = False
    out = mask.copy()
    out[too_small[lab]] = False
    return out


def clean_mask(arr, classes: dict, min_size=500):
    """clean_mask for a {label: class_id} table -> {label: (N,H,W,2) bool, 'bkgd': ...}"""
    from scipy.ndimage import binary_fill_holes
    agg = np.zeros(arr.shape, bool)
    out = {}
    for k, cid in classes.items():
        m = moving_avg_mask(arr == cid)
        clean = np.stack([remove_small_objects(binary_fill_holes(m[i]), min_size) for i in range(m.shape[0])])
        agg |= clean
        out[k] = np.repeat(clean[..., None], 2, axis=3)
    out['bkgd'] = np.repeat((~agg)[..., None], 2, axis=3)
    return out
