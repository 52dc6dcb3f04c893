"""CPU tests of the product's downstream waveform pipeline (tee_optical_flow_b200/waveforms.py) against the test
harness' independent restatement (oracle/downstream_ref.py) and against hand-checked cases of the two third-party
helpers (tsmoothie SpectralSmoother, peakutils indexes) -- reference: cardiac_cycle_detection.py:87-143,
peak_detection.py:16-375, optical_flow_utils.py:40-49."""
import numpy as np
import pytest

from oracle import downstream_ref as R
from tee_optical_flow_b200 import waveforms as Wv
from tee_optical_flow_b200.config import CardiacCycleConfig, PeakDetectionConfig


def _curves(seed, n):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    base = np.sin(2 * np.pi * t / rng.uniform(18, 40) + rng.uniform(0, 6))
    return base, rng


@pytest.mark.parametrize("n,frac,pad", [(62, 0.2, 20), (62, 0.3, 20), (62, 0.5, 20), (24, 0.3, 10), (298, 0.2, 20), (21, 0.5, 20)])
def test_spectral_smoother_equals_padded_rfft_lowpass(n, frac, pad):
    x, rng = _curves(n, n)
    x = x + 0.3 * rng.standard_normal(n)
    got = Wv.SpectralSmoother(frac, pad).smooth(x).smooth_data
    assert got.shape == (1, n)
    # independent statement: np.pad symmetric + rfft cut
    p = np.pad(x, pad, mode="symmetric")
    F = np.fft.rfft(p); F[int(len(F) * frac):] = 0
    want = np.fft.irfft(F, n=len(p))[pad:-pad]
    assert np.array_equal(got[0], want)
    assert np.array_equal(got[0], R.spectral_smooth(x, frac, pad))
    with pytest.raises(ValueError):
        Wv.SpectralSmoother(frac, n).smooth(x)
    with pytest.raises(ValueError):
        Wv.SpectralSmoother(1.0, pad)


def test_peak_indexes_known_cases():
    y = np.array([0, 1, 0, 2, 0, 3, 0, 2, 0, 1, 0], float)
    assert Wv.peak_indexes(y, thres=0.0, min_dist=1).tolist() == [1, 3, 5, 7, 9]
    assert Wv.peak_indexes(y, thres=0.5, min_dist=1).tolist() == [3, 5, 7]           # y > 1.5
    assert Wv.peak_indexes(y, thres=0.0, min_dist=2).tolist() == [1, 5, 9]           # 5 suppresses 3 and 7, then 1, 9
    assert Wv.peak_indexes(y, thres=0.0, min_dist=3).tolist() == [1, 5, 9]
    # plateaus: the peak is reported at the middle of the flat top (left half rises, the rest falls)
    assert Wv.peak_indexes(np.array([0, 1, 1, 1, 0], float), 0.0, 1).tolist() == [2]
    assert Wv.peak_indexes(np.array([0, 1, 1, 0], float), 0.0, 1).tolist() == [1]     # even run: dy = [1, 0, -1] -> [1, -1, -1]
    assert Wv.peak_indexes(np.array([1, 1, 0, 2, 0], float), 0.0, 1).tolist() == [3]   # leading flat run
    assert Wv.peak_indexes(np.array([0, 2, 0, 1, 1], float), 0.0, 1).tolist() == [1]   # trailing flat run
    assert Wv.peak_indexes(np.ones(9)).size == 0 and Wv.peak_indexes(np.array([3.0])).size == 0
    with pytest.raises(ValueError):
        Wv.peak_indexes(np.array([]))


@pytest.mark.parametrize("seed", range(40))
def test_peak_indexes_equals_harness_restatement(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 120))
    y = np.cumsum(rng.standard_normal(n))
    if seed % 3 == 0:
        y = np.round(y)                                  # many plateaus
    thres, md = float(rng.uniform(0, 0.9)), int(rng.integers(1, 12))
    assert Wv.peak_indexes(y, thres, md).tolist() == R.peak_indexes(y, thres, md).tolist()


def test_find_start_stop():
    assert Wv.find_start_stop(np.array([0, 1, 2, 5, 6, 9])) == [[0, 2], [5, 6], [9, 9]]
    assert Wv.find_start_stop(np.int64(4)) == [[4, 4]]            # np.squeeze of a single hit is 0-d
    assert Wv.find_start_stop(np.arange(7)) == [[0, 6]]
    with pytest.raises(IndexError):
        Wv.find_start_stop(np.array([], dtype=int))


@pytest.mark.parametrize("seed", range(12))
def test_angle_intervals_and_radlong_peaks_equal_harness(seed):
    """reference defaults (config.py:13-16, 75-82): smoothing 0.2 / 20, peaks 0.2 / 5 / subset picking / 0.3 / 20"""
    n = 62
    base, rng = _curves(100 + seed, n)
    ang = np.pi + 1.2 * base + 0.15 * rng.standard_normal(n)
    hi = 2.0 + np.maximum(base, 0) * 3 + 0.2 * rng.standard_normal(n)
    lo = -2.0 + np.minimum(base, 0) * 3 + 0.2 * rng.standard_normal(n)
    cc, pk = CardiacCycleConfig(), PeakDetectionConfig()
    sys_f, dia_f = Wv.angle_cycle_intervals(ang, cc)
    sys_r, dia_r = R.angle_detector_intervals(ang, cc.smooth_fraction, cc.pad_len)
    assert [list(map(int, s)) for s in sys_f] == [list(map(int, s)) for s in sys_r]
    assert [list(map(int, s)) for s in dia_f] == [list(map(int, s)) for s in dia_r]
    for subset in (True, False):
        kw = dict(smooth_fraction=pk.smooth_fraction, pad_len=pk.pad_len, peak_thres=pk.peak_thres, min_dist=pk.min_dist,
                  pick_peak_by_subset=subset)
        try:
            want = R.radlong_peak_indices(hi, lo, sys_r, n, **kw)
        except ValueError:                              # an empty fallback window: the reference raises too
            with pytest.raises(ValueError):
                Wv.calculate_radlong_peaks(hi, lo, np.arange(n) * 25.0, sys_f, dia_f, n, 'angle', **kw)
            continue
        got = Wv.calculate_radlong_peaks(hi, lo, np.arange(n) * 25.0, sys_f, dia_f, n, 'angle', **kw)
        assert [int(i) for i in got['sys_i']] == want['sys'] and [int(i) for i in got['e_i']] == want['e']
        assert [int(i) for i in got['l_i']] == want['l'] and [int(i) for i in got['a_i']] == want['a']
        assert [list(map(int, d)) for d in got['true_dia']] == want['true_dia']
        assert np.array_equal(got['sys_px'], np.arange(n)[got['sys_i']] * 25.0)
        assert np.array_equal(got['e_py'], got['filt_hi'][got['e_i']])


def test_single_peaks_follow_reference_rules():
    n = 62
    t = np.arange(n)
    curve = 1.0 + np.sin(2 * np.pi * t / 31.0) ** 2 + 0.05 * np.cos(t)
    sys_frames = [[6, 12], [34, 45]]
    res = Wv.calculate_single_peaks(curve, t * 20.0, sys_frames, [], n, 'angle', 0.2, 5, True)
    assert res['true_sys'] == sys_frames
    # diastole = head gap, tail gap, inner gaps -- in the reference's order (:289-299)
    assert res['true_dia'] == [[0, 5], [45, 61], [12, 34]]
    for (s0, s1), i in zip(sys_frames, res['sys_i']):
        assert s0 <= i <= s1 and curve[i] == curve[s0:s1 + 1].max()
    assert len(res['e_i']) == len(res['l_i']) == len(res['a_i']) == 3
    # a head gap too short for three windows: the reference's fallback argmax of an empty slice raises -- so does this
    with pytest.raises(ValueError):
        Wv.calculate_single_peaks(curve, t * 20.0, [[3, 12], [34, 45]], [], n, 'angle', 0.2, 5, True)
    # a systole run without a peak inside is dropped from true_sys (but still contributes an index) (:263-270)
    res2 = Wv.calculate_single_peaks(curve, t * 20.0, [[9, 14], [34, 45]], [], n, 'angle', 0.2, 5, True)
    assert res2['true_sys'] == [[34, 45]] and len(res2['sys_i']) == 2
    other = Wv.calculate_single_peaks(curve, t * 20.0, sys_frames, [[15, 33]], n, 'ecg', 0.2, 5, False, show_all_peaks=True)
    assert other['true_dia'] == [[15, 33]] and 'all_px' in other


def test_clip_waveform_indices_runs_on_analysis_dict():
    n = 62
    base, rng = _curves(7, n)
    analysis = {'ang_mode': (np.pi + 1.0 * base).astype(np.float32), 'mag_hi': (2 + base).astype(np.float32),
                'rad_hi': 1 + np.maximum(base, 0), 'rad_lo': -1 + np.minimum(base, 0),
                'long_hi': 1 + np.maximum(-base, 0), 'long_lo': -1 + np.minimum(-base, 0)}
    out = Wv.indices_of(Wv.clip_waveform_indices(analysis, n, frame_rate=40.0))
    assert set(out) == {'sys_frames', 'dia_frames', 'single', 'radial', 'longitudinal'}
    assert all(isinstance(i, int) for i in out['radial']['e_i'])
    assert out['sys_frames'] and out['dia_frames']


def test_whole_chain_on_oracle_flow_equals_harness(oracle):
    """a 64-frame synthetic clip solved by the CPU oracle, stored as fp16: the product's waveform pipeline fed with
    the harness' waveforms gives the harness' indices (reference defaults everywhere) -- and the chain does not
    degenerate on the synthetic clip (there ARE systole runs and peaks to compare)"""
    from tee_optical_flow_b200.synth import make_clip, make_masks
    n, h, w = 64, 120, 160
    fr = make_clip(seed=0, n_frames=n, H=h, W=w, peak_disp=3.0, period=32.0)
    masks = make_masks(0, n, h, w)
    om = oracle.OracleDualTVL1(err_mode=0)
    flows = [om.calc(fr[i], fr[i + 1]) for i in range(n - 1)]
    flows.append(flows[-1])
    f16 = (np.stack(flows) * np.float32(2.0)).astype(np.float16)
    want = R.clip_indices(f16, masks["rv"], masks["av"], n - 2)
    got = Wv.indices_of(Wv.clip_waveform_indices(want["waveforms"], n - 2, frame_rate=40.0))
    for k in ("sys_frames", "dia_frames", "single", "radial", "longitudinal"):
        assert got[k] == want[k], k
    assert len(want["sys_frames"]) >= 1 and len(want["radial"]["e_i"]) >= 1
