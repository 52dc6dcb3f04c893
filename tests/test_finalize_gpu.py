"""GPU parity tests for the stages around the solver: WASE background compensation, fp16 HDF5 layout, and the
masked radial / longitudinal decomposition with its per-frame reductions, against the numpy restatement of the
reference's host code (oracle/downstream_ref.py) on the same inputs.

Bars: order statistics, percentiles, histograms and the angle mode are exact (==).  The WASE scalar is a mean
of ~1e7 float32 values: the reference sums them pairwise in float32, the engine in float64 -> relative tolerance
2e-6 on the scalar (stated in DESIGN.md), and the compensated flow is compared with that tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, H, W = 26, 96, 128


@pytest.fixture(scope="module")
def clip():
    from tee_optical_flow_b200.synth import make_clip, make_masks
    frames = make_clip(seed=3, n_frames=N, H=H, W=W, peak_disp=4.0, period=12.0)
    masks = make_masks(3, N, H, W, period=12.0)
    return frames, masks


@pytest.fixture(scope="module")
def engine():
    from tee_optical_flow_b200.engine import TVL1Engine
    e = TVL1Engine(device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def ref():
    from oracle import downstream_ref
    return downstream_ref


def test_wase_matches_reference_formula(engine, clip, ref):
    frames, masks = clip
    plain, _ = engine.calc_clip(frames, duplicate_last=False)
    engine.set_wase_masks(masks["bkgd"])
    try:
        comp, _ = engine.calc_clip(frames, duplicate_last=False)
        bgs = engine.last_backgrounds()
    finally:
        engine.set_wase_masks(None)
    assert len(bgs) == N - 1
    for i in (0, 7, N - 2):
        want_bg = ref.wase_background(plain[i], masks["bkgd"])          # np.mean(masked_flow[masked_flow != 0])
        assert abs(bgs[i] - want_bg) <= 2e-6 * abs(want_bg) + 1e-12
        assert np.array_equal(comp[i], plain[i] - bgs[i])               # flow - background, float32
        assert np.abs(comp[i] - (plain[i] - want_bg)).max() <= 2e-6 * abs(want_bg) + 1e-9
    # the weight-map identity the kernel relies on
    w = ref.wase_weight_map(masks["bkgd"]).astype(np.float64)
    f = plain[3].astype(np.float64)
    nz = f != 0
    assert abs((w * f * nz).sum() / (w * nz).sum() - bgs[3]) <= 1e-6 * abs(bgs[3])


def test_calculate_optical_flow_wrapper(engine, clip, ref):
    from tee_optical_flow_b200.exceptions import OpticalFlowCalculationError
    from tee_optical_flow_b200.flow import calculate_optical_flow
    frames, masks = clip
    plain = calculate_optical_flow(frames[2], frames[3], masks, engine, bkgd_comp='none', OF_algo='TVL1')
    assert np.array_equal(plain, engine.calc(frames[2], frames[3]))
    comp = calculate_optical_flow(frames[2], frames[3], masks, engine, bkgd_comp='WASE', OF_algo='TVL1')
    want = plain - ref.wase_background(plain, masks["bkgd"])
    assert np.abs(comp - want).max() < 1e-6
    with pytest.raises(OpticalFlowCalculationError):
        calculate_optical_flow(frames[2], frames[3], masks, engine, bkgd_comp='median', OF_algo='TVL1')
    with pytest.raises(OpticalFlowCalculationError):
        calculate_optical_flow(frames[2], frames[3], masks, engine, OF_algo='farneback')


def test_process_frames_layout(clip, oracle):
    """the in-memory mirror of the HDF5 layout that OpticalFlowDataset reads (optical_flow_dataset.py:45-111)"""
    from tee_optical_flow_b200.exceptions import ConfigurationError
    from tee_optical_flow_b200.flow import process_frames
    frames, masks = clip
    out = process_frames(frames, masks, pixel_spacing=0.05, frame_rate=40.0, mode='RVIO_2class',
                         frames_are_prepared=True)
    assert out['flow'].shape == (N, H, W, 2) and out['flow'].dtype == np.float16
    assert out['echo'].shape == (N, H, W) and out['echo'].dtype == np.float16
    assert out['attrs']['nframes'] == N and out['attrs']['units_converted'] is True
    assert out['attrs']['labels'] == ['rv', 'av', 'bkgd'] and out['rv'].dtype == bool
    assert np.array_equal(out['flow'][-1], out['flow'][-2])
    # pair 5 == oracle flow * conversion_factor -> float16
    ref_flow = oracle.OracleDualTVL1(err_mode=1).calc(frames[5], frames[6])
    want = (ref_flow * (0.05 * 40.0)).astype(np.float16)
    assert np.array_equal(out['flow'][5], want)
    with pytest.raises(ConfigurationError):
        process_frames(frames, masks, mode='otsu', bkgd_comp='WASE', frames_are_prepared=True)
    with pytest.raises(ConfigurationError):
        process_frames(frames, masks, mode='bogus', frames_are_prepared=True)


@pytest.fixture(scope="module")
def stored(engine, clip):
    frames, masks = clip
    _, f16 = engine.calc_clip(frames, out_scale=2.0, duplicate_last=True, want_f32=False, want_f16=True)
    return f16


def test_decomposition_and_percentiles_exact(engine, clip, stored, ref):
    frames, masks = clip
    nframes = N - 2                                   # OpticalFlowDataset.nframes = attrs['nframes'] - 2
    cent = ref.calc_av_centroid(masks["av"], nframes)
    vel = stored.astype(np.float32)
    masked = vel * masks["rv"]
    res = engine.analyze_clip(stored, masks["rv"], cent, nframes, 1, 99)
    # magnitude 99th percentile of the non-zero entries (analysis.py:260) and angle mode (cardiac_cycle_detection.py:108-114)
    _, _, _, _, hi_ref = ref.hist3d(masked, nframes)
    assert np.array_equal(res["mag_hi"], hi_ref.astype(np.float32))
    assert np.array_equal(res["ang_mode"], ref.angle_mode(masked, nframes).astype(np.float32))
    # radial / longitudinal projections (analysis.py:137-163) + percentiles 1 / 99 (analysis.py:204-205)
    rad, lng = ref.comp_magnitude(masked, cent)
    _, _, rhi, rlo = ref.bidirectional_hist(rad, nframes)
    _, _, lhi, llo = ref.bidirectional_hist(lng, nframes)
    assert np.array_equal(res["rad_hi"], rhi) and np.array_equal(res["rad_lo"], rlo)
    assert np.array_equal(res["long_hi"], lhi) and np.array_equal(res["long_lo"], llo)
    assert res["rad_min"] == rad.min() and res["rad_max"] == rad.max()
    assert res["long_min"] == lng.min() and res["long_max"] == lng.max()


def test_histograms_exact(engine, clip, stored, ref):
    from tee_optical_flow_b200 import analysis as A
    frames, masks = clip
    nframes = N - 2
    cent = ref.calc_av_centroid(masks["av"], nframes)
    masked = stored.astype(np.float32) * masks["rv"]
    mag, ang, mag_e, ang_e, hi = A.calculate_3dhist(engine, stored, masks["rv"], nframes, centroid_list=cent)
    mag_r, ang_r, mag_er, ang_er, hi_r = ref.hist3d(masked, nframes)
    assert np.array_equal(mag, mag_r) and np.array_equal(ang, ang_r)
    assert np.array_equal(mag_e, mag_er) and np.array_equal(ang_e, ang_er) and np.array_equal(hi, hi_r)
    d = A.calculate_3dhist_radlong(engine, stored, masks["rv"], cent, nframes)
    rad, lng = ref.comp_magnitude(masked, cent)
    for key, arr in (("radial", rad), ("longitudinal", lng)):
        f_r, e_r, h_r, l_r = ref.bidirectional_hist(arr, nframes)
        f, e, h, l = d[key]
        assert np.array_equal(f, f_r) and np.array_equal(e, e_r[:-1]) and np.array_equal(h, h_r) and np.array_equal(l, l_r)


def test_empty_frames_follow_reference_fallbacks(engine, clip, stored, ref):
    """frames whose mask is empty: percentile carried forward / array extreme for the first frame"""
    from tee_optical_flow_b200 import analysis as A
    frames, masks = clip
    nframes = N - 2
    m = masks["rv"].copy()
    m[0] = False; m[5] = False; m[6] = False
    cent = ref.calc_av_centroid(masks["av"], nframes)
    masked = stored.astype(np.float32) * m
    d = A.calculate_3dhist_radlong(engine, stored, m, cent, nframes)
    rad, lng = ref.comp_magnitude(masked, cent)
    f_r, e_r, h_r, l_r = ref.bidirectional_hist(rad, nframes)
    f, e, h, l = d["radial"]
    assert np.array_equal(h, h_r) and np.array_equal(l, l_r) and np.array_equal(f, f_r)


def test_downstream_indices_identical(engine, clip, stored, ref, oracle):
    """north star: systole/diastole frame indices (AngleDetector) and e'/l'/a' peak frame indices computed from
    the engine's fp16 flow + GPU reductions equal those from the oracle's fp16 flow + the reference's numpy code"""
    frames, masks = clip
    nframes = N - 2
    om = oracle.OracleDualTVL1(err_mode=0)             # OpenCV-faithful oracle
    flows = [om.calc(frames[i], frames[i + 1]) for i in range(N - 1)]
    flows.append(flows[-1])
    ref16 = (np.stack(flows) * np.float32(2.0)).astype(np.float16)
    cent = ref.calc_av_centroid(masks["av"], nframes)
    masked_ref = ref16.astype(np.float32) * masks["rv"]
    mode_ref = ref.angle_mode(masked_ref, nframes)
    rad, lng = ref.comp_magnitude(masked_ref, cent)
    _, _, rhi, rlo = ref.bidirectional_hist(rad, nframes)
    res = engine.analyze_clip(stored, masks["rv"], cent, nframes, 1, 99)
    kw = dict(smooth_fraction=0.2, pad_len=10)
    sys_ref, dia_ref = ref.angle_detector_intervals(mode_ref, **kw)
    sys_gpu, dia_gpu = ref.angle_detector_intervals(res["ang_mode"], **kw)
    assert [list(map(int, s)) for s in sys_gpu] == [list(map(int, s)) for s in sys_ref]
    assert [list(map(int, s)) for s in dia_gpu] == [list(map(int, s)) for s in dia_ref]
    pk = dict(smooth_fraction=0.3, pad_len=10, peak_thres=0.5, min_dist=3)
    a = ref.radlong_peak_indices(rhi, rlo, sys_ref, nframes, **pk)
    b = ref.radlong_peak_indices(res["rad_hi"], res["rad_lo"], sys_gpu, nframes, **pk)
    assert a == b


def test_frame_prep_matches_host_formula(engine):
    """img2uint8(rgb2gray(frame)) (calculate_optical_flow.py:588, optical_flow_utils.py:30-31) on the GPU"""
    from oracle.frame_prep_ref import prepare_frames
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (5, 70, 90, 3), dtype=np.uint8)
    rgb[1] = rgb[1] // 3 + 40                        # min > 0: the (sic) division by max, not by the range
    rgb[2] = 0                                       # all-black frame
    gray_rgb = np.repeat(rng.integers(0, 256, (1, 70, 90, 1), dtype=np.uint8), 3, axis=-1)
    rgb[3] = gray_rgb[0]                             # gray2rgb input (R = G = B)
    want = prepare_frames(rgb)
    got = engine.prepare_frames(rgb)
    assert got.dtype == np.uint8 and got.shape == (5, 70, 90)
    with np.errstate(all="ignore"):
        assert np.array_equal(got, want)


def test_process_frames_rgb_input(engine, clip):
    from oracle.frame_prep_ref import prepare_frames
    from tee_optical_flow_b200.flow import process_frames
    frames, masks = clip
    rgb = np.stack([frames[:6]] * 3, axis=-1)
    rgb[..., 1] = rgb[..., 1] // 2
    out = process_frames(rgb, {k: v[:6] for k, v in masks.items()}, engine=engine)
    gray = prepare_frames(rgb)
    _, want = engine.calc_clip(gray, want_f32=False, want_f16=True)
    assert np.array_equal(out['flow'], want)
    assert out['attrs']['units_converted'] is False and out['echo'].shape == (6,) + frames.shape[1:]


def _class_map(seed, N, H, W):
    """noisy SAM-like argmax map: two classes with holes, specks and frame-to-frame flicker"""
    from tee_optical_flow_b200.synth import make_masks
    rng = np.random.default_rng(seed)
    m = make_masks(seed, N, H, W, period=9.0)
    cm = np.zeros((N, H, W), np.uint8)
    cm[m["rv"][..., 0]] = 1
    cm[m["av"][..., 0]] = 2
    flick = rng.random((N, H, W))
    cm[flick < 0.04] = 0                                   # holes / drop-outs
    cm[(flick > 0.985) & (cm == 0)] = 1                    # specks
    cm[(flick > 0.97) & (flick <= 0.985) & (cm == 0)] = 2
    cm[3, 10:40, 10:60] = 1                                # a blob present in a single frame (voted away)
    cm[5:9, 60:75, 5:25] = 2                               # a second AV component, smaller than the main one
    return cm


def test_clean_mask_matches_reference(engine, ref):
    """clean_mask (calculate_optical_flow.py:113-182): temporal vote, binary_fill_holes, remove_small_objects, bkgd"""
    from tee_optical_flow_b200.config import OpticalFlowCalculationConfig
    from tee_optical_flow_b200.masks import MODE_CLASSES, clean_mask
    cm = _class_map(11, 14, 90, 120)
    cfg = OpticalFlowCalculationConfig(min_mask_size=60)
    got = clean_mask(engine, cm, mode='RVIO_2class', config=cfg)
    want = ref.clean_mask(cm, MODE_CLASSES['RVIO_2class'], min_size=60)
    assert set(got) == {'rv', 'av', 'bkgd'}
    for k in want:
        assert got[k].shape == want[k].shape == (14, 90, 120, 2) and got[k].dtype == bool
        assert np.array_equal(got[k], want[k]), k
    assert clean_mask(engine, cm, mode='nope') is None
    assert want['rv'][3, 20, 30, 0] == False and got['rv'].any()


def test_av_centroid_matches_reference(engine, ref):
    """calc_AV_centroid (analysis.py:39-86): largest 8-connected component, fallbacks, Savitzky-Golay"""
    from tee_optical_flow_b200.masks import MODE_CLASSES, calc_AV_centroid
    cm = _class_map(12, 16, 90, 120)
    masks = ref.clean_mask(cm, MODE_CLASSES['RVIO_2class'], min_size=30)
    av = masks['av'].copy()
    av[0] = False                                          # first frame empty -> image centre
    av[7] = False                                          # empty in the middle -> copy previous
    for filt in (False, True):
        got = np.asarray(calc_AV_centroid(engine, av, 14, filter=filt))
        want = ref.calc_av_centroid(av, 14, do_filter=filt)
        assert got.shape == (14, 2)
        assert np.array_equal(got, np.asarray(want)), filt
