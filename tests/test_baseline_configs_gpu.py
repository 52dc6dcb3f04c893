"""GPU parity tests at the FULL sizes of BASELINE.json's configs, through the C ABI, against the CPU oracle.

  config 2  64-frame 600x800 uint8 clip, TV-L1 defaults: every one of the 63 pairs bit-identical to the oracle with
            the float64 error sum (flow AND per-level iteration counters); against the OpenCV-faithful serial float32
            error sum the north-star tolerance, mean EPE <= 1e-2 px, with max |d| printed.
  config 3  the same clip with RVIO_2class masks, saliency off AND on, radial / longitudinal decomposition: stored fp16
            flow, per-frame waveforms and the systole / diastole + e' / l' / a' frame indices with the REFERENCE'S
            DEFAULT downstream parameters (optical_flow/config.py:13-16, 75-82) equal to the oracle chain.
  config 5  1024x1024, 7 scales, 10 warps, WASE on: 4 pairs.

The oracle solves ~9 pairs/s on the box's host cores, so a whole clip costs seconds; its flows are computed once per
module.  Saliency parity is unpinned: that branch checks the engine against the restated oracle only.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPE_TOL = 1e-2          # px, north_star
H, W, N = 600, 800, 64


def _epe(a, b):
    return np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1))


def _oracle_clip(oracle, frames, err_mode, **kw):
    om = oracle.OracleDualTVL1(err_mode=err_mode, **kw)
    flows, counters = [], []
    for i in range(len(frames) - 1):
        flows.append(om.calc(frames[i], frames[i + 1]))
        counters.append(om.last_counters.copy())
    return np.stack(flows), np.stack(counters)


@pytest.fixture(scope="module")
def clip():
    from tee_optical_flow_b200.synth import make_clip, make_masks
    return make_clip(seed=0, n_frames=N, H=H, W=W), make_masks(0, N, H, W)


@pytest.fixture(scope="module")
def cycle_clip():
    """config 3 clip: same size and texture, motion period 24 frames.  On the period-32 clip the last systole run
    ends two frames before the end, a tail gap on which the reference's e' / l' / a' windows are empty and its pickers
    raise; with period 24 the cardiac cycles are complete and every stage of the downstream chain produces indices."""
    from tee_optical_flow_b200.synth import make_clip, make_masks
    return make_clip(seed=0, n_frames=N, H=H, W=W, period=24.0), make_masks(0, N, H, W, period=24.0)


@pytest.fixture(scope="module")
def engine():
    from tee_optical_flow_b200.engine import TVL1Engine
    e = TVL1Engine(device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def engine_flows(engine, clip):
    frames, _ = clip
    f32, _ = engine.calc_clip(frames, duplicate_last=False)
    counters, info = engine.last_counters()
    return f32, counters, info


@pytest.fixture(scope="module")
def oracle_flows_em0(oracle, clip):
    oracle.set_threads(0)
    return _oracle_clip(oracle, clip[0], 0)


def test_config2_whole_clip_bitexact_with_counters(engine_flows, oracle, clip):
    frames, _ = clip
    flow, counters, info = engine_flows
    assert info["n_pairs"] == N - 1 and flow.shape == (N - 1, H, W, 2)
    oracle.set_threads(0)
    ref, ref_c = _oracle_clip(oracle, frames, 1)
    bad = [i for i in range(N - 1) if not np.array_equal(flow[i].view(np.uint32), ref[i].view(np.uint32))]
    assert not bad, f"pairs {bad} differ from the oracle; first mean EPE {_epe(flow[bad[0]], ref[bad[0]]).mean():.3e}"
    assert np.array_equal(counters, ref_c[:, :counters.shape[1]])


def test_config2_schedulers_agree(engine_flows, clip):
    """the stepped scheduler (one launch per phase step) and the two-iteration passes produce the same bits and
    counters as the default dataflow kernel with single-iteration passes"""
    from tee_optical_flow_b200.engine import TVL1Engine
    frames, _ = clip
    flow, counters, _ = engine_flows
    with TVL1Engine(device=0) as eng:
        eng._set("stepped", 1)
        f2, _ = eng.calc_clip(frames[:9], duplicate_last=False)
        c2, info = eng.last_counters()
    assert info["solver_launches"] > 8
    assert np.array_equal(f2.view(np.uint32), flow[:8].view(np.uint32)) and np.array_equal(c2, counters[:8])
    with TVL1Engine(device=0) as eng:
        eng._set("spec_factor", 1.5)                     # two-iteration passes: same bits, same counters
        f3, _ = eng.calc_clip(frames[:9], duplicate_last=False)
        c3, info = eng.last_counters()
    assert info["double_steps"] > 0
    assert np.array_equal(f3.view(np.uint32), flow[:8].view(np.uint32)) and np.array_equal(c3, counters[:8])


def test_config2_tolerance_against_opencv_faithful_error_sum(engine_flows, oracle_flows_em0):
    flow, counters, _ = engine_flows
    ref0, ref0_c = oracle_flows_em0
    epe = _epe(flow, ref0)
    per_pair = epe.reshape(N - 1, -1).mean(1)
    flips = int((counters != ref0_c[:, :counters.shape[1]]).any(axis=(1, 2)).sum())
    print(f"config 2, 63 pairs vs serial-float32 oracle: mean EPE {epe.mean():.3e} px, worst pair {per_pair.max():.3e}, "
          f"max |d| {np.abs(flow - ref0).max():.3e} px, pairs whose exit decisions differ: {flips}")
    assert epe.mean() <= EPE_TOL and per_pair.max() <= EPE_TOL


def _downstream_equal(engine, flow16_gpu, flow16_ref, masks, frame_rate):
    """engine chain (GPU reductions + product waveform pipeline) vs oracle chain (numpy restatement of the reference)
    with the reference's default downstream configs"""
    from oracle import downstream_ref as R
    from tee_optical_flow_b200 import waveforms as Wv
    from tee_optical_flow_b200.masks import calc_AV_centroid
    nframes = N - 2                                       # OpticalFlowDataset.nframes = attrs['nframes'] - 2
    want = R.clip_indices(flow16_ref, masks["rv"], masks["av"], nframes)
    cent = np.asarray(calc_AV_centroid(engine, masks["av"], nframes, filter=True))
    assert np.array_equal(cent, want["centroids"])
    res = engine.analyze_clip(flow16_gpu, masks["rv"], cent, nframes, 1, 99)
    got = Wv.indices_of(Wv.clip_waveform_indices(res, nframes, frame_rate=frame_rate, strict=False))
    keys = ("sys_frames", "dia_frames", "single", "radial", "longitudinal")
    return {k: got[k] for k in keys}, {k: want[k] for k in keys}, res, want["waveforms"]


def test_config3_no_saliency_layout_waveforms_and_indices(engine, cycle_clip, oracle):
    """process_frames(no_saliency=True, RVIO_2class masks) at full size: stored fp16 flow == oracle flow * conversion
    -> fp16; waveforms and indices from the engine chain == the oracle chain on the OpenCV-faithful oracle's flow"""
    from oracle.frame_prep_ref import prepare_frames
    from tee_optical_flow_b200.flow import process_frames
    frames, masks = cycle_clip
    rgb = np.stack([frames] * 3, axis=-1)                 # a greyscale DICOM after gray2rgb (:536)
    prepared = prepare_frames(rgb)                        # img2uint8(rgb2gray(.)): the solver's input (:588)
    ps, fr = 0.05, 40.0
    out = process_frames(rgb, masks, pixel_spacing=ps, frame_rate=fr, mode='RVIO_2class', bkgd_comp='none',
                         no_saliency=True, engine=engine)
    assert out['attrs']['nframes'] == N and out['attrs']['labels'] == ['rv', 'av', 'bkgd']
    assert out['attrs']['units_converted'] is True and out['flow'].dtype == np.float16
    oracle.set_threads(0)
    ref0, _ = _oracle_clip(oracle, prepared, 0)
    cf = np.float32(ps * fr)
    ref16 = (np.concatenate([ref0, ref0[-1:]]) * cf).astype(np.float16)
    epe = _epe(out['flow'].astype(np.float32)[:-1] / cf, ref0)
    print(f"config 3 (saliency off): mean EPE {epe.mean():.3e} px (includes the fp16 rounding of the stored flow)")
    assert epe.mean() <= EPE_TOL
    got, want, res, wf = _downstream_equal(engine, out['flow'], ref16, masks, fr)
    print("config 3 (saliency off) indices:", got)
    assert got == want
    assert any('raises' not in got[k] for k in ('single', 'radial', 'longitudinal')), "degenerate clip: every picker raised"
    if np.array_equal(out['flow'], ref16):                # no exit decision flipped: the waveforms are then identical too
        assert np.array_equal(res["mag_hi"], wf["mag_hi"].astype(np.float32))
        assert np.array_equal(res["rad_hi"], wf["rad_hi"]) and np.array_equal(res["long_lo"], wf["long_lo"])
        assert np.array_equal(res["ang_mode"], wf["ang_mode"].astype(np.float32))


def test_config3_saliency_on(engine, cycle_clip, oracle):
    """process_frames(no_saliency=False): the GPU saliency stage + TV-L1 on its float32 maps vs the restated saliency
    oracle + TV-L1 oracle; flow within the north-star tolerance, downstream indices equal.  (Saliency composition:
    parity unpinned, DESIGN.md -- this pins the engine to the restatement only.)"""
    from oracle import saliency_ref as S
    from tee_optical_flow_b200.exceptions import SaliencyParityWarning
    from tee_optical_flow_b200.flow import process_frames
    frames, masks = cycle_clip
    rng = np.random.default_rng(5)
    rgb = np.stack([frames, (frames.astype(np.float32) * 0.8).astype(np.uint8), 255 - frames], axis=-1)
    rgb[..., 1] += rng.integers(0, 3, frames.shape, dtype=np.uint8)
    with pytest.warns(SaliencyParityWarning):
        out = process_frames(rgb, masks, pixel_spacing=0.05, frame_rate=40.0, mode='RVIO_2class', no_saliency=False,
                             engine=engine)
    assert out['_engine_info']['saliency_parity'] == 'unpinned'
    sal = np.stack([S.compute_saliency(rgb[i]) for i in range(N)])
    assert np.array_equal(engine.compute_saliency(rgb), sal)
    oracle.set_threads(0)
    ref0, _ = _oracle_clip(oracle, sal, 0)
    cf = np.float32(0.05 * 40.0)
    ref16 = (np.concatenate([ref0, ref0[-1:]]) * cf).astype(np.float16)
    epe = _epe(out['flow'].astype(np.float32)[:-1] / cf, ref0)
    print(f"config 3 (saliency on): mean EPE {epe.mean():.3e} px (includes the fp16 rounding of the stored flow)")
    assert epe.mean() <= EPE_TOL
    got, want, _, _ = _downstream_equal(engine, out['flow'], ref16, masks, 40.0)
    print("config 3 (saliency on) indices:", got)
    assert got == want


def test_config5_1024_7scales_10warps_wase(oracle):
    """BASELINE config 5 geometry with WASE background compensation on: 4 pairs at 1024x1024"""
    from oracle import downstream_ref as R
    from tee_optical_flow_b200.engine import TVL1Engine
    from tee_optical_flow_b200.synth import make_clip, make_masks
    n = 5
    fr = make_clip(seed=30, n_frames=n, H=1024, W=1024, peak_disp=4.0, period=10.0)
    masks = make_masks(30, n, 1024, 1024, period=10.0)
    oracle.set_threads(0)
    ref1, ref1_c = _oracle_clip(oracle, fr, 1, nscales=7, warps=10)
    ref0, _ = _oracle_clip(oracle, fr, 0, nscales=7, warps=10)
    with TVL1Engine(device=0, nscales=7, warps=10) as eng:
        plain, _ = eng.calc_clip(fr, duplicate_last=False)
        counters, _ = eng.last_counters()
        eng.set_wase_masks(masks["bkgd"])
        comp, comp16 = eng.calc_clip(fr, out_scale=1.5, duplicate_last=True, want_f16=True)
        bgs = eng.last_backgrounds()
    assert np.array_equal(plain.view(np.uint32), ref1.view(np.uint32))
    assert counters.shape[1] == 7 and np.array_equal(counters, ref1_c[:, :7]) and (counters[:, :, 2] == 10).all()
    epe = _epe(plain, ref0)
    print(f"config 5, 4 pairs vs serial-float32 oracle: mean EPE {epe.mean():.3e} px, max |d| {np.abs(plain - ref0).max():.3e}")
    assert epe.mean() <= EPE_TOL
    for i in range(n - 1):
        want_bg = R.wase_background(plain[i], masks["bkgd"])           # np.mean(masked_flow[masked_flow != 0])
        assert abs(bgs[i] - want_bg) <= 2e-6 * abs(want_bg) + 1e-12
        assert np.array_equal(comp[i], (plain[i] - bgs[i]) * np.float32(1.5))
    assert np.array_equal(comp[n - 1], comp[n - 2]) and np.array_equal(comp16, comp.astype(np.float16))
