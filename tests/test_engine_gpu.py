"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle and the golden vectors.

Bar (integer-free floating-point path): the engine evaluates the same IEEE float32 operations in the same order
as the oracle, so the flow must be BIT-IDENTICAL to the oracle run with the float64 error sum (err_mode=1, the
engine's reduction) and the per-level iteration counters must be equal.  Against the OpenCV-faithful serial
float32 error sum (err_mode=0) the north-star tolerance applies: mean end-point error <= 1e-2 px.
"""
import ast

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPE_TOL = 1e-2   # px, north_star


@pytest.fixture(scope="module")
def engine():
    from tee_optical_flow_b200.engine import TVL1Engine
    e = TVL1Engine(device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def pairs(golden_dir):
    return np.load(golden_dir / "tvl1_pairs.npz")


def _fresh(**kw):
    from tee_optical_flow_b200.engine import TVL1Engine
    return TVL1Engine(device=0, **kw)


def _epe(a, b):
    return np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1))


NAMES = dict(lambda_="lambda_", tau="tau", theta="theta", nscales="nscales", warps="warps", epsilon="epsilon",
             inner="inner_iterations", outer="outer_iterations")


@pytest.mark.parametrize("case", ["u8_default", "u8_fast", "f32_default", "u8_params", "u8_tiny_pyramid_stop"])
def test_golden_pairs_bitexact(pairs, case):
    params = ast.literal_eval(str(pairs[f"{case}__params"]))
    with _fresh(**{NAMES[k]: v for k, v in params.items()}) as eng:
        flow = eng.calc(pairs[f"{case}__I0"], pairs[f"{case}__I1"])
        counters, info = eng.last_counters()
    want = pairs[f"{case}__flow_em1"]
    assert flow.dtype == np.float32 and flow.shape == want.shape
    assert np.all(flow == want), f"mean EPE {_epe(flow, want).mean():.3e}"
    gold_c = pairs[f"{case}__counters_em1"]
    assert np.array_equal(counters[0], gold_c[:counters.shape[1]])
    # OpenCV-faithful error sum: tolerance of the north star
    assert _epe(flow, pairs[f"{case}__flow_em0"]).mean() <= EPE_TOL


@pytest.mark.parametrize("shape", [(16, 16), (17, 23), (33, 65), (64, 64), (65, 129), (100, 37),
                                   (24, 1100), (20, 2100)])   # the last two: plane pitch 2048 / 4096
def test_ragged_sizes_vs_oracle(oracle, shape):
    """tile-boundary cases: sizes below / at / just above the 64x16 tile, odd sizes, pyramid stop < 16 px"""
    from tee_optical_flow_b200.synth import make_clip
    H, W = shape
    fr = make_clip(seed=H * 1000 + W, n_frames=2, H=H, W=W, peak_disp=2.0, period=8.0)
    with _fresh() as eng:
        flow = eng.calc(fr[0], fr[1])
        counters, _ = eng.last_counters()
    om = oracle.OracleDualTVL1(err_mode=1)
    ref = om.calc(fr[0], fr[1])
    assert np.all(flow == ref), f"{shape}: mean EPE {_epe(flow, ref).mean():.3e}"
    assert counters.shape[1] == om.last_nscales
    assert np.array_equal(counters[0], om.last_counters[:om.last_nscales])


def test_noise_images_hit_borders_and_iteration_caps(oracle):
    """white-noise frames: large incoherent flow, bicubic taps outside the image, inner/outer caps reached"""
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (48, 80), dtype=np.uint8)
    b = rng.integers(0, 256, (48, 80), dtype=np.uint8)
    with _fresh(nscales=3, warps=2, inner_iterations=7, outer_iterations=3) as eng:
        flow = eng.calc(a, b)
        counters, _ = eng.last_counters()
    om = oracle.OracleDualTVL1(nscales=3, warps=2, innnerIterations=7, outerIterations=3, err_mode=1)
    ref = om.calc(a, b)
    assert np.all(flow == ref)
    assert np.array_equal(counters[0], om.last_counters[:3])
    assert counters[0, :, 0].max() == 2 * 3 * 7      # caps reached: warps * outer * inner


@pytest.mark.parametrize("median", [1, 3, 5])
def test_median_sizes(oracle, median):
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=21, n_frames=2, H=70, W=90, peak_disp=4.0, period=8.0)
    with _fresh(median_filtering=median) as eng:
        flow = eng.calc(fr[0], fr[1])
    ref = oracle.OracleDualTVL1(medianFiltering=median, err_mode=1).calc(fr[0], fr[1])
    assert np.all(flow == ref)


def test_identical_frames_zero_flow(engine):
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=5, n_frames=1, H=96, W=128)
    flow = engine.calc(fr[0], fr[0])
    assert np.all(flow == 0)
    counters, _ = engine.last_counters()
    assert counters[0, :, 0].tolist() == [5] * 5     # first iteration has error 0 -> one iteration per warp


def test_clip_with_refill_matches_pairwise(oracle):
    """more pairs than slots: finished slots are refilled from the work counter; results do not depend on it"""
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=8, n_frames=10, H=80, W=96, peak_disp=5.0, period=6.0)
    with _fresh(max_slots=3) as eng:
        f3, h3 = eng.calc_clip(fr, duplicate_last=True, want_f16=True)
        c3, info = eng.last_counters()
    assert info["n_slots"] == 3 and info["n_pairs"] == 9
    with _fresh(max_slots=64) as eng:
        f64_, _ = eng.calc_clip(fr, duplicate_last=False)
        c64, _ = eng.last_counters()
    assert f3.shape == (10, 80, 96, 2) and f64_.shape == (9, 80, 96, 2)
    assert np.array_equal(f3[:9], f64_) and np.array_equal(c3, c64)
    assert np.array_equal(f3[9], f3[8])               # reference appends a copy of the last flow (:599)
    assert np.array_equal(h3, f3.astype(np.float16))  # fp16 pack is RNE like numpy astype (:403)
    om = oracle.OracleDualTVL1(err_mode=1)
    for i in (0, 4, 8):
        assert np.all(f3[i] == om.calc(fr[i], fr[i + 1]))


def test_device_path_equals_host_path_and_scale(engine):
    import torch
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=9, n_frames=5, H=64, W=80, peak_disp=3.0, period=8.0)
    host32, _ = engine.calc_clip(fr, out_scale=1.0)
    dev32, dev16 = engine.calc_clip(torch.from_numpy(fr).cuda(), out_scale=0.37, want_f16=True)
    scaled = host32 * np.float32(0.37)                # np.stack(flow_list) * conversion_factor (:600)
    assert np.array_equal(dev32.cpu().numpy(), scaled)
    assert np.array_equal(dev16.cpu().numpy(), scaled.astype(np.float16))


def test_sharded_pair_ranges_equal_unsharded(engine):
    """a clip split by pair range (1-frame overlap) gives the same bits as the whole clip"""
    import torch
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=10, n_frames=9, H=64, W=96, peak_disp=3.0, period=8.0)
    whole, _ = engine.calc_clip(fr, duplicate_last=False)
    d = torch.from_numpy(fr).cuda()
    lo, _ = engine.calc_pairs_device(d, np.arange(0, 4), np.arange(1, 5))
    hi, _ = engine.calc_pairs_device(d, np.arange(4, 8), np.arange(5, 9))
    assert np.array_equal(np.concatenate([lo.cpu().numpy(), hi.cpu().numpy()]), whole)


def test_determinism(engine):
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=12, n_frames=4, H=72, W=88, peak_disp=6.0, period=6.0)
    a, _ = engine.calc_clip(fr)
    b, _ = engine.calc_clip(fr)
    assert np.array_equal(a, b)


def test_translation_sign_convention(engine):
    from scipy.ndimage import gaussian_filter, shift
    rng = np.random.default_rng(3)
    base = gaussian_filter(rng.standard_normal((120, 160)), 3.0)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    I0 = base.astype(np.uint8)
    I1 = np.clip(np.rint(shift(base, (-1.0, 2.0), order=3, mode="reflect")), 0, 255).astype(np.uint8)
    inner = engine.calc(I0, I1)[20:-20, 20:-20]
    assert abs(inner[..., 0].mean() - 2.0) < 0.05 and abs(inner[..., 1].mean() + 1.0) < 0.05


def test_errors_follow_the_reference(engine):
    from tee_optical_flow_b200.exceptions import OpticalFlowCalculationError
    with pytest.raises(OpticalFlowCalculationError):
        engine.calc(np.zeros((32, 32), np.float64), np.zeros((32, 32), np.float64))
    with pytest.raises(OpticalFlowCalculationError):
        engine.calc(np.zeros((32, 32), np.uint8), np.zeros((32, 33), np.uint8))
    with pytest.raises(OpticalFlowCalculationError):
        engine.calc_clip(np.zeros((1, 32, 32), np.uint8))
    with pytest.raises(OpticalFlowCalculationError):
        engine.setMedianFiltering(7)
    assert engine.getMedianFiltering() == 5
    engine.setLambda(0.2)
    assert engine.getLambda() == 0.2
    engine.setLambda(0.15)


def test_full_size_600x800_pair_vs_oracle(engine, oracle):
    """BASELINE size: bit-identical to the oracle (float64 error sum); <= 1e-2 px mean EPE and max |d| stated
    against the OpenCV-faithful serial float32 error sum"""
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=0, n_frames=2, H=600, W=800)
    flow = engine.calc(fr[0], fr[1])
    counters, _ = engine.last_counters()
    o1 = oracle.OracleDualTVL1(err_mode=1)
    ref1 = o1.calc(fr[0], fr[1])
    assert np.all(flow == ref1)
    assert np.array_equal(counters[0], o1.last_counters)
    ref0 = oracle.OracleDualTVL1(err_mode=0).calc(fr[0], fr[1])
    epe = _epe(flow, ref0)
    print(f"600x800 vs serial-f32 oracle: mean EPE {epe.mean():.3e}, max |d| {np.abs(flow - ref0).max():.3e}")
    assert epe.mean() <= EPE_TOL


def test_full_size_clip_properties(engine):
    """64-frame 600x800 clip (BASELINE configs[1]) through the device path: size-independent properties"""
    import torch
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=1, n_frames=16, H=600, W=800)
    d = torch.from_numpy(fr).cuda()
    f32, f16 = engine.calc_clip(d, want_f16=True)
    counters, info = engine.last_counters()
    f32 = f32.cpu().numpy()
    assert np.isfinite(f32).all()
    assert np.array_equal(f32[-1], f32[-2])
    assert np.array_equal(f16.cpu().numpy(), f32.astype(np.float16))
    assert (counters[:, :, 2] == 5).all() and (counters[:, :, 0] >= 5).all()
    # reversed clip: pair (i+1 -> i) flow is roughly the negated forward flow inside the sector
    rev, _ = engine.calc_pairs_device(d, np.array([1], np.int32), np.array([0], np.int32))
    fwd = f32[0]
    m = np.abs(fwd).sum(-1) > 0.05
    assert np.abs(rev.cpu().numpy()[0][m] + fwd[m]).mean() < 0.1


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_exact_division_matches_ieee(engine, mode):
    """the shared-reciprocal division of the inner iteration is bit-identical to IEEE division (2^32 pairs)"""
    import ctypes as C
    bad = C.c_int64(-1)
    rc = engine._lib.teeflow_selftest_division(engine._h, mode, 1 << 32, 1234 + mode, C.byref(bad))
    assert rc == 0 and bad.value == 0


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fast_hypot_matches_double_precision_hypot(engine, mode):
    """the packed float-float hypot of the dual update never ACCEPTS a value that differs from
    (float)sqrt((double)a*a + (double)b*b) (2^31 operand pairs per mode); in the flow-gradient range it hands
    fewer than 0.1 % of the pairs to the exact form"""
    import ctypes as C
    bad, rej = C.c_int64(-1), C.c_int64(-1)
    n = 1 << 30
    rc = engine._lib.teeflow_selftest_hypot(engine._h, mode, n, 99 + mode, C.byref(bad), C.byref(rej))
    assert rc == 0 and bad.value == 0, (bad.value, rej.value)
    if mode == 0:
        assert rej.value < n // 1000, rej.value


@pytest.mark.parametrize("spec", [1.01, 2.0, 8.0])
def test_two_iteration_passes_change_nothing(oracle, spec):
    """temporal blocking of the inner loop (two iterations per pass, speculative exit test, spec_factor > 0) yields the
    same flow bits and the same per-level iteration counts as the oracle -- whatever the speculation threshold,
    including discarded passes"""
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=11, n_frames=4, H=150, W=203, peak_disp=6.0, period=8.0)
    with _fresh() as eng:
        eng._set("spec_factor", spec)
        flow, _ = eng.calc_clip(fr, duplicate_last=False)
        counters, info = eng.last_counters()
    assert info["double_steps"] > 0
    if spec < 1.5:
        assert info["double_steps_discarded"] > 0      # the low threshold must exercise the discard path
    om = oracle.OracleDualTVL1(err_mode=1)
    for i in range(3):
        ref = om.calc(fr[i], fr[i + 1])
        assert np.all(flow[i] == ref), f"pair {i}: mean EPE {_epe(flow[i], ref).mean():.3e}"
        assert np.array_equal(counters[i], om.last_counters[:counters.shape[1]])


def test_two_iteration_passes_full_size_equal_single_iteration_passes():
    """600x800: the flow bits and the executed iteration counts do not depend on the temporal-blocking option"""
    import torch
    from tee_optical_flow_b200.synth import make_clip
    fr = torch.from_numpy(make_clip(seed=5, n_frames=7, H=600, W=800)).cuda()
    out = []
    for spec in (0.0, 1.5):
        with _fresh() as eng:
            eng._set("spec_factor", spec)
            f32, _ = eng.calc_clip(fr)
            counters, info = eng.last_counters()
            out.append((f32.cpu().numpy(), counters, info))
    assert out[0][2]["double_steps"] == 0 and out[1][2]["double_steps"] > 0
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1], out[1][1])


def test_batch_of_clips_equals_per_clip():
    """BASELINE config 4 (one rank's share): several clips through one scheduler run, slots refilled across clips"""
    import torch
    from tee_optical_flow_b200.synth import make_clip
    clips = np.stack([make_clip(seed=s, n_frames=6, H=64, W=96, peak_disp=4.0, period=6.0) for s in (20, 21, 22)])
    with _fresh(max_slots=4) as eng:
        _, b16 = eng.calc_batch(torch.from_numpy(clips).cuda(), out_scale=1.25)
        b16 = b16.cpu().numpy()
        assert b16.shape == (3, 6, 64, 96, 2)
        for i in range(3):
            _, c16 = eng.calc_clip(clips[i], out_scale=1.25, want_f32=False, want_f16=True)
            assert np.array_equal(b16[i], c16)


def test_batch_cut_into_several_scheduler_runs_changes_nothing():
    """calc_batch bounds its pyramid workspace by cutting a large batch into scheduler runs (max_frames_per_run): same
    flows, and last_counters() returns the counters of all runs in batch order"""
    import torch
    from tee_optical_flow_b200.synth import make_clip
    clips = np.stack([make_clip(seed=s, n_frames=5, H=48, W=80, peak_disp=3.0, period=6.0) for s in (40, 41, 42, 43, 44)])
    d = torch.from_numpy(clips).cuda()
    with _fresh() as eng:
        whole32, whole16 = eng.calc_batch(d, want_f32=True, want_f16=True)
        c_whole, i_whole = eng.last_counters()
        cut32, cut16 = eng.calc_batch(d, want_f32=True, want_f16=True, max_frames_per_run=10)     # 2 clips per run
        c_cut, i_cut = eng.last_counters()
    assert torch.equal(whole32, cut32) and torch.equal(whole16, cut16)
    assert i_cut["scheduler_runs"] == 3 and i_cut["n_pairs"] == i_whole["n_pairs"] == 20
    assert np.array_equal(c_whole, c_cut)


def test_long_clip_config_1024_7scales_10warps(oracle):
    """BASELINE config 5 geometry: 1024x1024, nscales=7, warps=10 -- one pair against the oracle, bit for bit"""
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=30, n_frames=2, H=1024, W=1024, peak_disp=4.0, period=10.0)
    with _fresh(nscales=7, warps=10) as eng:
        assert eng.level_sizes(1024, 1024) == [(1024, 1024), (819, 819), (655, 655), (524, 524), (419, 419),
                                               (335, 335), (268, 268)]
        flow = eng.calc(fr[0], fr[1])
        counters, _ = eng.last_counters()
    om = oracle.OracleDualTVL1(nscales=7, warps=10, err_mode=1)
    ref = om.calc(fr[0], fr[1])
    assert np.all(flow == ref)
    assert np.array_equal(counters[0], om.last_counters)
    assert (counters[0, :, 2] == 10).all()


def test_asynchronous_device_run_equals_synchronous(engine):
    """teeflow_calc_clip_async: returns after enqueueing the pyramid and the one dataflow launch; later work on the
    stream sees complete results, teeflow_finish() delivers the verdict and the statistics"""
    import torch
    from tee_optical_flow_b200.exceptions import OpticalFlowCalculationError
    from tee_optical_flow_b200.synth import make_clip
    fr = torch.from_numpy(make_clip(seed=14, n_frames=6, H=96, W=128, peak_disp=4.0, period=8.0)).cuda()
    want32, want16 = engine.calc_clip(fr, out_scale=0.5, want_f16=True)
    c_sync, _ = engine.last_counters()
    got32, got16 = engine.calc_clip(fr, out_scale=0.5, want_f16=True, asynchronous=True)
    with pytest.raises(OpticalFlowCalculationError):      # a second run before finish() is refused
        engine.calc_clip(fr)
    follow_up = got32 * 2                                 # enqueued behind the solver on the same stream
    engine.finish()
    c_async, info = engine.last_counters()
    assert info["solver_launches"] == 1 and info["n_pairs"] == 5
    assert torch.equal(got32, want32) and torch.equal(got16, want16) and torch.equal(follow_up, want32 * 2)
    assert np.array_equal(c_sync, c_async)
    engine.finish()                                       # idempotent
