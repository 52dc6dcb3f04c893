"""CPU tests: the C oracle against the committed golden vectors (genuine cv2 primitives; cv2-composed solver)."""
import ast

import numpy as np
import pytest

f32 = np.float32


def _eq(a, b):
    """numeric equality (so that -0.0 == +0.0), every element"""
    return a.shape == b.shape and bool(np.all(a == b))


@pytest.fixture(scope="module")
def prim(golden_dir):
    return np.load(golden_dir / "primitives_cv2.npz")


@pytest.fixture(scope="module")
def pairs(golden_dir):
    return np.load(golden_dir / "tvl1_pairs.npz")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_resize_matches_cv2_bitexact(oracle, prim, tag):
    img = prim[f"img_{tag}"]
    down = oracle.resize_by_factor(img, 0.8)
    assert _eq(down, prim[f"down_{tag}"])
    up = oracle.resize_linear(prim[f"down_{tag}"], img.shape)
    assert _eq(up, prim[f"up_{tag}"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_remap_matches_cv2_bitexact(oracle, prim, tag):
    out = oracle.remap_cubic(prim[f"img_{tag}"], prim[f"mapx_{tag}"], prim[f"mapy_{tag}"])
    assert _eq(out, prim[f"remap_{tag}"])


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("k", [3, 5])
def test_median_matches_cv2_bitexact(oracle, prim, tag, k):
    assert _eq(oracle.median_blur(prim[f"img_{tag}"], k), prim[f"median{k}_{tag}"])


def test_median_networks_zero_one(oracle):
    """zero-one principle: a comparator network selects the median of every input iff it does for all 0/1 inputs"""
    assert oracle.lib().oracle_median_network_failures(9) == 0
    assert oracle.lib().oracle_median_network_failures(25) == 0


def test_pyramid_sizes(oracle):
    """dsize = cvRound(size * 0.8): 600x800 -> 480x640 -> 384x512 -> 307x410 -> 246x328"""
    import ctypes as C
    h, w = 600, 800
    got = []
    for _ in range(4):
        dh, dw = C.c_int(), C.c_int()
        oracle.lib().oracle_scaled_size(h, w, 0.8, C.byref(dh), C.byref(dw))
        h, w = dh.value, dw.value
        got.append((h, w))
    assert got == [(480, 640), (384, 512), (307, 410), (246, 328)]


def test_gradient_boundaries(oracle):
    rng = np.random.default_rng(0)
    a = rng.random((7, 9)).astype(f32)
    dx, dy = oracle.centered_gradient(a)
    assert dx[3, 0] == f32(0.5) * (a[3, 1] - a[3, 0])
    assert dx[3, 8] == f32(0.5) * (a[3, 8] - a[3, 7])
    assert dy[0, 4] == f32(0.5) * (a[1, 4] - a[0, 4])
    assert dy[6, 4] == f32(0.5) * (a[6, 4] - a[5, 4])
    assert dx[2, 4] == f32(0.5) * (a[2, 5] - a[2, 3])


CASES = ["u8_default", "u8_fast", "f32_default", "u8_params", "u8_tiny_pyramid_stop"]


def _model(oracle, params, em):
    m = oracle.OracleDualTVL1(err_mode=em)
    names = dict(lambda_="setLambda", tau="setTau", theta="setTheta", nscales="setScalesNumber",
                 warps="setWarpingsNumber", epsilon="setEpsilon", inner="setInnerIterations",
                 outer="setOuterIterations", scale_step="setScaleStep", median="setMedianFiltering")
    for k, v in params.items():
        getattr(m, names[k])(v)
    return m


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("em", [0, 1])
def test_solver_matches_cv2_composition_bitexact(oracle, pairs, case, em):
    params = ast.literal_eval(str(pairs[f"{case}__params"]))
    m = _model(oracle, params, em)
    flow = m.calc(pairs[f"{case}__I0"], pairs[f"{case}__I1"])
    want = pairs[f"{case}__flow_em{em}"]
    assert flow.shape == want.shape and flow.dtype == np.float32
    assert _eq(flow, want), f"max |d| = {np.abs(flow - want).max()}"
    assert np.array_equal(m.last_counters, pairs[f"{case}__counters_em{em}"])


def test_pyramid_stops_below_16px(oracle, pairs):
    m = _model(oracle, dict(nscales=6), 0)
    m.calc(pairs["u8_tiny_pyramid_stop__I0"], pairs["u8_tiny_pyramid_stop__I1"])
    # 40x44 -> 32x35 -> 26x28 -> 21x22 -> 17x18 -> 14x14 (<16: dropped)
    assert m.last_nscales == 5
    assert m.last_counters[5].tolist() == [0, 0, 0]


def test_identical_frames_give_zero_flow(oracle):
    from tee_optical_flow_b200.synth import make_clip
    fr = make_clip(seed=5, n_frames=1, H=64, W=80)
    m = oracle.OracleDualTVL1()
    flow = m.calc(fr[0], fr[0])
    assert np.all(flow == 0)
    # first inner iteration has error 0 -> immediate exit: one iteration per warp
    assert m.last_counters[:, 0].tolist() == [5] * 5


def test_translation_sign_convention(oracle):
    """I1(x + u) ~= I0(x): content moving by (+2, -1) px gives flow[...,0] ~ +2, flow[...,1] ~ -1"""
    from scipy.ndimage import gaussian_filter, shift
    rng = np.random.default_rng(3)
    base = gaussian_filter(rng.standard_normal((120, 160)), 3.0)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    I0 = base.astype(np.uint8)
    I1 = np.clip(np.rint(shift(base, (-1.0, 2.0), order=3, mode="reflect")), 0, 255).astype(np.uint8)
    flow = oracle.OracleDualTVL1().calc(I0, I1)
    inner = flow[20:-20, 20:-20]
    assert abs(inner[..., 0].mean() - 2.0) < 0.05
    assert abs(inner[..., 1].mean() + 1.0) < 0.05


def test_recovers_the_synthetic_motion_field(oracle):
    """Physical sanity check that does not depend on any OpenCV build: the benchmark clips are a texture advected by an
    analytic displacement field d_t (frame_t(x) = texture(x + d_t(x))), so the flow from frame t to t+1 is d_t - d_(t+1)
    to first order.  The restated solver recovers it to a few hundredths of a pixel inside the ultrasound sector."""
    from scipy.ndimage import binary_erosion
    from tee_optical_flow_b200.synth import make_clip, sector_mask
    H, W = 240, 320
    fr, truth = make_clip(seed=3, n_frames=4, H=H, W=W, peak_disp=4.0, period=12.0, return_truth=True)
    flow = oracle.OracleDualTVL1(err_mode=0).calc(fr[1], fr[2])
    dx, dy = truth[1][0] - truth[2][0], truth[1][1] - truth[2][1]
    inside = binary_erosion(sector_mask(H, W), iterations=12)
    epe = np.sqrt((flow[..., 0] - dx) ** 2 + (flow[..., 1] - dy) ** 2)[inside]
    assert np.hypot(dx, dy)[inside].mean() > 0.4          # there is motion to recover
    assert epe.mean() < 0.06 and np.percentile(epe, 99) < 0.3


def test_rejects_bad_input(oracle):
    m = oracle.OracleDualTVL1()
    with pytest.raises(ValueError):
        m.calc(np.zeros((8, 8), np.uint8), np.zeros((8, 9), np.uint8))
    with pytest.raises(ValueError):
        m.calc(np.zeros((8, 8), np.float64), np.zeros((8, 8), np.float64))
