"""Saliency input stage (StaticSaliencyFineGrained restatement, PARITY UNPINNED for the composition):
CPU: the closed forms the GPU kernels use (fixed-point grey, binomial 5x5 blur, float32 integral) against the
genuine cv2 primitives of this image; GPU: the kernels against oracle/saliency_ref.py, bit for bit."""
import numpy as np
import pytest

from oracle import saliency_ref as S


def _images():
    from tee_optical_flow_b200.synth import make_clip
    rng = np.random.default_rng(3)
    clip = make_clip(seed=2, n_frames=3, H=150, W=203, peak_disp=4.0, period=8.0)
    rgb = np.repeat(clip[..., None], 3, -1)
    rgb[..., 1] = np.roll(rgb[..., 1], 3, axis=2)          # make the channels differ
    noise = rng.integers(0, 256, (3, 150, 203, 3), dtype=np.uint8)
    flat = np.zeros((3, 150, 203, 3), np.uint8); flat[1] = 255; flat[2, 40:90, 60:120] = 200
    return np.concatenate([rgb, noise, flat])


def test_closed_forms_match_genuine_cv2_primitives():
    cv2 = pytest.importorskip("cv2")
    for img in _images():
        g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(S.gray_bgr2gray(img), g)
        b = cv2.GaussianBlur(g, (5, 5), 0)
        assert np.array_equal(S.gaussian5(g), b)
        assert np.array_equal(S.integral_f32(b), cv2.integral(b, sdepth=cv2.CV_32F))
        assert np.array_equal(S.fine_grained_u8(img, use_cv2=False), S.fine_grained_u8(img, use_cv2=True))


def test_saliency_map_properties():
    imgs = _images()
    m = S.compute_saliency(imgs[0])
    assert m.dtype == np.float32 and m.shape == imgs[0].shape[:2] and m.min() >= 0 and m.max() == 1.0
    assert not S.fine_grained_u8(imgs[6]).any()           # a constant image has no centre-surround contrast
    assert not S.fine_grained_u8(imgs[7]).any()


@pytest.mark.gpu
def test_gpu_saliency_bitexact_vs_oracle():
    from tee_optical_flow_b200.engine import TVL1Engine
    imgs = _images()
    with TVL1Engine(device=0) as eng:
        got = eng.compute_saliency(imgs)
        got_u8 = eng.compute_saliency(imgs, return_u8=True)
    for i, img in enumerate(imgs):
        ref_u8 = S.fine_grained_u8(img)
        assert np.array_equal(got_u8[i], ref_u8), f"frame {i}: {(got_u8[i] != ref_u8).sum()} pixels differ"
        assert np.array_equal(got[i].view(np.uint32), S.compute_saliency(img).view(np.uint32))


@pytest.mark.gpu
def test_gpu_saliency_full_size_and_process_frames():
    """600x800, more frames than one scratch chunk; process_frames(no_saliency=False) feeds the maps to the solver"""
    from tee_optical_flow_b200.engine import TVL1Engine
    from tee_optical_flow_b200.flow import process_frames
    from tee_optical_flow_b200.synth import make_clip
    from oracle import tvl1_oracle as O
    clip = make_clip(seed=4, n_frames=18, H=600, W=800)
    rgb = np.repeat(clip[..., None], 3, -1)
    with TVL1Engine(device=0) as eng:
        sal = eng.compute_saliency(rgb)
    for i in (0, 16, 17):
        assert np.array_equal(sal[i].view(np.uint32), S.compute_saliency(rgb[i]).view(np.uint32))
    small = rgb[:3, 200:320, 300:460]
    res = process_frames(small, {}, mode='otsu', no_saliency=False)
    ref = O.OracleDualTVL1(err_mode=1).calc(S.compute_saliency(small[0]), S.compute_saliency(small[1]))
    assert np.array_equal(res['flow'][0], ref.astype(np.float16))
    assert res['attrs']['no_saliency'] is False
