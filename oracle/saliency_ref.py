"""TEST INFRASTRUCTURE -- CPU restatement of cv2.saliency.StaticSaliencyFineGrained (opencv_contrib
modules/saliency/src/staticSaliencyFineGrained.cpp), the input stage of the reference when no_saliency=False
(optical_flow/calculate_optical_flow.py:560, :586).  Only tests/, __graft_entry__.smoke() and bench.py may import it.

PARITY UNPINNED: the `saliency` module is not in this image's cv2 (opencv-python-headless 4.13) and its source is
not under /root/reference; the COMPOSITION below is restated from the published source as recalled (Montabone &
Soto's fine-grained saliency: grey image, two 5x5 Gaussian blurs, integral image, on/off centre-surround
differences at six neighbourhood sizes, mixing).  The PRIMITIVES are the genuine ones of this image's cv2
(cvtColor, GaussianBlur, integral), and tests/test_saliency.py pins the closed forms the GPU uses against them.

Quirks kept on purpose: the reference hands RGB frames to a BGR2GRAY conversion; (uchar) casts truncate; a sum of the
normalised on and off responses above 255 wraps like the x86 cast (low byte); empty on/off maxima (division by
zero in the C++ code, undefined there) give 0 here.
"""
from __future__ import annotations

import numpy as np

NEIGHBORHOODS = (3 * 4, 3 * 4 * 2, 3 * 4 * 2 * 2, 7 * 4, 7 * 4 * 2, 7 * 4 * 2 * 2)
f32 = np.float32


def gray_bgr2gray(img_u8: np.ndarray) -> np.ndarray:
    """cvtColor(COLOR_BGR2GRAY) for uint8 (cv2 4.13: 15-bit fixed point) on whatever channel order comes in."""
    c0, c1, c2 = (img_u8[..., i].astype(np.int64) for i in range(3))
    return ((c0 * 3735 + c1 * 19235 + c2 * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def gaussian5(gray_u8: np.ndarray) -> np.ndarray:
    """GaussianBlur(Size(5,5), 0) for uint8: kernel [1 4 6 4 1]/16 per axis, BORDER_REFLECT_101, one rounding."""
    w = np.array([1, 4, 6, 4, 1], np.int64)
    H, W = gray_u8.shape
    p = np.pad(gray_u8.astype(np.int64), 2, mode="reflect")
    s = np.zeros((H, W), np.int64)
    for i in range(5):
        for j in range(5):
            s += w[i] * w[j] * p[i:i + H, j:j + W]
    return ((s + 128) >> 8).astype(np.uint8)


def integral_f32(gray_u8: np.ndarray) -> np.ndarray:
    """integral(src, CV_32F): sum[y+1][x+1] = sum[y][x+1] + (row prefix of row y up to x), float32, serial in y."""
    H, W = gray_u8.shape
    out = np.zeros((H + 1, W + 1), f32)
    for y in range(H):
        out[y + 1, 1:] = out[y, 1:] + np.cumsum(gray_u8[y].astype(f32), dtype=f32)
    return out


def _scaled(integ: np.ndarray, gray: np.ndarray, n: int):
    H, W = gray.shape
    yy, xx = np.mgrid[0:H, 0:W]
    p1x = np.clip(xx - n + 1, 0, W); p1y = np.clip(yy - n + 1, 0, H)
    p2x = np.clip(xx + n + 1, 0, W); p2y = np.clip(yy + n + 1, 0, H)
    v = ((integ[p2y, p2x] + integ[p1y, p1x]) - integ[p2y, p1x]) - integ[p1y, p2x]          # float32, this order
    g = gray.astype(f32)
    v = (v - g) / ((p2x - p1x) * (p2y - p1y) - 1).astype(f32)
    mean_on = g - v
    mean_off = v - g
    on = np.where(mean_on > 0, mean_on, f32(0)).astype(np.uint8)      # (uchar) truncation, values <= 255
    off = np.where(mean_off > 0, mean_off, f32(0)).astype(np.uint8)
    return on, off


def _norm255(num: np.ndarray, den: int) -> np.ndarray:
    """(uchar)(255. * (float)(num / (float)den)); den == 0 is undefined in the C++ code -> 0"""
    if den == 0:
        return np.zeros(num.shape, np.uint8)
    q = num.astype(f32) / f32(den)
    return (np.float64(255.0) * q.astype(np.float64)).astype(np.int64).astype(np.uint8)


def fine_grained_u8(img_u8: np.ndarray, use_cv2: bool = True) -> np.ndarray:
    """calcIntensityChannel: (H,W,3) or (H,W) uint8 -> (H,W) uint8 intensity conspicuity map"""
    if use_cv2:
        import cv2
        gray = cv2.cvtColor(img_u8, cv2.COLOR_BGR2GRAY) if img_u8.ndim == 3 else img_u8.copy()
        gray = cv2.GaussianBlur(gray, (5, 5), 0)
        gray = cv2.GaussianBlur(gray, (5, 5), 0)
        integ = cv2.integral(gray, sdepth=cv2.CV_32F)
    else:
        gray = gray_bgr2gray(img_u8) if img_u8.ndim == 3 else img_u8.copy()
        gray = gaussian5(gaussian5(gray))
        integ = integral_f32(gray)
    sum_on = np.zeros(gray.shape, np.uint16)
    sum_off = np.zeros(gray.shape, np.uint16)
    for n in NEIGHBORHOODS:
        on, off = _scaled(integ, gray, n)
        sum_on += on
        sum_off += off
    on8 = _norm255(sum_on, int(sum_on.max()))
    off8 = _norm255(sum_off, int(sum_off.max()))
    max_val = max(int(on8.max()), int(off8.max()))
    if max_val == 0:
        return np.zeros(gray.shape, np.uint8)
    s = (on8.astype(np.int64) + off8.astype(np.int64)).astype(f32)
    v = np.float64(255.0) * s.astype(np.float64) / np.float64(f32(max_val))
    return (v.astype(np.int64) & 255).astype(np.uint8)


def compute_saliency(img_u8: np.ndarray, use_cv2: bool = True) -> np.ndarray:
    """computeSaliency: float32 map in [0, 1] (dst.convertTo(saliencyMap, CV_32F, 1/255.f))"""
    return fine_grained_u8(img_u8, use_cv2).astype(f32) * f32(1.0 / 255.0)
