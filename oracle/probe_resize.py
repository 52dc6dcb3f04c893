"""Probe: float32 model of cv2.resize(INTER_LINEAR) (fx/fy form and dsize form) vs real cv2, bit level."""
import numpy as np, cv2
f32 = np.float32

def lin_coeffs(ssz, dsz, scale, clamp=True):
    d = np.arange(dsz, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx.astype(f32)
    if not clamp:   # vertical direction: weights kept, row indices clipped by the caller
        return sx, (f32(1) - fx).astype(f32), fx.astype(f32)
    lo = sx < 0
    fx[lo] = 0; sx[lo] = 0
    hi = sx >= ssz - 1
    fx[hi] = 0; sx[hi] = ssz - 1
    return sx, (f32(1) - fx).astype(f32), fx.astype(f32)

def resize_model(src, dW, dH, scale_x, scale_y, variant):
    H, W = src.shape
    sx, ax0, ax1 = lin_coeffs(W, dW, scale_x)
    sy, ay0, ay1 = lin_coeffs(H, dH, scale_y, clamp=False)
    sx1 = np.minimum(sx + 1, W - 1); sy1 = np.clip(sy + 1, 0, H - 1); sy = np.clip(sy, 0, H - 1)
    if variant == "plain":     # h: S0*a0 + S1*a1 ; v: r0*b0 + r1*b1
        hr = src[:, sx] * ax0[None, :] + src[:, sx1] * ax1[None, :]
        out = hr[sy, :] * ay0[:, None] + hr[sy1, :] * ay1[:, None]
    elif variant == "fma_v":   # vertical with fma(r1,b1, r0*b0)
        hr = src[:, sx] * ax0[None, :] + src[:, sx1] * ax1[None, :]
        a = (hr[sy, :] * ay0[:, None]).astype(np.float64)
        out = (a + hr[sy1, :].astype(np.float64) * ay1[:, None].astype(np.float64)).astype(f32)
    elif variant == "fma_hv":
        a = (src[:, sx] * ax0[None, :]).astype(np.float64)
        hr = (a + src[:, sx1].astype(np.float64) * ax1[None, :].astype(np.float64)).astype(f32)
        a = (hr[sy, :] * ay0[:, None]).astype(np.float64)
        out = (a + hr[sy1, :].astype(np.float64) * ay1[:, None].astype(np.float64)).astype(f32)
    return out.astype(f32)

rng = np.random.default_rng(1)
for (H, W) in [(600, 800), (384, 512), (307, 410), (61, 77)]:
    src = (rng.random((H, W)) * 255).astype(f32)
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        ref = cv2.resize(src, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
        dH, dW = ref.shape
        for variant in ("plain", "fma_v", "fma_hv"):
            mod = resize_model(src, dW, dH, 1.0 / 0.8, 1.0 / 0.8, variant)
            print("down", (H, W), "->", (dH, dW), "ipp", ipp, variant, "maxdiff", np.abs(ref - mod).max(), "n_neq", int((ref != mod).sum()))
    # upsample form (explicit dsize), like the flow up-sampling
    small = cv2.resize(src, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
    sh, sw = small.shape
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        ref = cv2.resize(small, (W, H), interpolation=cv2.INTER_LINEAR)
        for variant in ("plain", "fma_v", "fma_hv"):
            mod = resize_model(small, W, H, 1.0 / (W / sw), 1.0 / (H / sh), variant)
            print("up  ", (sh, sw), "->", (H, W), "ipp", ipp, variant, "maxdiff", np.abs(ref - mod).max(), "n_neq", int((ref != mod).sum()))

print("---- mismatch positions")
cv2.ipp.setUseIPP(False)
rng = np.random.default_rng(1)
src = (rng.random((307, 410)) * 255).astype(f32)
ref = cv2.resize(src, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
mod = resize_model(src, 328, 246, 1.25, 1.25, "plain")
bad = np.argwhere(ref != mod)
print(bad[:30].tolist())
sx, a0, a1 = lin_coeffs(410, 328, 1.25)
sy, b0, b1 = lin_coeffs(307, 246, 1.25)
for y, x in bad[:5]:
    print(y, x, "sy", sy[y], b0[y], b1[y], "sx", sx[x], a0[x], a1[x], ref[y, x], mod[y, x])
