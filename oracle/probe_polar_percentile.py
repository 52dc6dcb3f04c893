"""Probe: float32 model of cv2.cartToPolar, and the exact arithmetic of np.percentile / np.histogram."""
import numpy as np, cv2
f32 = np.float32
rng = np.random.default_rng(0)
x = (rng.standard_normal((300, 400)) * 2).astype(f32); y = (rng.standard_normal((300, 400)) * 2).astype(f32)
x[:20] = 0; y[:10] = 0; y[40:50] = 0
mag, ang = cv2.cartToPolar(x, y)

def cart_to_polar_model(x, y):
    mag = np.sqrt(x * x + y * y)
    p1 = f32(0.9997878412794807) * f32(180 / np.pi); p3 = f32(-0.3258083974640975) * f32(180 / np.pi)
    p5 = f32(0.1555786518463281) * f32(180 / np.pi); p7 = f32(-0.04432655554792128) * f32(180 / np.pi)
    ax, ay = np.abs(x), np.abs(y)
    eps = f32(2.220446049250313e-16)
    with np.errstate(all="ignore"):
        c1 = ay / (ax + eps); c2 = ax / (ay + eps)
    def poly(c):
        cc = c * c
        return (((p7 * cc + p5) * cc + p3) * cc + p1) * c
    a = np.where(ax >= ay, poly(c1), f32(90) - poly(c2)).astype(f32)
    a = np.where(x < 0, f32(180) - a, a).astype(f32)
    a = np.where(y < 0, f32(360) - a, a).astype(f32)
    return mag.astype(f32), (a * f32(np.pi / 180)).astype(f32)

m2, a2 = cart_to_polar_model(x, y)
print("mag equal", np.array_equal(mag, m2), np.abs(mag - m2).max(), "ang equal", np.array_equal(ang, a2), np.abs(ang - a2).max(), (ang != a2).sum())
bad = np.argwhere(ang != a2)[:5]
for (i, j) in bad: print(x[i, j], y[i, j], ang[i, j], a2[i, j])

# percentile: float32 and float64 semantics
v = rng.standard_normal(10007).astype(f32)
for q in (1, 99):
    r = np.percentile(v, q)
    s = np.sort(v); n = len(s)
    virt = (n - 1) * (q / 100.0)
    lo = int(np.floor(virt)); g = virt - lo
    cand64 = float(s[lo]) + (float(s[min(lo + 1, n - 1)]) - float(s[lo])) * g
    g32 = f32(g)
    cand32 = s[lo] + (s[lo + 1] - s[lo]) * g32
    # numpy _lerp: a + (b-a)*t, with subtract(b, diff*(1-t)) where t>=0.5
    d = s[lo + 1] - s[lo]
    lerp = s[lo] + d * g32 if g < 0.5 else s[lo + 1] - d * (f32(1) - g32)
    print(q, type(r), r.dtype, repr(r), "cand64", repr(f32(cand64)), "cand32", repr(cand32), "lerp32", repr(lerp))
v64 = v.astype(np.float64) * 1.2345678
for q in (1, 99):
    r = np.percentile(v64, q)
    s = np.sort(v64); n = len(s); virt = (n - 1) * (q / 100.0); lo = int(np.floor(virt)); g = virt - lo
    d = s[lo + 1] - s[lo]
    lerp = s[lo] + d * g if g < 0.5 else s[lo + 1] - d * (1 - g)
    print(q, repr(r), "lerp64", repr(lerp), "plain", repr(s[lo] + d * g))
