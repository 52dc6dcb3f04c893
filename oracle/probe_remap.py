"""Probe: is a float32 numpy model of cv2.remap(INTER_CUBIC, BORDER_CONSTANT) bit-exact with the real cv2?"""
import numpy as np, cv2
f32 = np.float32

def cubic_tab():
    A = f32(-0.75)
    tab = np.zeros((32, 4), f32)
    scale = f32(1.0) / f32(32)
    for i in range(32):
        x = f32(i) * scale
        x1 = x + f32(1)
        c0 = ((A * x1 - f32(5) * A) * x1 + f32(8) * A) * x1 - f32(4) * A
        c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + f32(1)
        ox = f32(1) - x
        c2 = ((A + f32(2)) * ox - (A + f32(3))) * ox * ox + f32(1)
        c3 = f32(1) - c0 - c1 - c2
        tab[i] = [c0, c1, c2, c3]
    return tab

def remap_model(src, mapx, mapy, order="rows"):
    H, W = src.shape
    tab = cubic_tab()
    ix = np.rint(mapx * f32(32)).astype(np.int64)
    iy = np.rint(mapy * f32(32)).astype(np.int64)
    sx = (ix >> 5) - 1
    sy = (iy >> 5) - 1
    fx = ix & 31
    fy = iy & 31
    out = np.zeros(mapx.shape, f32)
    pad = np.zeros((H + 8, W + 8), f32)
    pad[4:4 + H, 4:4 + W] = src
    inside = (sx >= 0) & (sx <= W - 4) & (sy >= 0) & (sy <= H - 4)
    fully_out = (sx + 3 < 0) | (sx >= W) | (sy + 3 < 0) | (sy >= H)
    sxc = np.clip(sx, -4, W); syc = np.clip(sy, -4, H)
    rows = []
    seq = np.zeros(mapx.shape, f32)
    for k1 in range(4):
        r = None
        for k2 in range(4):
            w = tab[fy, k1] * tab[fx, k2]
            yy = syc + k1; xx = sxc + k2
            valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            v = pad[np.clip(yy + 4, 0, H + 7), np.clip(xx + 4, 0, W + 7)]
            term = v * w
            r = term if r is None else r + term
            seq = np.where(valid, seq + term, seq)
        rows.append(r)
    fast = ((rows[0] + rows[1]) + rows[2]) + rows[3]
    out = np.where(inside, fast, seq)
    out = np.where(fully_out, f32(0), out)
    return out

rng = np.random.default_rng(0)
H, W = 120, 160
src = (rng.random((H, W)) * 255).astype(f32)
yy, xx = np.mgrid[0:H, 0:W].astype(f32)
u1 = (rng.standard_normal((H, W)) * 3).astype(f32)
u2 = (rng.standard_normal((H, W)) * 3).astype(f32)
mapx = xx + u1; mapy = yy + u2
for ipp in (True, False):
    cv2.ipp.setUseIPP(ipp)
    for opt in (True, False):
        cv2.setUseOptimized(opt)
        ref = cv2.remap(src, mapx, mapy, cv2.INTER_CUBIC)
        mod = remap_model(src, mapx, mapy)
        d = np.abs(ref - mod)
        print("ipp", ipp, "opt", opt, "maxdiff", d.max(), "n_neq", int((ref != mod).sum()), "of", ref.size)

cv2.ipp.setUseIPP(True); cv2.setUseOptimized(True)
ref = cv2.remap(src, mapx, mapy, cv2.INTER_CUBIC)
mod = remap_model(src, mapx, mapy)
bad = np.argwhere(ref != mod)
ix = np.rint(mapx * f32(32)).astype(np.int64); iy = np.rint(mapy * f32(32)).astype(np.int64)
for (y, x) in bad[:25]:
    print(y, x, "sx", (ix[y,x]>>5)-1, "sy", (iy[y,x]>>5)-1, "fx", ix[y,x]&31, "fy", iy[y,x]&31, ref[y,x], mod[y,x])
