"""Second, independent restatement of OpenCV's CPU DualTVL1: numpy float32 pointwise steps composed with the
GENUINE cv2.resize / cv2.remap / cv2.medianBlur of the OpenCV in this image.

TEST INFRASTRUCTURE ONLY (never imported by the product).  Shares no code with oracle/tvl1_oracle.c; it exists
to (a) cross-check the C oracle and (b) generate the golden fixtures under tests/golden/
(tests/golden/make_golden.py).  Reference call sites: optical_flow/calculate_optical_flow.py:577-578,642.

``ipp=False`` switches OpenCV's IPP dispatch off so that cv2.resize runs OpenCV's own (open-source) linear
resize, which the C oracle reproduces bit for bit; with ``ipp=True`` Intel IPP's resize is used and differs by
<= 3e-5 grey levels on 0..255 images (oracle/probe_resize.py).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
FLT_EPSILON = np.finfo(np.float32).eps
FLT_MAX = np.finfo(np.float32).max


def centered_gradient(I):
    Ip = np.pad(I, 1, mode="edge")
    dx = f32(0.5) * (Ip[1:-1, 2:] - Ip[1:-1, :-2])
    dy = f32(0.5) * (Ip[2:, 1:-1] - Ip[:-2, 1:-1])
    return dx.astype(f32), dy.astype(f32)


def forward_gradient(u):
    ux = np.zeros_like(u)
    uy = np.zeros_like(u)
    ux[:, :-1] = u[:, 1:] - u[:, :-1]
    uy[:-1, :] = u[1:, :] - u[:-1, :]
    return ux, uy


def divergence(v1, v2):
    div = np.empty_like(v1)
    div[1:, 1:] = (v1[1:, 1:] - v1[1:, :-1]) + (v2[1:, 1:] - v2[:-1, 1:])
    div[0, 1:] = (v1[0, 1:] - v1[0, :-1]) + v2[0, 1:]
    div[1:, 0] = (v1[1:, 0] + v2[1:, 0]) - v2[:-1, 0]
    div[0, 0] = v1[0, 0] + v2[0, 0]
    return div


def inner_iteration(I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, l_t, theta, taut, err_mode):
    l_t, theta, taut = f32(l_t), f32(theta), f32(taut)
    rho = rho_c + (I1wx * u1 + I1wy * u2)
    lg = l_t * grad
    c1 = rho < -lg
    c2 = (~c1) & (rho > lg)
    c3 = (~c1) & (~c2) & (grad > FLT_EPSILON)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = np.where(c3, -rho / grad, f32(0)).astype(f32)
    d1 = np.where(c1, l_t * I1wx, np.where(c2, -l_t * I1wx, np.where(c3, fi * I1wx, f32(0)))).astype(f32)
    d2 = np.where(c1, l_t * I1wy, np.where(c2, -l_t * I1wy, np.where(c3, fi * I1wy, f32(0)))).astype(f32)
    v1 = u1 + d1
    v2 = u2 + d2
    div1 = divergence(p11, p12)
    div2 = divergence(p21, p22)
    u1n = v1 + theta * div1
    u2n = v2 + theta * div2
    term = (u1n - u1) * (u1n - u1) + (u2n - u2) * (u2n - u2)
    if err_mode == 0:
        err = float(np.cumsum(term.ravel(), dtype=f32)[-1])   # serial float32 accumulation, raster order
    else:
        err = float(np.sum(term.ravel().astype(np.float64)))
    u1x, u1y = forward_gradient(u1n)
    u2x, u2y = forward_gradient(u2n)
    g1 = np.sqrt(u1x.astype(np.float64) ** 2 + u1y.astype(np.float64) ** 2).astype(f32)
    g2 = np.sqrt(u2x.astype(np.float64) ** 2 + u2y.astype(np.float64) ** 2).astype(f32)
    ng1 = f32(1) + taut * g1
    ng2 = f32(1) + taut * g2
    p11n = (p11 + taut * u1x) / ng1
    p12n = (p12 + taut * u1y) / ng1
    p21n = (p21 + taut * u2x) / ng2
    p22n = (p22 + taut * u2y) / ng2
    return u1n, u2n, p11n, p12n, p21n, p22n, err


class Cv2ComposedDualTVL1:
    def __init__(self, tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01, inner=30, outer=10,
                 scale_step=0.8, median=5, err_mode=0, ipp=False):
        self.tau, self.lambda_, self.theta = tau, lambda_, theta
        self.nscales, self.warps, self.epsilon = nscales, warps, epsilon
        self.inner, self.outer, self.scale_step, self.median = inner, outer, scale_step, median
        self.err_mode, self.ipp = err_mode, ipp
        self.last_counters = None

    def setLambda(self, v):
        self.lambda_ = float(v)

    def _proc_one_scale(self, cv2, I0, I1, u1, u2, counters):
        H, W = I0.shape
        scaled_eps = f32(self.epsilon * self.epsilon * (H * W))
        I1x, I1y = centered_gradient(I1)
        p11 = np.zeros((H, W), f32); p12 = np.zeros((H, W), f32)
        p21 = np.zeros((H, W), f32); p22 = np.zeros((H, W), f32)
        l_t = f32(self.lambda_ * self.theta)
        taut = f32(self.tau / self.theta)
        theta = f32(self.theta)
        yy, xx = np.mgrid[0:H, 0:W]
        xx = xx.astype(f32); yy = yy.astype(f32)
        for _ in range(self.warps):
            m1 = xx + u1
            m2 = yy + u2
            I1w = cv2.remap(I1, m1, m2, cv2.INTER_CUBIC)
            I1wx = cv2.remap(I1x, m1, m2, cv2.INTER_CUBIC)
            I1wy = cv2.remap(I1y, m1, m2, cv2.INTER_CUBIC)
            grad = I1wx * I1wx + I1wy * I1wy
            rho_c = ((I1w - I1wx * u1) - I1wy * u2) - I0
            counters[2] += 1
            error = FLT_MAX
            n_outer = 0
            while error > scaled_eps and n_outer < self.outer:
                if self.median > 1:
                    u1 = cv2.medianBlur(u1, self.median)
                    u2 = cv2.medianBlur(u2, self.median)
                    counters[1] += 1
                n_inner = 0
                while error > scaled_eps and n_inner < self.inner:
                    u1, u2, p11, p12, p21, p22, err = inner_iteration(
                        I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, l_t, theta, taut, self.err_mode)
                    error = f32(err)
                    counters[0] += 1
                    n_inner += 1
                n_outer += 1
        return u1, u2

    def calc(self, I0, I1, flow=None):
        import cv2
        old_ipp = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(bool(self.ipp))
        try:
            mult = f32(1.0) if I0.dtype == np.uint8 else f32(255.0)
            I0s = [I0.astype(f32) * mult]
            I1s = [I1.astype(f32) * mult]
            nscales = self.nscales
            for s in range(1, nscales):
                a = cv2.resize(I0s[s - 1], None, fx=self.scale_step, fy=self.scale_step,
                               interpolation=cv2.INTER_LINEAR)
                b = cv2.resize(I1s[s - 1], None, fx=self.scale_step, fy=self.scale_step,
                               interpolation=cv2.INTER_LINEAR)
                I0s.append(a); I1s.append(b)
                if a.shape[0] < 16 or a.shape[1] < 16:
                    nscales = s
                    break
            counters = np.zeros((self.nscales, 3), np.int32)
            u1 = np.zeros(I0s[nscales - 1].shape, f32)
            u2 = np.zeros(I0s[nscales - 1].shape, f32)
            for s in range(nscales - 1, -1, -1):
                u1, u2 = self._proc_one_scale(cv2, I0s[s], I1s[s], u1, u2, counters[s])
                if s == 0:
                    break
                Hn, Wn = I0s[s - 1].shape
                u1 = cv2.resize(u1, (Wn, Hn), interpolation=cv2.INTER_LINEAR) * f32(1.0 / self.scale_step)
                u2 = cv2.resize(u2, (Wn, Hn), interpolation=cv2.INTER_LINEAR) * f32(1.0 / self.scale_step)
            self.last_counters = counters
            return np.stack([u1, u2], axis=-1).astype(f32)
        finally:
            cv2.ipp.setUseIPP(old_ipp)
