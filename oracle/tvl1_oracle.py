"""ctypes front-end of the CPU oracle (oracle/tvl1_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (tee_optical_flow_b200) never imports this module.

The object mirrors the duck type the reference drives (optical_flow/calculate_optical_flow.py:577-578,642):
``createOptFlow_DualTVL1()`` -> ``.setLambda(x)`` -> ``.calc(I0, I1, None) -> (H, W, 2) float32``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libtvl1_oracle.so"
_lib = None


class _Params(C.Structure):
    _fields_ = [
        ("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("epsilon", C.c_double),
        ("scale_step", C.c_double), ("nscales", C.c_int), ("warps", C.c_int), ("inner_iterations", C.c_int),
        ("outer_iterations", C.c_int), ("median_filtering", C.c_int), ("err_mode", C.c_int),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [_HERE / "tvl1_oracle.c", _HERE / "Makefile"]
    if (not force and _LIB_PATH.exists()
            and all(_LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs)):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", str(_HERE), "-B", "libtvl1_oracle.so"], check=True, env=env,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        fp = C.POINTER(C.c_float)
        L.tvl1_oracle_calc.restype = C.c_int
        L.tvl1_oracle_calc.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       fp, C.POINTER(C.c_int)]
        L.oracle_resize_linear.restype = None
        L.oracle_resize_linear.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.oracle_remap_cubic.restype = None
        L.oracle_remap_cubic.argtypes = [fp, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_int]
        L.oracle_median_blur.restype = C.c_int
        L.oracle_median_blur.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int]
        L.oracle_centered_gradient.restype = None
        L.oracle_centered_gradient.argtypes = [fp, C.c_int, C.c_int, fp, fp]
        L.oracle_warp_step.restype = None
        L.oracle_warp_step.argtypes = [fp] * 6 + [C.c_int, C.c_int] + [fp] * 4
        L.oracle_inner_iteration.restype = C.c_double
        L.oracle_inner_iteration.argtypes = [fp] * 10 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
        L.oracle_cubic_table.restype = None
        L.oracle_cubic_table.argtypes = [fp]
        L.oracle_scaled_size.restype = None
        L.oracle_scaled_size.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_set_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def set_threads(n: int = 0) -> int:
    """OpenMP threads used by the oracle (0 = all host cores); returns the value in effect."""
    return lib().oracle_set_threads(int(n) if n > 0 else (os.cpu_count() or 1))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------------- primitives
def resize_linear(src, dsize_hw, scale_xy=None) -> np.ndarray:
    """cv::resize(INTER_LINEAR) on float32.  scale_xy=None -> explicit-dsize form (scale = 1/(dst/src))."""
    src = _f32c(src)
    sH, sW = src.shape
    dH, dW = dsize_hw
    if scale_xy is None:
        scale_xy = (1.0 / (dW / sW), 1.0 / (dH / sH))
    dst = np.empty((dH, dW), np.float32)
    lib().oracle_resize_linear(_fp(src), sH, sW, _fp(dst), dH, dW, float(scale_xy[0]), float(scale_xy[1]))
    return dst


def resize_by_factor(src, f: float) -> np.ndarray:
    """cv::resize(src, Size(), f, f, INTER_LINEAR): dsize = cvRound(size*f), scale = 1/f."""
    src = _f32c(src)
    dH, dW = C.c_int(), C.c_int()
    lib().oracle_scaled_size(src.shape[0], src.shape[1], float(f), C.byref(dH), C.byref(dW))
    return resize_linear(src, (dH.value, dW.value), (1.0 / f, 1.0 / f))


def remap_cubic(src, mapx, mapy) -> np.ndarray:
    src, mapx, mapy = _f32c(src), _f32c(mapx), _f32c(mapy)
    dst = np.empty(mapx.shape, np.float32)
    lib().oracle_remap_cubic(_fp(src), src.shape[0], src.shape[1], _fp(mapx), _fp(mapy), _fp(dst),
                             mapx.shape[0], mapx.shape[1])
    return dst


def median_blur(src, ksize: int = 5) -> np.ndarray:
    src = _f32c(src)
    dst = np.empty_like(src)
    rc = lib().oracle_median_blur(_fp(src), _fp(dst), src.shape[0], src.shape[1], int(ksize))
    if rc != 0:
        raise ValueError(f"unsupported median ksize {ksize}")
    return dst


def centered_gradient(src):
    src = _f32c(src)
    dx, dy = np.empty_like(src), np.empty_like(src)
    lib().oracle_centered_gradient(_fp(src), src.shape[0], src.shape[1], _fp(dx), _fp(dy))
    return dx, dy


def warp_step(I0, I1, u1, u2):
    """centeredGradient(I1) + buildFlowMap + 3x remap + calcGradRho -> (I1wx, I1wy, grad, rho_c)."""
    I0, I1, u1, u2 = map(_f32c, (I0, I1, u1, u2))
    I1x, I1y = centered_gradient(I1)
    H, W = I0.shape
    outs = [np.empty((H, W), np.float32) for _ in range(4)]
    lib().oracle_warp_step(_fp(I0), _fp(I1), _fp(I1x), _fp(I1y), _fp(u1), _fp(u2), H, W, *[_fp(o) for o in outs])
    return tuple(outs)


def inner_iteration(I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, l_t, theta, taut, err_mode=1):
    """One primal-dual iteration; u*/p* are updated IN PLACE (must be C-contiguous float32). Returns the error."""
    arrs = [I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22]
    for a in arrs:
        assert a.dtype == np.float32 and a.flags.c_contiguous
    H, W = u1.shape
    return lib().oracle_inner_iteration(*[_fp(a) for a in arrs], H, W, np.float32(l_t), np.float32(theta),
                                        np.float32(taut), int(err_mode))


def cubic_table() -> np.ndarray:
    t = np.empty((32, 4), np.float32)
    lib().oracle_cubic_table(_fp(t))
    return t


# ---------------------------------------------------------------------------------------------- the solver
class OracleDualTVL1:
    """Drop-in for cv2.optflow.DualTVL1OpticalFlow (CPU): same setters/getters, same calc signature."""

    def __init__(self, tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01, innnerIterations=30,
                 outerIterations=10, scaleStep=0.8, gamma=0.0, medianFiltering=5, useInitialFlow=False,
                 err_mode=0):
        if gamma != 0.0 or useInitialFlow:
            raise NotImplementedError("gamma != 0 / useInitialFlow are never set by the reference")
        self.p = _Params(tau, lambda_, theta, epsilon, scaleStep, nscales, warps, innnerIterations,
                         outerIterations, medianFiltering, err_mode)
        self.last_counters = None
        self.last_nscales = None

    # OpenCV-named accessors used by the reference (calculate_optical_flow.py:578) and by the parity tests
    def setLambda(self, v): self.p.lambda_ = float(v)
    def getLambda(self): return self.p.lambda_
    def setTau(self, v): self.p.tau = float(v)
    def setTheta(self, v): self.p.theta = float(v)
    def setEpsilon(self, v): self.p.epsilon = float(v)
    def setScaleStep(self, v): self.p.scale_step = float(v)
    def setScalesNumber(self, v): self.p.nscales = int(v)
    def setWarpingsNumber(self, v): self.p.warps = int(v)
    def setInnerIterations(self, v): self.p.inner_iterations = int(v)
    def setOuterIterations(self, v): self.p.outer_iterations = int(v)
    def setMedianFiltering(self, v): self.p.median_filtering = int(v)

    def calc(self, I0, I1, flow=None) -> np.ndarray:
        I0 = np.ascontiguousarray(I0)
        I1 = np.ascontiguousarray(I1)
        if I0.shape != I1.shape or I0.ndim != 2 or I0.dtype != I1.dtype:
            raise ValueError("I0/I1 must be 2-D arrays of identical shape and dtype")
        if I0.dtype == np.uint8:
            is_f32 = 0
        elif I0.dtype == np.float32:
            is_f32 = 1
        else:
            raise ValueError("DualTVL1 accepts CV_8UC1 or CV_32FC1")
        H, W = I0.shape
        out = np.empty((H, W, 2), np.float32)
        counters = np.zeros((self.p.nscales, 3), np.int32)
        rc = lib().tvl1_oracle_calc(C.byref(self.p), I0.ctypes.data, I1.ctypes.data, is_f32, H, W, _fp(out),
                                    counters.ctypes.data_as(C.POINTER(C.c_int)))
        if rc < 0:
            raise RuntimeError(f"tvl1_oracle_calc failed: {rc}")
        self.last_counters = counters
        self.last_nscales = rc
        return out


def create_reference_model(**kw):
    """The real cv2.optflow when this interpreter has opencv-contrib, else the restated oracle."""
    try:
        import cv2
        if hasattr(cv2, "optflow"):
            return cv2.optflow.createOptFlow_DualTVL1(), "cv2.optflow"
    except Exception:
        pass
    return OracleDualTVL1(**kw), "restated-oracle"
