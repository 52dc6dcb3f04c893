"""TVL1Engine -- the B200 TV-L1 solver behind the reference's operator interface.

The reference selects its optical-flow operator by the string ``OF_algo`` and then only ever calls
``OF_model.setLambda(x)`` and ``OF_model.calc(I0, I1, None)`` on it
(optical_flow/calculate_optical_flow.py:564-578, 627-646).  ``TVL1Engine`` has that duck type (plus the other
OpenCV-named setters/getters of ``cv2.optflow.DualTVL1OpticalFlow``), so it can be handed to the reference's
``calculate_optical_flow(..., OF_model=engine)`` unchanged, and adds the batched ``calc_clip`` that solves all
frame pairs of a clip concurrently on the GPU.

All arithmetic happens in libteeflow.so (hand-written sm_100a CUDA, include/teeflow.h).  PyTorch is used only
for device buffers and streams.  No CPU fallback: without the library or a GPU this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .exceptions import EngineUnavailableError, OpticalFlowCalculationError

_PARAM_KEYS = {
    "tau": "tau", "lambda_": "lambda", "theta": "theta", "epsilon": "epsilon", "scale_step": "scale_step",
    "nscales": "nscales", "warps": "warps", "inner_iterations": "inner_iterations",
    "outer_iterations": "outer_iterations", "median_filtering": "median_filtering", "max_slots": "max_slots",
}


def _dtype_code(dt) -> int:
    if dt == np.uint8:
        return _lib.TEEFLOW_U8
    if dt == np.float32:
        return _lib.TEEFLOW_F32
    # same restriction as OpenCV: CV_8UC1 or CV_32FC1 (cv2.error in the reference)
    raise OpticalFlowCalculationError(f"TV-L1 input must be uint8 or float32, got {dt}")


class TVL1Engine:
    """Drop-in for ``cv2.optflow.createOptFlow_DualTVL1()`` (calculate_optical_flow.py:577)."""

    def __init__(self, device: Optional[int] = None, **params):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        p = _lib.TeeflowParams()
        self._lib.teeflow_default_params(C.byref(p))
        for k, v in params.items():
            if k not in _PARAM_KEYS:
                raise TypeError(f"unknown TV-L1 parameter {k!r}")
            setattr(p, k, v)
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except ImportError:  # pragma: no cover
                pass
        self.device = int(device)
        rc = self._lib.teeflow_create(C.byref(p), self.device, C.byref(self._h))
        if rc != 0:
            msg = (self._lib.teeflow_last_error(None) or b"").decode()
            self._h = C.c_void_p()
            if rc == _lib.ERR_CUDA:
                raise EngineUnavailableError(msg)
            raise OpticalFlowCalculationError(f"teeflow_create failed ({rc}): {msg}")

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.teeflow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _check(self, rc: int):
        if rc < 0:
            msg = (self._lib.teeflow_last_error(self._h) or b"").decode()
            raise OpticalFlowCalculationError(f"libteeflow error {rc}: {msg}")
        return rc

    # ------------------------------------------------------------------ parameters (OpenCV names)
    def _set(self, key, v):
        self._check(self._lib.teeflow_set_param(self._h, key.encode(), float(v)))

    def _get(self, key):
        out = C.c_double()
        self._check(self._lib.teeflow_get_param(self._h, key.encode(), C.byref(out)))
        return out.value

    def setLambda(self, v): self._set("lambda", v)
    def getLambda(self): return self._get("lambda")
    def setTau(self, v): self._set("tau", v)
    def getTau(self): return self._get("tau")
    def setTheta(self, v): self._set("theta", v)
    def getTheta(self): return self._get("theta")
    def setEpsilon(self, v): self._set("epsilon", v)
    def getEpsilon(self): return self._get("epsilon")
    def setScaleStep(self, v): self._set("scale_step", v)
    def getScaleStep(self): return self._get("scale_step")
    def setScalesNumber(self, v): self._set("nscales", v)
    def getScalesNumber(self): return int(self._get("nscales"))
    def setWarpingsNumber(self, v): self._set("warps", v)
    def getWarpingsNumber(self): return int(self._get("warps"))
    def setInnerIterations(self, v): self._set("inner_iterations", v)
    def getInnerIterations(self): return int(self._get("inner_iterations"))
    def setOuterIterations(self, v): self._set("outer_iterations", v)
    def getOuterIterations(self): return int(self._get("outer_iterations"))
    def setMedianFiltering(self, v): self._set("median_filtering", v)
    def getMedianFiltering(self): return int(self._get("median_filtering"))

    def setGamma(self, v):
        if float(v) != 0.0:
            raise OpticalFlowCalculationError("gamma != 0 is not supported (the reference never sets it)")

    def getGamma(self): return 0.0

    def setUseInitialFlow(self, v):
        if v:
            raise OpticalFlowCalculationError("useInitialFlow is not supported (the reference never sets it)")

    def getUseInitialFlow(self): return False

    # ------------------------------------------------------------------ OF_model.calc (one pair, host arrays)
    def calc(self, I0, I1, flow=None) -> np.ndarray:
        """``OF_model.calc(saliency_1, saliency_2, None)`` (calculate_optical_flow.py:642): two (H, W) uint8 or
        float32 images -> (H, W, 2) float32, channel 0 = x displacement, 1 = y."""
        I0 = np.ascontiguousarray(I0)
        I1 = np.ascontiguousarray(I1)
        if I0.ndim != 2 or I0.shape != I1.shape or I0.dtype != I1.dtype:
            raise OpticalFlowCalculationError("I0 and I1 must be 2-D arrays of identical shape and dtype")
        code = _dtype_code(I0.dtype)
        H, W = I0.shape
        out = np.empty((H, W, 2), np.float32)
        self._batch_counters = None
        self._check(self._lib.teeflow_calc_pair_host(self._h, I0.ctypes.data, I1.ctypes.data, code, H, W,
                                                     out.ctypes.data))
        return out

    # ------------------------------------------------------------------ all pairs of a clip
    def calc_clip(self, frames, out_scale: float = 1.0, duplicate_last: bool = True, want_f32: bool = True,
                  want_f16: bool = False, asynchronous: bool = False):
        """Flow of every consecutive pair of ``frames`` (N, H, W) -- the reference's pair loop
        (calculate_optical_flow.py:584-600) in one call.

        numpy frames  -> host path (H2D and D2H inside the call), numpy results.
        torch CUDA tensor -> device path, torch results on the same device.
        Returns (flow_f32 or None, flow_f16 or None), each (N_out, H, W, 2) with N_out = N-1 (+1 when
        ``duplicate_last``: the reference appends a copy of the last flow, :599).
        ``asynchronous`` (device path): return as soon as the work is enqueued on the current CUDA stream; the result
        tensors are complete for later work on that stream, `finish()` (or last_counters()) fetches the verdict.
        """
        if isinstance(frames, np.ndarray):
            return self._calc_clip_host(frames, out_scale, duplicate_last, want_f32, want_f16)
        return self._calc_clip_device(frames, out_scale, duplicate_last, want_f32, want_f16, asynchronous=asynchronous)

    def _calc_clip_host(self, frames, out_scale, duplicate_last, want_f32, want_f16):
        frames = np.ascontiguousarray(frames)
        if frames.ndim != 3:
            raise OpticalFlowCalculationError("frames must be (N, H, W)")
        code = _dtype_code(frames.dtype)
        N, H, W = frames.shape
        n_out = N - 1 + (1 if duplicate_last else 0)
        f32 = np.empty((n_out, H, W, 2), np.float32) if want_f32 else None
        f16 = np.empty((n_out, H, W, 2), np.float16) if want_f16 else None
        self._batch_counters = None
        self._check(self._lib.teeflow_calc_clip_host(
            self._h, frames.ctypes.data, code, N, H, W, f32.ctypes.data if want_f32 else None,
            f16.ctypes.data if want_f16 else None, float(np.float32(out_scale)), int(duplicate_last)))
        return f32, f16

    def finish(self) -> None:
        """Completes an asynchronous device run (`calc_clip(..., asynchronous=True)`): waits for it on its stream,
        raises if the scheduler reported a problem.  last_counters() calls it implicitly."""
        self._check(self._lib.teeflow_finish(self._h))

    def _calc_clip_device(self, frames, out_scale, duplicate_last, want_f32, want_f16, out_f32=None, out_f16=None,
                          asynchronous: bool = False):
        import torch
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda):
            raise OpticalFlowCalculationError("frames must be a numpy array or a CUDA torch tensor")
        if frames.device.index != self.device:
            raise OpticalFlowCalculationError(f"frames live on cuda:{frames.device.index}, engine on cuda:{self.device}")
        if frames.dim() != 3:
            raise OpticalFlowCalculationError("frames must be (N, H, W)")
        if frames.dtype == torch.uint8:
            code = _lib.TEEFLOW_U8
        elif frames.dtype == torch.float32:
            code = _lib.TEEFLOW_F32
        else:
            raise OpticalFlowCalculationError(f"TV-L1 input must be uint8 or float32, got {frames.dtype}")
        frames = frames.contiguous()
        N, H, W = frames.shape
        n_out = N - 1 + (1 if duplicate_last else 0)
        f32 = out_f32 if out_f32 is not None else (
            torch.empty((n_out, H, W, 2), dtype=torch.float32, device=frames.device) if want_f32 else None)
        f16 = out_f16 if out_f16 is not None else (
            torch.empty((n_out, H, W, 2), dtype=torch.float16, device=frames.device) if want_f16 else None)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        self._batch_counters = None
        entry = self._lib.teeflow_calc_clip_async if asynchronous else self._lib.teeflow_calc_clip
        self._keep = frames                        # an asynchronous run reads the frames after this call returns
        self._check(entry(
            self._h, frames.data_ptr(), code, N, H, W, H * W, f32.data_ptr() if f32 is not None else None,
            f16.data_ptr() if f16 is not None else None, float(np.float32(out_scale)), int(duplicate_last),
            C.c_void_p(stream)))
        return f32, f16

    def calc_pairs_device(self, frames, pair_a, pair_b, out_index=None, dup_index=None, n_out=None,
                          out_scale: float = 1.0, want_f32: bool = True, want_f16: bool = False, out_f32=None, out_f16=None):
        """Generic form: arbitrary (frame a -> frame b) pairs over a CUDA frame tensor (N, H, W); used for
        batches of clips and for sharding a clip by pair range.  out_f32 / out_f16: preallocated
        (n_out, H, W, 2) CUDA tensors to write into instead of allocating."""
        import torch
        pair_a = np.ascontiguousarray(pair_a, np.int32)
        pair_b = np.ascontiguousarray(pair_b, np.int32)
        n_pairs = len(pair_a)
        out_index = np.arange(n_pairs, dtype=np.int32) if out_index is None else np.ascontiguousarray(out_index, np.int32)
        dup_index = np.full(n_pairs, -1, np.int32) if dup_index is None else np.ascontiguousarray(dup_index, np.int32)
        if n_out is None:
            n_out = int(max(out_index.max(initial=-1), dup_index.max(initial=-1)) + 1)
        frames = frames.contiguous()
        N, H, W = frames.shape
        code = _lib.TEEFLOW_U8 if frames.dtype == torch.uint8 else _lib.TEEFLOW_F32
        for t, dt in ((out_f32, torch.float32), (out_f16, torch.float16)):
            if t is not None and (t.dtype != dt or tuple(t.shape) != (n_out, H, W, 2) or not t.is_contiguous() or t.device != frames.device):
                raise OpticalFlowCalculationError("preallocated output must be a contiguous (n_out, H, W, 2) tensor of the right dtype")
        f32 = out_f32 if out_f32 is not None else \
            (torch.empty((n_out, H, W, 2), dtype=torch.float32, device=frames.device) if want_f32 else None)
        f16 = out_f16 if out_f16 is not None else \
            (torch.empty((n_out, H, W, 2), dtype=torch.float16, device=frames.device) if want_f16 else None)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        self._batch_counters = None
        self._check(self._lib.teeflow_calc_pairs(
            self._h, frames.data_ptr(), code, N, H, W, H * W, ip(pair_a), ip(pair_b), ip(out_index), ip(dup_index),
            n_pairs, f32.data_ptr() if f32 is not None else None, f16.data_ptr() if f16 is not None else None,
            float(np.float32(out_scale)), C.c_void_p(stream)))
        return f32, f16

    def calc_batch(self, clips, out_scale: float = 1.0, duplicate_last: bool = True, want_f32: bool = False,
                   want_f16: bool = True, max_frames_per_run: int = 1024):
        """A batch of equally shaped clips (B, N, H, W) on one GPU (BASELINE config 4, one rank's share): the pairs of
        up to max_frames_per_run frames' worth of clips go through ONE scheduler run, so slots freed by one clip are
        refilled with pairs of the next (no drain between clips); larger batches are cut into such runs, which bounds
        the pyramid workspace (~24 MB per 600x800 frame).  Returns (f32, f16) shaped (B, N_out, H, W, 2); the
        iteration counters of all runs are concatenated in last_counters()."""
        import torch
        if not (isinstance(clips, torch.Tensor) and clips.is_cuda and clips.dim() == 4):
            raise OpticalFlowCalculationError("clips must be a CUDA tensor (B, N, H, W)")
        B, N, H, W = clips.shape
        if N < 2:
            raise OpticalFlowCalculationError("a clip needs at least 2 frames")
        n_out = N if duplicate_last else N - 1
        per_run = max(1, int(max_frames_per_run) // N)
        f32 = torch.empty((B, n_out, H, W, 2), dtype=torch.float32, device=clips.device) if want_f32 else None
        f16 = torch.empty((B, n_out, H, W, 2), dtype=torch.float16, device=clips.device) if want_f16 else None
        counters, infos = [], []
        for b0 in range(0, B, per_run):
            b1 = min(b0 + per_run, B)
            nb = b1 - b0
            a = (np.arange(nb)[:, None] * N + np.arange(N - 1)[None, :]).ravel().astype(np.int32)
            o = (np.arange(nb)[:, None] * n_out + np.arange(N - 1)[None, :]).ravel().astype(np.int32)
            d = np.full(nb * (N - 1), -1, np.int32)
            if duplicate_last:
                d[N - 2::N - 1] = o[N - 2::N - 1] + 1
            self.calc_pairs_device(clips[b0:b1].reshape(nb * N, H, W), a, a + 1, o, d, n_out=nb * n_out, out_scale=out_scale,
                                   want_f32=want_f32, want_f16=want_f16,
                                   out_f32=None if f32 is None else f32[b0:b1].view(nb * n_out, H, W, 2),
                                   out_f16=None if f16 is None else f16[b0:b1].view(nb * n_out, H, W, 2))
            if B > per_run:
                c, info = self.last_counters()
                counters.append(c); infos.append(info)
        if counters:
            self._batch_counters = (np.concatenate(counters, axis=0), infos)
        else:
            self._batch_counters = None
        return f32, f16

    # ------------------------------------------------------------------ frame prep
    def prepare_frames(self, rgb):
        """img2uint8(rgb2gray(frame)) per frame (calculate_optical_flow.py:588) on the GPU.  rgb: (N,H,W,3) uint8,
        numpy -> numpy (N,H,W) uint8, CUDA torch tensor -> CUDA torch tensor."""
        import torch
        is_np = isinstance(rgb, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(rgb)) if is_np else rgb
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
            raise OpticalFlowCalculationError("rgb frames must be (N, H, W, 3) uint8")
        t = t.to(torch.device("cuda", self.device)).contiguous()
        N, H, W, _ = t.shape
        out = torch.empty((N, H, W), dtype=torch.uint8, device=t.device)
        stream = torch.cuda.current_stream(t.device).cuda_stream
        self._check(self._lib.teeflow_prepare_frames(self._h, t.data_ptr(), N, H, W, out.data_ptr(), C.c_void_p(stream)))
        return out.cpu().numpy() if is_np else out

    def compute_saliency(self, rgb, return_u8: bool = False):
        """cv2.saliency.StaticSaliencyFineGrained.computeSaliency per frame (calculate_optical_flow.py:560, :586)
        on the GPU: (N,H,W,3) uint8 -> (N,H,W) float32 in [0,1] (or the uint8 map with return_u8).  numpy -> numpy,
        CUDA torch tensor -> CUDA torch tensor.  Parity status: csrc/saliency_kernels.cuh."""
        import torch
        is_np = isinstance(rgb, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(rgb)) if is_np else rgb
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
            raise OpticalFlowCalculationError("rgb frames must be (N, H, W, 3) uint8")
        t = t.to(torch.device("cuda", self.device)).contiguous()
        N, H, W, _ = t.shape
        out = torch.empty((N, H, W), dtype=torch.uint8 if return_u8 else torch.float32, device=t.device)
        stream = torch.cuda.current_stream(t.device).cuda_stream
        self._check(self._lib.teeflow_saliency_fine_grained(
            self._h, t.data_ptr(), N, H, W, None if return_u8 else out.data_ptr(), out.data_ptr() if return_u8 else None,
            C.c_void_p(stream)))
        torch.cuda.current_stream(t.device).synchronize()
        return out.cpu().numpy() if is_np else out

    # ------------------------------------------------------------------ WASE background compensation
    def set_wase_masks(self, bkgd_mask, cache: bool = False) -> None:
        """bkgd_comp='WASE' (calculate_optical_flow.py:649-652): `bkgd_mask` is mask_dict['bkgd'], (N, H, W, 2) bool
        for ALL frames (numpy or CUDA torch tensor).  Builds the weight map w = sum_n bkgd[n] on the GPU and makes
        every following calc subtract the per-pair scalar mean(flow*bkgd != 0).  None switches it off.
        cache=True keeps the last weight map keyed by the mask OBJECT (the per-pair wrapper is called with the same
        mask_dict for every pair of a clip); the caller must not mutate that array in between."""
        import torch
        if bkgd_mask is None:
            if not cache:
                self._wase_w = None
                self._wase_key = None
            self._check(self._lib.teeflow_set_wase(self._h, None, 0, 0))
            return
        key = (id(bkgd_mask), tuple(bkgd_mask.shape))
        if cache and getattr(self, "_wase_key", None) == key and getattr(self, "_wase_w", None) is not None:
            w = self._wase_w
            self._check(self._lib.teeflow_set_wase(self._h, w.data_ptr(), w.shape[0], w.shape[1]))
            return
        m = bkgd_mask if isinstance(bkgd_mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bkgd_mask))
        if m.dim() != 4 or m.shape[-1] != 2:
            raise OpticalFlowCalculationError("bkgd mask must be (N, H, W, 2)")
        m = m.to(device=f"cuda:{self.device}").contiguous().view(torch.uint8) if m.dtype == torch.bool else \
            m.to(device=f"cuda:{self.device}", dtype=torch.uint8).contiguous()
        N, H, W, _ = m.shape
        w = torch.empty((H, W, 2), dtype=torch.float32, device=m.device)
        stream = torch.cuda.current_stream(m.device).cuda_stream
        self._check(self._lib.teeflow_wase_weights(self._h, m.data_ptr(), N, H, W, w.data_ptr(), C.c_void_p(stream)))
        torch.cuda.current_stream(m.device).synchronize()
        self._wase_w = w                      # keep the caller-owned buffer alive
        self._wase_key = key if cache else None
        self._wase_ref = bkgd_mask if cache else None   # pins id(): the key cannot be reused by another object
        self._check(self._lib.teeflow_set_wase(self._h, w.data_ptr(), H, W))

    def last_backgrounds(self) -> np.ndarray:
        st = _lib.TeeflowStats()
        self._check(self._lib.teeflow_get_stats(self._h, C.byref(st)))
        out = np.zeros(max(st.n_pairs, 1), np.float32)
        self._check(self._lib.teeflow_get_backgrounds(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), len(out)))
        return out[:st.n_pairs]

    # ------------------------------------------------------------------ decomposition + per-frame reductions
    def analyze_clip(self, flow_f16, mask, centroids, nframes: int, perc_lo: float = 1, perc_hi: float = 99) -> dict:
        """Per-frame waveforms of one label (analysis.py:215-327, cardiac_cycle_detection.py:100-116) on the GPU.
        flow_f16: (N,H,W,2) float16 stored flow; mask: (N,H,W,2) bool; centroids: (nframes,2) float64 (row, col)."""
        import torch
        dev = torch.device("cuda", self.device)
        f = flow_f16 if isinstance(flow_f16, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(flow_f16))
        m = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mask))
        if f.dtype != torch.float16 or f.dim() != 4 or f.shape[-1] != 2 or tuple(m.shape) != tuple(f.shape):
            raise OpticalFlowCalculationError("flow must be (N,H,W,2) float16 and mask (N,H,W,2) bool")
        f = f.to(dev).contiguous()
        m = (m.to(dev).contiguous().view(torch.uint8) if m.dtype == torch.bool else m.to(dev, torch.uint8).contiguous())
        N, H, W, _ = f.shape
        if not (1 <= nframes <= N):
            raise OpticalFlowCalculationError("nframes out of range")
        cent = np.ascontiguousarray(centroids, np.float64)
        if cent.shape != (nframes, 2):
            raise OpticalFlowCalculationError("centroids must be (nframes, 2) float64 (row, col)")
        res = dict(mag_hi=np.empty(nframes, np.float32), ang_mode=np.empty(nframes, np.float32),
                   rad_hi=np.empty(nframes), rad_lo=np.empty(nframes), long_hi=np.empty(nframes),
                   long_lo=np.empty(nframes), counts=np.empty((nframes, 4), np.int64))
        out = _lib.TeeflowAnalysis()
        for k, a in res.items():
            ct = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double, np.dtype(np.int64): C.c_int64}[a.dtype]
            setattr(out, k, a.ctypes.data_as(C.POINTER(ct)))
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._lib.teeflow_analyze_clip(self._h, f.data_ptr(), m.data_ptr(),
                                                   cent.ctypes.data_as(C.POINTER(C.c_double)), nframes, H, W,
                                                   float(perc_lo), float(perc_hi), C.byref(out), C.c_void_p(stream)))
        for k in ("mag_min", "mag_max", "ang_min", "ang_max"):
            res[k] = np.float32(getattr(out, k))
        for k in ("rad_min", "rad_max", "long_min", "long_max"):
            res[k] = np.float64(getattr(out, k))
        res["nframes"] = nframes
        return res

    def analysis_histogram(self, quantity: str, nframes: int, first, last, nbins: int = 1000):
        """np.histogram(non-zero, bins=nbins, range=(first, last)) per frame for 'mag' | 'ang' | 'rad' | 'long' of the
        last analyze_clip.  Returns (freq[nframes, nbins] int64 WITHOUT the reference's +1, edges)."""
        import torch
        qi = {"mag": 0, "ang": 1, "rad": 2, "long": 3}[quantity]
        dt = np.float32 if qi < 2 else np.float64
        first, last = dt(first), dt(last)
        if first == last:                       # numpy: first_edge -= 0.5; last_edge += 0.5
            first, last = dt(first - dt(0.5)), dt(last + dt(0.5))
        edges = np.linspace(first, last, nbins + 1, endpoint=True, dtype=dt)
        freq = np.empty((nframes, nbins), np.int64)
        stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        self._check(self._lib.teeflow_analysis_histogram(self._h, qi, edges.ctypes.data, nbins,
                                                         freq.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(stream)))
        return freq, edges

    # ------------------------------------------------------------------ accounting
    def last_counters(self) -> Tuple[np.ndarray, dict]:
        """(counters[n_pairs, n_levels, 3] = inner iterations / median passes / warps executed per level,
        info dict) of the last calc -- the inputs of the roofline accounting."""
        bc = getattr(self, "_batch_counters", None)
        if bc is not None:              # a calc_batch that was cut into several scheduler runs: counters of all of them
            c, infos = bc
            info = dict(infos[-1])
            for k in ("n_pairs", "device_ms", "pyramid_ms", "solver_ms", "solver_launches", "kernel_launches",
                      "double_steps", "double_steps_discarded"):
                info[k] = sum(i[k] for i in infos)
            info["scheduler_runs"] = len(infos)
            return c, info
        st = _lib.TeeflowStats()
        self._check(self._lib.teeflow_get_stats(self._h, C.byref(st)))
        n = st.n_pairs
        buf = np.zeros((max(n, 1), _lib.TEEFLOW_MAX_LEVELS, 3), np.int32)
        self._check(self._lib.teeflow_get_counters(self._h, buf.ctypes.data_as(C.POINTER(C.c_int32)), max(n, 1)))
        info = {k: getattr(st, k) for k, _ in _lib.TeeflowStats._fields_ if k != "reserved"}
        return buf[:n, :st.n_levels].copy(), info

    def flow_stats(self):
        """Diagnostic builds (-DTEEFLOW_FLOW_STATS=1): (enabled, cycles[16], counts[16]) of the last dataflow run."""
        buf = (C.c_uint64 * 32)()
        on = self._check(self._lib.teeflow_get_flow_stats(self._h, buf))
        return bool(on), np.array(buf[:16], np.float64), np.array(buf[16:], np.float64)

    def time_launches(self, n: int) -> None:
        """Diagnostics: time the first n solver launches of every following calc (teeflow_time_launches)."""
        self._check(self._lib.teeflow_time_launches(self._h, int(n)))

    def launch_times_ms(self) -> np.ndarray:
        buf = (C.c_float * 64)()
        n = self._check(self._lib.teeflow_get_launch_times(self._h, buf, 64))
        return np.array(buf[:n], np.float32)

    def level_sizes(self, H: int, W: int):
        hs = (C.c_int32 * _lib.TEEFLOW_MAX_LEVELS)()
        ws = (C.c_int32 * _lib.TEEFLOW_MAX_LEVELS)()
        L = self._check(self._lib.teeflow_level_sizes(self._h, H, W, hs, ws))
        return [(hs[i], ws[i]) for i in range(L)]


def createOptFlow_DualTVL1(**params) -> TVL1Engine:
    """Spelling of the reference's factory call (cv2.optflow.createOptFlow_DualTVL1, :577)."""
    return TVL1Engine(**params)
