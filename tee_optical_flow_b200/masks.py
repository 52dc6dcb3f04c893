"""Mask post-processing and AV centroid on the GPU, under the reference's names.

    clean_mask(engine, arr, mode, config)       calculate_optical_flow.py:113-182 (moving_avg_mask :91-111,
                                                binary_fill_holes, remove_small_objects, 2-channel repeat, 'bkgd')
    calc_AV_centroid(engine, mask_arr, nframes) analysis.py:39-86 (label -> largest region -> centroid, fallbacks,
                                                Savitzky-Golay filter)

The connected-component labelling, hole filling, size filtering, temporal vote and the per-frame centroid run in
libteeflow.so; the reference's tiny per-clip bookkeeping (empty-frame fallback, scipy's savgol_filter) stays in
Python.  The class map itself comes from the SAM segmentor, which is outside this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from .config import OpticalFlowCalculationConfig, default_optical_flow_config
from .engine import TVL1Engine
from .exceptions import OpticalFlowCalculationError

MODE_CLASSES = {
    'A4C': {'lv_inner': 1, 'lv': 2, 'la_inner': 3, 'la': 4, 'rv_inner': 5, 'ra_inner': 6, 'rv': 7, 'ra': 8},
    'RVIO_2class': {'rv': 1, 'av': 2},
    'MouseRV_A4C': {'rv': 1, 'rv_inner': 2},
}


def _cuda_u8(engine: TVL1Engine, a):
    import torch
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    dev = torch.device("cuda", engine.device)
    if t.dtype == torch.bool:
        return t.to(dev).contiguous().view(torch.uint8)
    return t.to(dev, torch.uint8).contiguous()


def clean_label(engine: TVL1Engine, class_map, class_id: int, window: int = 4, threshold: float = 0.49,
                min_size: int = 500):
    """one label of clean_mask: (arr == class_id) -> moving average vote -> fill holes -> remove small objects.
    class_map: (N,H,W) uint8 (numpy or CUDA tensor).  Returns a CUDA bool tensor (N,H,W)."""
    import torch
    cm = _cuda_u8(engine, class_map)
    if cm.dim() != 3:
        raise OpticalFlowCalculationError("class map must be (N, H, W)")
    N, H, W = cm.shape
    out = torch.empty((N, H, W), dtype=torch.uint8, device=cm.device)
    stream = torch.cuda.current_stream(cm.device).cuda_stream
    engine._check(engine._lib.teeflow_clean_masks(engine._h, cm.data_ptr(), N, H, W, int(class_id), int(window),
                                                  float(threshold), int(min_size), out.data_ptr(), C.c_void_p(stream)))
    return out.view(torch.bool)


def clean_mask(engine: TVL1Engine, arr, mode: str = 'A4C', verbose: bool = False,
               config: Optional[OpticalFlowCalculationConfig] = None) -> Optional[Dict[str, np.ndarray]]:
    """Reference semantics: returns {label: (N,H,W,2) bool, ..., 'bkgd': (N,H,W,2) bool} or None for a bad mode."""
    import torch
    if config is None:
        config = default_optical_flow_config()
    if mode not in MODE_CLASSES:
        return None
    cm = _cuda_u8(engine, arr)
    agg = torch.zeros(cm.shape, dtype=torch.bool, device=cm.device)
    out: Dict[str, np.ndarray] = {}
    for label, cid in MODE_CLASSES[mode].items():
        # note: the reference calls moving_avg_mask WITHOUT the config, i.e. with its defaults n=4, threshold=0.49
        m = clean_label(engine, cm, cid, 4, 0.49, config.min_mask_size)
        agg |= m
        out[label] = np.repeat(m.cpu().numpy()[..., None], 2, axis=3)
    out['bkgd'] = np.repeat((~agg).cpu().numpy()[..., None], 2, axis=3)
    return out


def calc_AV_centroid(engine: TVL1Engine, mask_arr, nframes: int, filter: bool = True, savgol_window: int = 10,
                     savgol_poly: int = 4, verbose: bool = False):
    """analysis.py:39-86.  mask_arr: (N,H,W,C) bool; returns the centroid list (row, col) per frame."""
    import torch
    m = _cuda_u8(engine, mask_arr)
    if m.dim() != 4:
        raise OpticalFlowCalculationError("mask array must be (N, H, W, C)")
    N, H, W, Cn = m.shape
    if not (1 <= nframes <= N):
        raise OpticalFlowCalculationError("nframes out of range")
    cent = np.empty((nframes, 2), np.float64)
    ncomp = np.empty(nframes, np.int32)
    stream = torch.cuda.current_stream(m.device).cuda_stream
    engine._check(engine._lib.teeflow_av_centroids(engine._h, m.data_ptr(), Cn, nframes, H, W,
                                                   cent.ctypes.data_as(C.POINTER(C.c_double)),
                                                   ncomp.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(stream)))
    centroid_list = []
    for i in range(nframes):
        if ncomp[i] >= 1:
            centroid_list.append((cent[i, 0], cent[i, 1]))
        else:
            if len(centroid_list) > 0:
                centroid_list.append(centroid_list[i - 1])        # copy previous if empty
            else:
                centroid_list.append((mask_arr.shape[1] / 2, mask_arr.shape[2] / 2))
            print('WARNING: EMPTY MASK at Frame ', i)
    if filter:
        if len(centroid_list) < savgol_window:
            print('ERROR: Cannot apply savgol filter! List smaller than window')
        else:
            from scipy.signal import savgol_filter
            centroid_list = savgol_filter(centroid_list, savgol_window, savgol_poly, axis=0)
    return centroid_list
