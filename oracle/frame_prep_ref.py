"""CPU restatement of the reference's TV-L1 input stage for `no_saliency=True` (TEST INFRASTRUCTURE ONLY).

    saliency_2 = img2uint8(rgb2gray(nparr[i]))            optical_flow/calculate_optical_flow.py:588
    img2uint8(img) = img_as_ubyte((img - min) / max)      optical_flow/optical_flow_utils.py:30-31

`skimage` is not installed in this image; its two functions are restated from their documented behaviour:
`skimage.color.rgb2gray` converts to float (uint8 -> /255) and takes 0.2125 R + 0.7154 G + 0.0721 B;
`img_as_ubyte` of a float image in [0, 1] is rint(x * 255).  Pinned by tests/test_flow_cpu.py on hand-computed
values; the product's GPU kernels (teeflow_prepare_frames) are compared against this module, never the other way.
"""
from __future__ import annotations

import numpy as np


def rgb2gray(rgb: np.ndarray) -> np.ndarray:
    a = np.asarray(rgb)
    a = a.astype(np.float64) / 255.0 if a.dtype == np.uint8 else a.astype(np.float64)
    # evaluation order of the weighted sum: (R*c0 + G*c1) + B*c2 (numpy's `@` over the last axis sums left to right)
    return (a[..., 0] * 0.2125 + a[..., 1] * 0.7154) + a[..., 2] * 0.0721


def img2uint8(img: np.ndarray) -> np.ndarray:
    x = (img - np.min(img)) / np.max(img)          # sic: divides by max, not by the range
    return np.clip(np.rint(x * 255.0), 0, 255).astype(np.uint8)


def prepare_frames(nparr: np.ndarray) -> np.ndarray:
    """(N,H,W[,3]) -> (N,H,W) uint8: gray2rgb when 3-D (:536), then img2uint8(rgb2gray(frame)) per frame."""
    if nparr.ndim == 3:
        nparr = np.stack([nparr] * 3, axis=-1)
    return np.stack([img2uint8(rgb2gray(nparr[i])) for i in range(nparr.shape[0])])
