"""Frame pre-processing of the reference -- numpy restatement, TEST INFRASTRUCTURE.

  preprocess_image   generators/colibri_common.py:622-656 (C# twin: WebRTCNetCoreSandbox/Program.cs:397-445):
                     aspect-preserving resize of the uint8 RGB frame so that its long side is the network size,
                     /255, ImageNet mean/std, zero pad bottom/right to S x S.
  resize_linear_u8   the third-party step inside it: ``cv2.resize(image, (w, h))`` = INTER_LINEAR on 8-bit data
                     (OpenCV resize.cpp: 11-bit fixed-point coefficients ``cvRound(f * 2048)``, horizontal pass in int32,
                     vertical pass ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``).

Pinned in tests/test_oracle_preprocess.py against cv2 4.13 (this image) and against the reference method itself:
bit-exact for every down-scaling case tried (camera frames); when UP-scaling to a width that is not a multiple of
OpenCV's SIMD step a few tail columns (< 0.2 % of the pixels) differ by one LSB (its scalar tail rounds differently).

Arithmetic of the normalisation as numpy executes the reference lines: ``image /= 255.`` stays float32;
``image -= mean`` and ``image /= std`` (Python lists -> float64 arrays) are evaluated in float64 and rounded back to
float32 by the in-place assignment.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def resized_shape(height: int, width: int, size: int) -> Tuple[int, int, float]:
    """colibri_common.py:633-640 -> (resized_height, resized_width, scale)."""
    if height > width:
        scale = size / height
        return size, int(width * scale), scale
    scale = size / width
    return int(height * scale), size, scale


def _coeffs(dn: int, sn: int):
    scale = sn / dn
    d = np.arange(dn)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(np.float32)
    lo = s < 0
    f[lo] = 0
    s[lo] = 0
    hi = s >= sn - 1
    f[hi] = 0
    s[hi] = sn - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    return s, np.minimum(s + 1, sn - 1), a0, a1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) for uint8 HWC input, default INTER_LINEAR."""
    sh, sw = img.shape[:2]
    if (sh, sw) == (dh, dw):
        return img.copy()
    sx, sx1, ax0, ax1 = _coeffs(dw, sw)
    sy, sy1, ay0, ay1 = _coeffs(dh, sh)
    i64 = img.astype(np.int64)
    hor = i64[:, sx, :] * ax0[None, :, None] + i64[:, sx1, :] * ax1[None, :, None]
    s0, s1 = hor[sy], hor[sy1]
    out = (((ay0[:, None, None] * (s0 >> 4)) >> 16) + ((ay1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess_image(image: np.ndarray, size: int) -> Tuple[np.ndarray, float]:
    """colibri_common.py:622-656: uint8 RGB (H, W, 3) -> (float32 (S, S, 3), scale)."""
    h, w = image.shape[:2]
    rh, rw, scale = resized_shape(h, w, size)
    img = resize_linear_u8(image, rw, rh).astype(np.float32)
    img /= np.float32(255.0)
    img = (img.astype(np.float64) - np.asarray(MEAN, np.float64)).astype(np.float32)
    img = (img.astype(np.float64) / np.asarray(STD, np.float64)).astype(np.float32)
    out = np.zeros((size, size, 3), np.float32)
    out[:rh, :rw] = img
    return out, scale
