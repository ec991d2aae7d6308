"""Frame pre-processing of the reference -- numpy restatement, TEST INFRASTRUCTURE.

  preprocess_image   generators/colibri_common.py:622-656 (C# twin: WebRTCNetCoreSandbox/Program.cs:397-445):
                     aspect-preserving resize of the uint8 RGB frame so that its long side is the network size,
                     /255, ImageNet mean/std, zero pad bottom/right to S x S.
  resize_linear_u8   the third-party step inside it: ``cv2.resize(image, (w, h))`` = INTER_LINEAR on 8-bit data
                     (OpenCV resize.cpp: 11-bit fixed-point coefficients ``cvRound(f * 2048)``, horizontal pass in int32,
                     vertical pass ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``).

Pinned in tests/test_oracle_preprocess.py against cv2 4.13 (this image) and against the reference method itself:
bit-exact for down- and up-scaling (the border rows of an up-scaled image follow OpenCV's clamped-row rule, _coeffs).

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


def _coeffs(dn: int, sn: int, vertical: bool = False):
    """OpenCV resize.cpp coefficient tables.  Horizontal: a tap outside the image gets weight 0 (fx = 0 at the
    borders).  Vertical: the weights are kept and the ROW INDICES are clamped instead, so on the border rows of an
    up-scaled image both taps read the same row with two separately truncated products."""
    scale = sn / dn
    d = np.arange(dn)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) for uint8 HWC input, default INTER_LINEAR."""
    sh, sw = img.shape[:2]
    if (sh, sw) == (dh, dw):
        return img.copy()
    sx, sx1, ax0, ax1 = _coeffs(dw, sw)
    sy, sy1, ay0, ay1 = _coeffs(dh, sh, vertical=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# The C# receiver's frame path (unity-sandbox/WebRTCNetCoreSandbox/Program.cs:128-200, 381-445), I420 bytes in:
#   i420_to_bgr_yv12   Program.cs:146-160: the I420 buffer (Y, U, V planes) is wrapped as a (rows * 3/2) x cols
#                      single-channel Mat and converted with ``Cv2.CvtColor(.., YUV2BGR_YV12)`` -- YV12 expects
#                      (Y, V, U), so the two chroma planes are deliberately read swapped.  OpenCV's 4:2:0 conversion
#                      (imgproc/src/color_yuv.simd.hpp: ITU-R BT.601, 20-bit fixed point).
#   center_crop_and_rescale  Program.cs:381-395 (called with crop 256 -> 512 x 512 at :170-173): central ROI, then
#                      cv2.resize INTER_LINEAR.
#   resize_and_normalize_cs  Program.cs:397-445: aspect-preserving resize, ConvertTo CV_32F, Cv2.Divide(255.0f),
#                      Cv2.Subtract(mean), Cv2.Divide(std) -- OpenCV evaluates these on CV_32F data in FLOAT32 with the
#                      scalars converted to float32 (unlike the numpy path above, which goes through float64), zero pad.
# Pinned against cv2 4.13 executing the same calls (tests/test_oracle_preprocess.py).
# ---------------------------------------------------------------------------------------------------------------------
_CY, _CUB, _CUG, _CVG, _CVR, _SHIFT = 1220542, 2116026, -409993, -852492, 1673527, 20


def i420_to_bgr_yv12(buf: np.ndarray, height: int, width: int) -> np.ndarray:
    """cv2.cvtColor(buf.reshape(height * 3 // 2, width), COLOR_YUV2BGR_YV12) for even height / width: (H, W, 3) uint8."""
    flat = np.asarray(buf, np.uint8).reshape(-1)
    n = height * width
    y = flat[:n].reshape(height, width).astype(np.int64)
    first = flat[n:n + n // 4].reshape(height // 2, width // 2).astype(np.int64)    # read as V (it holds the I420 U plane)
    second = flat[n + n // 4:n + n // 2].reshape(height // 2, width // 2).astype(np.int64)
    vv = np.repeat(np.repeat(first, 2, 0), 2, 1) - 128
    uu = np.repeat(np.repeat(second, 2, 0), 2, 1) - 128
    yy = np.maximum(0, y - 16) * _CY
    half = 1 << (_SHIFT - 1)
    r = (yy + half + _CVR * vv) >> _SHIFT
    g = (yy + half + _CVG * vv + _CUG * uu) >> _SHIFT
    b = (yy + half + _CUB * uu) >> _SHIFT
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def center_crop_and_rescale(img: np.ndarray, crop: int, out_w: int, out_h: int) -> np.ndarray:
    off_w, off_h = (img.shape[1] - crop) // 2, (img.shape[0] - crop) // 2
    return resize_linear_u8(img[off_h:off_h + crop, off_w:off_w + crop], out_w, out_h)


def resize_and_normalize_cs(img: np.ndarray, size: int) -> Tuple[np.ndarray, float]:
    """Program.cs:397-445 on a uint8 (H, W, 3) Mat: float32 (S, S, 3) in the Mat's channel order, and the scale."""
    h, w = img.shape[:2]
    if h > w:
        scale = np.float32(size) / np.float32(h)
        rh, rw = size, int(np.float32(w) * scale)
    else:
        scale = np.float32(size) / np.float32(w)
        rh, rw = int(np.float32(h) * scale), size
    v = resize_linear_u8(img, rw, rh).astype(np.float32)
    v = v / np.float32(255.0)
    v = v - np.asarray(MEAN, np.float32)
    v = v / np.asarray(STD, np.float32)
    out = np.zeros((size, size, 3), np.float32)
    out[:rh, :rw] = v
    return out, float(scale)


def csharp_frame_to_tensor(i420: np.ndarray, height: int, width: int, size: int, crop: int = 256, mid: int = 512):
    """The receiver's whole pre-processing (Program.cs:137-200): I420 frame -> float32 (S, S, 3) tensor (HWC; the
    following BlobFromImage only transposes to CHW) and the resize scale."""
    bgr = i420_to_bgr_yv12(i420, height, width)
    return resize_and_normalize_cs(center_crop_and_rescale(bgr, crop, mid, mid), size)
