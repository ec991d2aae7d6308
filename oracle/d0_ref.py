"""EfficientDet-d0 detection variant (BASELINE.json configs[4]) -- numpy restatement
(TEST INFRASTRUCTURE).

  anchors      efficientdet/utils.py:76-139   Anchors.forward: (y1, x1, y2, x2), anchor_scale 4, ratios
               (1,1),(1.4,0.7),(0.7,1.4), scales 2^(0,1/3,2/3); float64 maths, cast to float32
  decode       efficientdet/utils.py:7-35     BBoxTransform -> (xmin, ymin, xmax, ymax)
  clip         efficientdet/utils.py:38-52    x1,y1 >= 0; x2 <= W-1; y2 <= H-1
  postprocess  utils/utils.py:90-128          score = max_c; keep > threshold; torchvision batched_nms

Pinned: anchors / decode / clip against the reference classes themselves (importable), postprocess against the
reference function body executed with this image's torchvision (tests/test_oracle_pins.py).  Third-party arithmetic:
torchvision 0.9.2 (README.md:120) ``batched_nms`` = offset every box by ``class_id * (max_coordinate + 1)`` and run
plain NMS (IoU = inter / (area_i + area_j - inter) on the OFFSET boxes, suppress iff > threshold, score-descending).
"""
from __future__ import annotations

import itertools
from typing import Dict, List

import numpy as np

f32 = np.float32
SCALES = np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)])
RATIOS = [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)]
STRIDES = [8, 16, 32, 64, 128]


def anchors(size: int, anchor_scale: float = 4.0) -> np.ndarray:
    """Anchors.forward (efficientdet/utils.py:76-139): (N,4) float32 (y1,x1,y2,x2); order level -> (y,x) -> (scale,ratio)."""
    out = []
    for stride in STRIDES:
        lvl = []
        for scale, ratio in itertools.product(SCALES, RATIOS):
            base = anchor_scale * stride * scale
            ax2, ay2 = base * ratio[0] / 2.0, base * ratio[1] / 2.0
            x = np.arange(stride / 2, size, stride)
            y = np.arange(stride / 2, size, stride)
            xv, yv = np.meshgrid(x, y)
            xv, yv = xv.reshape(-1), yv.reshape(-1)
            boxes = np.vstack((yv - ay2, xv - ax2, yv + ay2, xv + ax2)).swapaxes(0, 1)
            lvl.append(boxes[:, None, :])
        out.append(np.concatenate(lvl, axis=1).reshape(-1, 4))
    return np.vstack(out).astype(np.float32)


def decode_clip(anc: np.ndarray, regression: np.ndarray, width: int, height: int) -> np.ndarray:
    """BBoxTransform + ClipBoxes, float32 in the reference's op order.  regression (B,N,4) = (dy,dx,dh,dw)."""
    a = anc.astype(f32)[None]
    r = regression.astype(f32)
    yca = (a[..., 0] + a[..., 2]) / f32(2)
    xca = (a[..., 1] + a[..., 3]) / f32(2)
    ha = a[..., 2] - a[..., 0]
    wa = a[..., 3] - a[..., 1]
    w = np.exp(r[..., 3]) * wa
    h = np.exp(r[..., 2]) * ha
    yc = r[..., 0] * ha + yca
    xc = r[..., 1] * wa + xca
    ymin = yc - h / f32(2.)
    xmin = xc - w / f32(2.)
    ymax = yc + h / f32(2.)
    xmax = xc + w / f32(2.)
    return np.stack([np.maximum(xmin, f32(0)), np.maximum(ymin, f32(0)),
                     np.minimum(xmax, f32(width - 1)), np.minimum(ymax, f32(height - 1))], axis=2).astype(f32)


def nms_offset(boxes: np.ndarray, scores: np.ndarray, classes: np.ndarray, iou_thr: float) -> np.ndarray:
    """torchvision 0.9.2 batched_nms (coordinate trick) + nms_kernel.cpp, float32."""
    n = len(scores)
    if n == 0:
        return np.zeros((0,), np.int64)
    max_coord = boxes.max()
    off = (classes.astype(f32) * f32(max_coord + f32(1))).astype(f32)
    b = (boxes.astype(f32) + off[:, None]).astype(f32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(f32)
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    dead = np.zeros(n, bool)
    keep: List[int] = []
    thr = f32(iou_thr)
    for i in order:
        if dead[i]:
            continue
        keep.append(int(i))
        w = np.maximum(f32(0), np.minimum(x2[i], x2) - np.maximum(x1[i], x1)).astype(f32)
        h = np.maximum(f32(0), np.minimum(y2[i], y2) - np.maximum(y1[i], y1)).astype(f32)
        inter = (w * h).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = (inter / ((areas[i] + areas).astype(f32) - inter).astype(f32)).astype(f32)
        dead |= ovr > thr
    return np.asarray(keep, np.int64)


def postprocess(regression: np.ndarray, classification: np.ndarray, size: int, threshold: float = 0.2,
                iou_threshold: float = 0.2) -> List[Dict[str, np.ndarray]]:
    """utils/utils.py:90-128 for every image: dict(rois (k,4) x1,y1,x2,y2, class_ids (k,), scores (k,), anchor_idx (k,))
    in NMS keep order (score descending)."""
    anc = anchors(size)
    boxes = decode_clip(anc, regression, size, size)
    out = []
    for b in range(regression.shape[0]):
        cls = classification[b].astype(f32)
        score = cls.max(axis=1)
        sel = np.nonzero(score > f32(threshold))[0]
        if len(sel) == 0:
            out.append({"rois": np.zeros((0, 4), f32), "class_ids": np.zeros((0,), np.int64),
                        "scores": np.zeros((0,), f32), "anchor_idx": np.zeros((0,), np.int64)})
            continue
        classes = cls[sel].argmax(axis=1)
        keep = nms_offset(boxes[b, sel], score[sel], classes, iou_threshold)
        out.append({"rois": boxes[b, sel][keep], "class_ids": classes[keep].astype(np.int64),
                    "scores": score[sel][keep], "anchor_idx": sel[keep].astype(np.int64)})
    return out
