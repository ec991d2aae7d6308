"""numpy restatement of the reference post-processing chain (TEST INFRASTRUCTURE).

Pinned parts
  * anchors_for_shape  -- golden files onnx-models/anchors_256.txt,
    translation_anchors_{256,512}.txt (copied as tests/golden/*.npy) and the reference
    function itself (tests/test_oracle_pins.py).
  * decode (bbox_transform_inv / ClipBoxes / translation_transform_inv / CalculateTxTy) --
    the reference's own TF-free code (hmdegopose/layers.py lines 1-259) executed on the same
    inputs (tests/test_oracle_pins.py) + tests/golden/post_golden_*.npz.

PARITY UNPINNED parts (third-party arithmetic absent from /root/reference and from this image)
  * tensorflow (requirements.txt: bare ``tensorflow``, unpinned): tf.image.non_max_suppression,
    tf.nn.top_k, tf.where/gather/pad as called from hmdegopose/layers.py:264-400.  Restated
    from the published TF 2.x algorithm (core/kernels/image/non_max_suppression_op.cc):
    candidates ordered by score descending, ties -> lower index first; IoU with per-axis
    min/max normalised corners, 0 if either area <= 0; drop iff IoU > threshold; stop at
    max_output_size.  top_k: descending, ties -> lower index first.  Cross-checked against
    torchvision.ops.nms (same formula, strict >, stable order) in tests/test_oracle_post.py.
  * OpenCvSharp4 4.5.2 / OpenCV 4.5.2 cv::dnn::NMSBoxes (Program.cs:901): its outcome never
    changes the C# result (the arg-max-score candidate always survives NMS), see csharp_best().
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# anchors  (generators/utils/anchors.py)
# --------------------------------------------------------------------------------------
SIZES = (32, 64, 128, 256, 512)          # anchors.py:59
STRIDES = (8, 16, 32, 64, 128)           # anchors.py:60
RATIOS = np.array([1, 0.5, 2], dtype=np.float32)                                   # anchors.py:63
SCALES = np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)], dtype=np.float32)  # anchors.py:64
LEVELS = (3, 4, 5, 6, 7)


def base_anchors(size: float) -> np.ndarray:
    """generate_anchors (anchors.py:385-419): 9 boxes, a = scale*3 + ratio, float64 maths on
    float32-rounded ratios/scales."""
    out = np.zeros((9, 4), dtype=np.float64)
    for si in range(3):
        for ri in range(3):
            a = si * 3 + ri
            side = np.float64(size) * np.float64(SCALES[si])
            area = side * side
            w = np.sqrt(area / np.float64(RATIOS[ri]))
            h = w * np.float64(RATIOS[ri])
            # reference: x1 = 0 - w*0.5, x2 = w - w*0.5 (anchors.py:414-415)
            out[a] = (0.0 - w * 0.5, 0.0 - h * 0.5, w - w * 0.5, h - h * 0.5)
    return out


def level_shapes(size_hw: Tuple[int, int]) -> List[Tuple[int, int]]:
    """guess_shapes (anchors.py:257-270)."""
    return [((size_hw[0] + 2 ** l - 1) // 2 ** l, (size_hw[1] + 2 ** l - 1) // 2 ** l) for l in LEVELS]


def anchors_for_shape(size_hw: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
    """anchors_for_shape (anchors.py:273-318) + shift (:321-347) + translation_shift (:350-382).
    Row order: level -> y -> x -> a.  Returns float32 (N,4) x1,y1,x2,y2 and (N,3) cx,cy,stride."""
    boxes, trans = [], []
    for (h, w), size, stride in zip(level_shapes(size_hw), SIZES, STRIDES):
        base = base_anchors(size)
        cx = (np.arange(w, dtype=np.float64) + 0.5) * stride
        cy = (np.arange(h, dtype=np.float64) + 0.5) * stride
        gx, gy = np.meshgrid(cx, cy)                      # (h, w), row = y
        shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)  # (K,4)
        boxes.append((shifts[:, None, :] + base[None, :, :]).reshape(-1, 4))
        t = np.repeat(shifts[:, None, :2], 9, axis=1).reshape(-1, 2)
        trans.append(np.concatenate([t, np.full((t.shape[0], 1), float(stride))], axis=1))
    return (np.concatenate(boxes, 0).astype(np.float32), np.concatenate(trans, 0).astype(np.float32))


# --------------------------------------------------------------------------------------
# decode  (hmdegopose/layers.py:122-249, loss.py:12-51) -- all float32, op order as written
# --------------------------------------------------------------------------------------
def decode_boxes(anchors: np.ndarray, regression: np.ndarray, width: int, height: int) -> np.ndarray:
    """bbox_transform_inv (layers.py:169-200) then ClipBoxes (layers.py:122-136).
    anchors (N,4) f32; regression (B,N,4) = (ty,tx,th,tw)."""
    a = anchors.astype(f32)[None]
    d = regression.astype(f32)
    cxa = (a[..., 0] + a[..., 2]) / f32(2)
    cya = (a[..., 1] + a[..., 3]) / f32(2)
    wa = a[..., 2] - a[..., 0]
    ha = a[..., 3] - a[..., 1]
    ty, tx, th, tw = d[..., 0], d[..., 1], d[..., 2], d[..., 3]
    w = np.exp(tw) * wa
    h = np.exp(th) * ha
    cy = ty * ha + cya
    cx = tx * wa + cxa
    ymin = cy - h / f32(2.)
    xmin = cx - w / f32(2.)
    ymax = cy + h / f32(2.)
    xmax = cx + w / f32(2.)
    x1 = np.clip(xmin, f32(0), f32(width - 1))
    y1 = np.clip(ymin, f32(0), f32(height - 1))
    x2 = np.clip(xmax, f32(0), f32(width - 1))
    y2 = np.clip(ymax, f32(0), f32(height - 1))
    return np.stack([x1, y1, x2, y2], axis=-1).astype(f32)


def decode_translation(tanchors: np.ndarray, raw: np.ndarray, cam: np.ndarray) -> np.ndarray:
    """translation_transform_inv (layers.py:142-166) then CalculateTxTy (layers.py:212-249).
    tanchors (N,3) = (cx,cy,stride); raw (B,N,3) = (dx,dy,Tz); cam (B,6) =
    [fx,fy,px,py,tz_scale,image_scale].  Returns (B,N,3) = (tx,ty,tz)."""
    ta = tanchors.astype(f32)[None]
    d = raw.astype(f32)
    c = cam.astype(f32)
    stride = ta[..., 2]
    x = ta[..., 0] + d[..., 0] * stride
    y = ta[..., 1] + d[..., 1] * stride
    tz_raw = d[..., 2]
    fx, fy, px, py, tzs, ims = (c[:, i:i + 1] for i in range(6))
    x = x / ims
    y = y / ims
    tz = tz_raw * tzs
    x = x - px
    y = y - py
    tx = (x * tz) / fx
    ty = (y * tz) / fy
    return np.stack([tx, ty, tz], axis=-1).astype(f32)


# --------------------------------------------------------------------------------------
# TensorFlow filter_detections (hmdegopose/layers.py:264-400)   [parity unpinned, see header]
# --------------------------------------------------------------------------------------
def iou_tf(bi: np.ndarray, bj: np.ndarray) -> np.float32:
    """TF non_max_suppression_op.cc IOU(): fp32, corners normalised with min/max per axis."""
    y0i, y1i = min(bi[0], bi[2]), max(bi[0], bi[2])
    x0i, x1i = min(bi[1], bi[3]), max(bi[1], bi[3])
    y0j, y1j = min(bj[0], bj[2]), max(bj[0], bj[2])
    x0j, x1j = min(bj[1], bj[3]), max(bj[1], bj[3])
    area_i = f32(f32(y1i - y0i) * f32(x1i - x0i))
    area_j = f32(f32(y1j - y0j) * f32(x1j - x0j))
    if area_i <= 0 or area_j <= 0:
        return f32(0.0)
    ih = max(f32(min(y1i, y1j) - max(y0i, y0j)), f32(0.0))
    iw = max(f32(min(x1i, x1j) - max(x0i, x0j)), f32(0.0))
    inter = f32(ih * iw)
    return f32(inter / f32(f32(area_i + area_j) - inter))


def nms_tf(boxes: np.ndarray, scores: np.ndarray, max_out: int, iou_thr: float) -> np.ndarray:
    """tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold).
    Returns indices into boxes, in selection order."""
    boxes = boxes.astype(f32)
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))  # score desc, idx asc
    thr = f32(iou_thr)
    sel: List[int] = []
    for i in order:
        if len(sel) >= max_out:
            break
        keep = True
        for j in reversed(sel):                       # TF iterates selected boxes backwards
            if iou_tf(boxes[i], boxes[j]) > thr:
                keep = False
                break
        if keep:
            sel.append(int(i))
    return np.asarray(sel, dtype=np.int64)


def nms_tf_fast(boxes: np.ndarray, scores: np.ndarray, max_out: int, iou_thr: float) -> np.ndarray:
    """Vectorised equivalent of nms_tf (same fp32 op order) for large candidate sets."""
    b = boxes.astype(f32)
    n = len(scores)
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    lo0 = np.minimum(b[:, 0], b[:, 2]); hi0 = np.maximum(b[:, 0], b[:, 2])
    lo1 = np.minimum(b[:, 1], b[:, 3]); hi1 = np.maximum(b[:, 1], b[:, 3])
    area = ((hi0 - lo0) * (hi1 - lo1)).astype(f32)
    thr = f32(iou_thr)
    alive = np.ones(n, dtype=bool)
    sel: List[int] = []
    for i in order:
        if not alive[i]:
            continue
        sel.append(int(i))
        if len(sel) >= max_out:
            break
        d0 = np.maximum(np.minimum(hi0[i], hi0) - np.maximum(lo0[i], lo0), f32(0)).astype(f32)
        d1 = np.maximum(np.minimum(hi1[i], hi1) - np.maximum(lo1[i], lo1), f32(0)).astype(f32)
        inter = (d0 * d1).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = (inter / ((area[i] + area).astype(f32) - inter).astype(f32)).astype(f32)
        iou = np.where((area[i] <= 0) | (area <= 0), f32(0), iou)
        alive &= ~(iou > thr)
    return np.asarray(sel, dtype=np.int64)


def filter_detections(boxes: np.ndarray, classification: np.ndarray, rotation: np.ndarray,
                      translation: np.ndarray, hand: np.ndarray, score_threshold: float = 0.5,
                      max_detections: int = 100, nms_threshold: float = 0.5,
                      fast: bool = True) -> Dict[str, np.ndarray]:
    """filter_detections for ONE image (layers.py:264-400) with class_specific_filter=True, nms=True
    and the constants TrainModelWithLoss passes (train.py:78-81).

    Returns dict(boxes[100,4], scores[100], labels[100] int32, rotation[100,3], translation[100,3],
    hand[100,H], anchor_idx[100] int32) padded with -1."""
    nms = nms_tf_fast if fast else nms_tf
    thr = f32(score_threshold)
    pairs_a: List[np.ndarray] = []
    pairs_c: List[np.ndarray] = []
    for c in range(classification.shape[1]):
        s = classification[:, c].astype(f32)
        idx = np.nonzero(s > thr)[0]                                 # tf.where(greater), ascending
        keep = nms(boxes[idx], s[idx], max_detections, nms_threshold)
        pairs_a.append(idx[keep])
        pairs_c.append(np.full(len(keep), c, dtype=np.int64))
    anchors_idx = np.concatenate(pairs_a) if pairs_a else np.zeros((0,), np.int64)
    labels = np.concatenate(pairs_c) if pairs_c else np.zeros((0,), np.int64)
    sc = classification[anchors_idx, labels].astype(f32)
    k = min(max_detections, len(sc))
    top = np.lexsort((np.arange(len(sc)), -sc.astype(np.float64)))[:k]   # tf.nn.top_k: ties -> lower index
    a = anchors_idx[top]

    def pad(x, width=None):
        shape = (max_detections,) if width is None else (max_detections, width)
        out = np.full(shape, -1, dtype=x.dtype)
        out[:k] = x
        return out

    return {
        "boxes": pad(boxes[a].astype(f32), 4),
        "scores": pad(sc[top]),
        "labels": pad(labels[top].astype(np.int32)),
        "rotation": pad(rotation[a].astype(f32), rotation.shape[1]),
        "translation": pad(translation[a].astype(f32), translation.shape[1]),
        "hand": pad(hand[a].astype(f32), hand.shape[1]),
        "anchor_idx": pad(a.astype(np.int32)),
        "count": np.int32(k),
    }


def detect(regression, classification, rotation, translation_raw, hand, cam, size: int,
           score_threshold: float = 0.5, max_detections: int = 100, nms_threshold: float = 0.5):
    """TrainModelWithLoss.forward inference branch after the network (train.py:34-83), for every
    image of the batch (the reference keeps only the last one: layers.py:466-482)."""
    anchors, tanchors = anchors_for_shape((size, size))
    boxes = decode_boxes(anchors, regression, size, size)
    trans = decode_translation(tanchors, translation_raw, cam)
    return [filter_detections(boxes[b], classification[b], rotation[b], trans[b], hand[b],
                              score_threshold, max_detections, nms_threshold)
            for b in range(regression.shape[0])]


# --------------------------------------------------------------------------------------
# C# receiver semantics (unity-sandbox/WebRTCNetCoreSandbox/Program.cs:488-960)
# --------------------------------------------------------------------------------------
def csharp_decode_boxes(anchors: np.ndarray, regression: np.ndarray, width: int, height: int) -> np.ndarray:
    """regress_boxes + clip_boxes (Program.cs:654-784) for ONE image, float32.
    NOTE the C# twin uses regression column 0 for the x-centre and column 1 for the y-centre
    (Program.cs:672,686,739-740), the opposite of layers.py:186; and clamps ymax with ``width``
    (Program.cs:773).  Reproduced verbatim."""
    a = anchors.astype(f32)
    d = regression.astype(f32)
    tx, ty, th, tw = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    cxa = (a[:, 0] + a[:, 2]) / f32(2)
    cya = (a[:, 1] + a[:, 3]) / f32(2)
    wa = a[:, 2] - a[:, 0]
    ha = a[:, 3] - a[:, 1]
    w = np.exp(tw) * wa
    h = np.exp(th) * ha
    cy = ty * ha + cya
    cx = tx * wa + cxa
    ymin = cy - h / f32(2)
    xmin = cx - w / f32(2)
    ymax = cy + h / f32(2)
    xmax = cx + w / f32(2)
    xmin = np.minimum(np.maximum(xmin, f32(0)), f32(width - 1))
    ymin = np.minimum(np.maximum(ymin, f32(0)), f32(height - 1))
    xmax = np.minimum(np.maximum(xmax, f32(0)), f32(width - 1))
    ymax = np.minimum(np.maximum(ymax, f32(0)), f32(width - 1))
    return np.stack([xmin, ymin, xmax, ymax], axis=1).astype(f32)


def csharp_best(regression, classification, rotation, translation_raw, cam, size: int,
                score_threshold: float = 0.5) -> np.ndarray:
    """What Program.cs:247-270 hands to the pose packet for ONE image: out10 =
    [score, rect.X, rect.Y, rect.Width, rect.Height, rx, ry, rz (rad), tx, ty, tz (m)][:10]... laid out as
    [score, X, Y, W, H, rx, ry, rz, tx, ty] would drop tz, so out has 11 floats:
    [score, X, Y, W, H, rx, ry, rz, tx, ty, tz].

    Program.cs:786-960: keep score > 0.5; Rect((int)x1,(int)y1,(int)x2,(int)y2) (so Width/Height
    hold x2/y2, Program.cs:840-845); NMSBoxes(top_k=10); choose the max-score survivor with a strict
    '>' scan.  The highest-scoring candidate always survives NMS and NMSBoxes orders candidates with a
    stable descending sort, so the result is the arg-max-score anchor, ties -> lowest anchor index.
    Zeros if nothing passes (Program.cs:929-932)."""
    out = np.zeros(11, dtype=f32)
    s = classification[:, 0].astype(f32)
    idx = np.nonzero(s > f32(score_threshold))[0]
    if len(idx) == 0:
        return out
    best = idx[np.lexsort((idx, -s[idx].astype(np.float64)))[0]]
    anchors, tanchors = anchors_for_shape((size, size))
    box = csharp_decode_boxes(anchors[best:best + 1], regression[best:best + 1], size, size)[0]
    tr = decode_translation(tanchors[best:best + 1], translation_raw[None, best:best + 1], cam[None])[0, 0]
    out[0] = s[best]
    out[1:5] = np.trunc(box).astype(f32)                 # (int) cast truncates toward zero
    out[5:8] = rotation[best].astype(f32) * f32(np.pi)   # residual_rotation *= (float)Math.PI
    out[8:11] = tr * f32(1 / 1000.0)                     # residual_translation *= (1 / 1000.0f)
    return out


def csharp_pose_packet(best11) -> bytes:
    """Program.cs:279-292: floatArray = {rvec.X, rvec.Y, rvec.Z, t.X, t.Y, t.Z}; Buffer.BlockCopy -> 24 bytes
    (little-endian fp32 on every platform .NET Core runs on); PoseDataChannel.cs:80-108 copies them back."""
    import struct
    b = np.asarray(best11, np.float32).reshape(-1)
    return struct.pack("<6f", *[float(v) for v in b[5:11]])
