"""Generate tests/golden/d0_golden_128.npz with the REFERENCE code (run in the build container only).

  anchors      : efficientdet.utils.Anchors()(image)                    (efficientdet/utils.py:76-139)
  boxes        : ClipBoxes(BBoxTransform(anchors, regression), image)   (efficientdet/utils.py:7-52)
  detections   : the body of utils.utils.postprocess (utils/utils.py:90-128), executed from its source text because
                 utils/utils.py imports webcolors / cv2 helpers that are absent here; torchvision.ops.boxes.batched_nms
                 of this image (coordinate trick below 4000 candidates = torchvision 0.9.2 arithmetic, README.md:120).

Usage:  python -m oracle.make_golden_d0
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference/pytorch-sandbox"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "d0_golden_128.npz")


def reference_postprocess():
    src = open(os.path.join(REF, "utils", "utils.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def postprocess("))
    end = next(i for i in range(start + 1, len(src)) if src[i] and not src[i].startswith((" ", "\t")))
    from torchvision.ops.boxes import batched_nms
    ns = {"torch": torch, "np": np, "batched_nms": batched_nms}
    exec("\n".join(src[start:end]), ns)
    return ns["postprocess"]


def inputs(size=128, batch=3, classes=7, seed=5):
    g = torch.Generator().manual_seed(seed)
    n = 9 * sum(((size + (1 << l) - 1) >> l) ** 2 for l in range(3, 8))
    reg = torch.randn(batch, n, 4, generator=g) * 0.35
    cls = torch.sigmoid(torch.randn(batch, n, classes, generator=g) * 1.2 - 3.0)
    cls[2] = cls[2] * 0.01          # third image: nothing passes the threshold (the reference's empty branch)
    return reg, cls


def main() -> None:
    sys.path.insert(0, REF)
    from efficientdet.utils import Anchors, BBoxTransform, ClipBoxes
    size = 128
    reg, cls = inputs(size)
    x = torch.zeros(reg.shape[0], 3, size, size)
    anc = Anchors(anchor_scale=4.0, pyramid_levels=[3, 4, 5, 6, 7])(x, x.dtype)
    boxes = ClipBoxes()(BBoxTransform()(anc, reg), x)
    post = reference_postprocess()
    thr, iou = 0.2, 0.2
    dets = post(x, anc, reg, cls, BBoxTransform(), ClipBoxes(), thr, iou)
    out = {"size": size, "threshold": thr, "iou_threshold": iou, "regression": reg.numpy(), "classification": cls.numpy(),
           "anchors": anc[0].numpy(), "boxes": boxes.numpy()}
    for b, d in enumerate(dets):
        out[f"rois_{b}"] = np.asarray(d["rois"], np.float32).reshape(-1, 4)
        out[f"class_ids_{b}"] = np.asarray(d["class_ids"], np.int64).reshape(-1)
        out[f"scores_{b}"] = np.asarray(d["scores"], np.float32).reshape(-1)
        print("image", b, "detections", len(out[f"scores_{b}"]))
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
