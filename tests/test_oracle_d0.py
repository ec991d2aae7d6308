"""EfficientDet-d0 variant: oracle/d0_ref.py against the golden vectors produced by the reference
code (oracle/make_golden_d0.py), and libhmdpose's host-side D0 anchor arithmetic against the oracle (no GPU)."""
import os

import numpy as np
import pytest

from oracle import d0_ref


@pytest.fixture(scope="module")
def gold(gold_dir):
    return np.load(os.path.join(gold_dir, "d0_golden_128.npz"))


def test_anchors_bit_exact(gold):
    a = d0_ref.anchors(int(gold["size"]))
    assert a.dtype == np.float32 and np.array_equal(a, gold["anchors"])


def test_decode_clip(gold):
    s = int(gold["size"])
    got = d0_ref.decode_clip(gold["anchors"], gold["regression"], s, s)
    assert np.abs(got - gold["boxes"]).max() <= 1e-4        # exp() is libm-dependent; everything else is exact
    assert got[..., :2].min() >= 0 and got[..., 2:].max() <= s - 1   # x2,y2 are not floored, x1,y1 not capped (utils.py:44-50)


def test_postprocess_matches_reference(gold):
    s = int(gold["size"])
    dets = d0_ref.postprocess(gold["regression"], gold["classification"], s, float(gold["threshold"]),
                              float(gold["iou_threshold"]))
    assert len(dets) == 3
    for b, d in enumerate(dets):
        assert np.array_equal(d["class_ids"], gold[f"class_ids_{b}"]), b
        assert np.array_equal(d["scores"], gold[f"scores_{b}"]), b
        assert d["rois"].shape == gold[f"rois_{b}"].shape
        if len(d["scores"]):
            assert np.abs(d["rois"] - gold[f"rois_{b}"]).max() <= 1e-4
            assert np.all(np.diff(d["scores"]) <= 0)            # keep order = score descending
    assert len(dets[2]["scores"]) == 0                          # the reference's empty branch


def test_class_offset_keeps_overlapping_boxes_of_different_classes():
    # two identical boxes: same class -> one survives, different classes -> both (batched_nms semantics)
    n = 9 * sum(((128 + (1 << l) - 1) >> l) ** 2 for l in range(3, 8))
    reg = np.zeros((1, n, 4), np.float32)
    cls = np.zeros((1, n, 3), np.float32)
    cls[0, 0, 1] = 0.9
    cls[0, 1, 1] = 0.8      # anchor 1 = same cell, different ratio; make the boxes identical through the regression
    a = d0_ref.anchors(128)
    ha, wa = a[:2, 2] - a[:2, 0], a[:2, 3] - a[:2, 1]
    reg[0, 1, 2], reg[0, 1, 3] = np.log(ha[0] / ha[1]), np.log(wa[0] / wa[1])
    same = d0_ref.postprocess(reg, cls, 128)[0]
    assert list(same["anchor_idx"]) == [0]
    cls[0, 1, 1], cls[0, 1, 2] = 0.0, 0.8
    diff = d0_ref.postprocess(reg, cls, 128)[0]
    assert list(diff["anchor_idx"]) == [0, 1] and list(diff["class_ids"]) == [1, 2]


@pytest.mark.parametrize("size", [128, 256, 512, 640])
def test_library_anchors_match_oracle(size):
    from hmd_ego_pose_b200.model import d0_anchors
    assert np.array_equal(d0_anchors(size), d0_ref.anchors(size))


@pytest.mark.refpin
def test_oracle_against_live_reference():
    ref = "/root/reference/pytorch-sandbox"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    import sys
    import torch
    sys.path.insert(0, ref)
    try:
        from efficientdet.utils import Anchors, BBoxTransform, ClipBoxes
    finally:
        sys.path.remove(ref)
    x = torch.zeros(2, 3, 512, 512)
    anc = Anchors(anchor_scale=4.0, pyramid_levels=[3, 4, 5, 6, 7])(x, x.dtype)
    assert np.array_equal(anc[0].numpy(), d0_ref.anchors(512))
    reg = torch.randn(2, anc.shape[1], 4, generator=torch.Generator().manual_seed(0)) * 0.4
    boxes = ClipBoxes()(BBoxTransform()(anc, reg), x).numpy()
    assert np.abs(d0_ref.decode_clip(d0_ref.anchors(512), reg.numpy(), 512, 512) - boxes).max() <= 2e-4


def test_class_offset_nms_matches_torchvision_batched_nms():
    """a20 third-party arithmetic: torchvision's batched_nms (coordinate trick below 4 000 boxes, the only path of the
    pinned 0.9.2) on random boxes with score ties and degenerate boxes -- same keep list, same order."""
    import torch
    from torchvision.ops.boxes import batched_nms
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.integers(1, 400))
        c = rng.random((n, 2)) * 500
        wh = rng.random((n, 2)) * 120
        if trial % 4 == 0:
            wh[rng.random(n) < 0.1] = 0.0                       # zero-area boxes
        boxes = np.concatenate([c, c + wh], axis=1).astype(np.float32)
        scores = rng.random(n).astype(np.float32)
        if trial % 3 == 0:
            scores = np.round(scores, 1)                        # ties: stable order by index
        classes = rng.integers(0, 5, n)
        thr = float(rng.choice([0.2, 0.5, 0.7]))
        want = batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(classes), thr).numpy()
        got = d0_ref.nms_offset(boxes, scores, classes, thr)
        if trial % 3 == 0:   # torch.sort is not stable for ties: compare as sets plus score order
            assert set(got.tolist()) == set(want.tolist()) and np.all(np.diff(scores[got]) <= 0)
        else:
            assert np.array_equal(got, want), trial
