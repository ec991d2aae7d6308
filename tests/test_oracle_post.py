"""Post-processing oracle: TF/OpenCV semantics restated (parity unpinned) -- internal consistency and
cross-checks against torchvision.ops.nms (same IoU formula, strict '>', stable descending order)."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import postprocess_ref as pp


def _random_boxes(rng, n, size=256.0):
    c = rng.random((n, 2)) * size
    wh = rng.random((n, 2)) * 60 + 2
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).clip(0, size - 1).astype(np.float32)
    return b


def test_nms_fast_equals_sequential_and_torchvision():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 300, 1500):
        boxes = _random_boxes(rng, n)
        scores = (rng.random(n) * 0.5 + 0.5).astype(np.float32)
        if n > 10:
            scores[5] = scores[9]          # exact tie -> lower index first
            boxes[3, 2] = boxes[3, 0]      # zero-area box: IoU 0 with everything
        a = pp.nms_tf(boxes, scores, 100, 0.5)
        b = pp.nms_tf_fast(boxes, scores, 100, 0.5)
        assert np.array_equal(a, b)
        if n:
            tv = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()[:100]
            assert np.array_equal(a, tv)


def test_filter_detections_padding_topk_and_classes():
    rng = np.random.default_rng(1)
    n, c = 2000, 3
    boxes = _random_boxes(rng, n)
    cls = rng.random((n, c)).astype(np.float32)
    rot, tr, hand = rng.random((n, 3), np.float32), rng.random((n, 3), np.float32), rng.random((n, 63), np.float32)
    d = pp.filter_detections(boxes, cls, rot, tr, hand)
    k = int(d["count"])
    assert k == 100 and (np.diff(d["scores"][:k]) <= 0).all()
    assert d["labels"].dtype == np.int32 and set(np.unique(d["labels"])) <= {0, 1, 2}
    assert np.array_equal(d["boxes"][:k], boxes[d["anchor_idx"][:k]])
    assert np.array_equal(d["scores"][:k], cls[d["anchor_idx"][:k], d["labels"][:k]])
    few = pp.filter_detections(boxes[:20], cls[:20] * 0.6, rot[:20], tr[:20], hand[:20])
    k = int(few["count"])
    assert k < 20 and (few["boxes"][k:] == -1).all() and (few["labels"][k:] == -1).all() and (few["hand"][k:] == -1).all()
    none = pp.filter_detections(boxes, cls * 0.4, rot, tr, hand)
    assert int(none["count"]) == 0 and (none["scores"] == -1).all()


def test_csharp_best_is_argmax_with_lowest_index_tiebreak():
    rng = np.random.default_rng(2)
    a, t = pp.anchors_for_shape((256, 256))
    n = len(a)
    reg = (rng.standard_normal((n, 4)) * 0.3).astype(np.float32)
    cls = (rng.random((n, 1)) * 0.9).astype(np.float32)
    cls[4000, 0] = cls[700, 0] = 0.97
    rot = rng.standard_normal((n, 3)).astype(np.float32)
    tr = rng.standard_normal((n, 3)).astype(np.float32)
    cam = np.array([480, 480, 128, 128, 1000, 1], np.float32)
    out = pp.csharp_best(reg, cls, rot, tr, cam, 256)
    assert out[0] == np.float32(0.97)
    assert np.allclose(out[5:8], rot[700] * np.float32(np.pi))
    assert (out[1:5] == np.trunc(out[1:5])).all()
    assert (pp.csharp_best(reg, cls * 0.5, rot, tr, cam, 256) == 0).all()


def test_csharp_selection_pinned_against_opencv_nmsboxes():
    """a19 (Program.cs:786-960): the receiver thresholds at 0.5, runs CvDnn.NMSBoxes(score 0.5, nms 0.5, top_k 10) on
    int Rects whose Width/Height hold x2/y2, then scans the survivors with a strict '>' for the best score.  The oracle
    (postprocess_ref.csharp_best) replaces that by "arg-max score, ties -> lowest index".  Pinned here against OpenCV's
    own NMSBoxes (cv2.dnn, the library OpenCvSharp4 wraps): same winner on random, tie-heavy and degenerate inputs."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for trial in range(300):
        n = int(rng.integers(1, 60))
        x1 = rng.integers(0, 200, n); y1 = rng.integers(0, 200, n)
        x2 = x1 + rng.integers(0, 80, n); y2 = y1 + rng.integers(0, 80, n)       # zero-area boxes included
        rects = [[int(a), int(b), int(c), int(d)] for a, b, c, d in zip(x1, y1, x2, y2)]   # (X, Y, "Width"=x2, "Height"=y2)
        scores = (0.5 + 0.5 * rng.random(n)).astype(np.float32)
        if trial % 3 == 0:
            scores = np.round(scores, 1)                                          # many exact ties
        cand = np.nonzero(scores > np.float32(0.5))[0]
        if len(cand) == 0:
            continue
        keep = cv2.dnn.NMSBoxes([rects[i] for i in cand], [float(scores[i]) for i in cand], 0.5, 0.5, top_k=10)
        keep = [int(k) for k in np.asarray(keep).reshape(-1)]
        assert len(keep) >= 1
        best, best_score = -1, 0.0                                               # Program.cs:934-951: strict '>' scan
        for k in keep:
            if scores[cand[k]] > best_score:
                best, best_score = int(cand[k]), float(scores[cand[k]])
        want = int(cand[np.lexsort((cand, -scores[cand].astype(np.float64)))[0]])
        assert scores[best] == scores[want]
        assert best == want, (trial, best, want)
