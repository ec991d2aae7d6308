"""8f-1 oracle pins: oracle/preprocess_ref.py against OpenCV (cv2.resize) and against the reference method
(generators/colibri_common.py:622-656, executed from its source text in the build container)."""
import os

import numpy as np
import pytest

from oracle import preprocess_ref as pr

cv2 = pytest.importorskip("cv2")
SHAPES = [(480, 640, 256), (504, 896, 256), (720, 1280, 512), (300, 200, 256), (256, 256, 256), (1080, 1920, 512),
          (360, 640, 256), (128, 128, 256)]


def _frame(h, w, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)      # blocky structure + noise
    img = np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w]
    return (img.astype(np.int32) + rng.integers(-20, 21, (h, w, 3))).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,size", SHAPES)
def test_resize_matches_opencv(h, w, size):
    img = _frame(h, w, seed=h + w)
    rh, rw, _ = pr.resized_shape(h, w, size)
    ref = cv2.resize(img, (rw, rh))
    got = pr.resize_linear_u8(img, rw, rh)
    diff = np.abs(ref.astype(np.int32) - got.astype(np.int32))
    if rw <= w and rh <= h:
        assert diff.max() == 0                              # down-scaling (camera frames): bit-exact
    else:
        assert diff.max() <= 1 and (diff > 0).mean() < 5e-3  # up-scaling: OpenCV's scalar tail columns round differently


@pytest.mark.refpin
def test_preprocess_matches_reference_method():
    path = "/root/reference/pytorch-sandbox/generators/colibri_common.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    src = open(path).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def preprocess_image("))
    end = next(i for i in range(start + 1, len(src)) if src[i].strip().startswith("def "))
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in src[start:end])
    ns = {"cv2": cv2, "np": np}
    exec(body, ns)

    class Self:
        pass
    for h, w, size in [(480, 640, 256), (504, 896, 256), (720, 1280, 512), (300, 200, 256)]:
        s = Self()
        s.image_size = size
        img = _frame(h, w, seed=3)
        ref, ref_scale = ns["preprocess_image"](s, img.copy())
        got, scale = pr.preprocess_image(img, size)
        assert ref.dtype == np.float32 and ref.shape == got.shape == (size, size, 3)
        assert scale == ref_scale and np.array_equal(ref, got)          # bit-exact incl. the float64 mean/std steps
