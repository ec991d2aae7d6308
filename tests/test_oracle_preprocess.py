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
    assert np.array_equal(ref, got)                         # down- and up-scaling: bit-exact


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


def _i420(h, w, seed):
    rng = np.random.default_rng(seed)
    rgb = _frame(h, w, seed)
    yuv = cv2.cvtColor(rgb, cv2.COLOR_RGB2YUV_I420).reshape(-1).copy()
    noise = rng.integers(-12, 13, yuv.shape)                      # leave the gamut a little: exercises the saturation
    return (yuv.astype(np.int32) + noise).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w", [(504, 896), (480, 640), (720, 1280), (360, 640)])
def test_i420_yv12_conversion_matches_opencv(h, w):
    """Program.cs:146-160: the I420 buffer converted with YUV2BGR_YV12 (chroma planes read swapped)."""
    buf = _i420(h, w, seed=h)
    ref = cv2.cvtColor(buf.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_YV12)
    assert np.array_equal(pr.i420_to_bgr_yv12(buf, h, w), ref)
    rnd = np.random.default_rng(1).integers(0, 256, h * w * 3 // 2, dtype=np.uint8)      # every byte value, no structure
    assert np.array_equal(pr.i420_to_bgr_yv12(rnd, h, w), cv2.cvtColor(rnd.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_YV12))


def _cv2_receiver_pipeline(buf, h, w, size, crop=256, mid=512):
    """The same OpenCV calls the C# receiver makes (Program.cs:146-200, 381-445), through cv2's Python binding."""
    bgr = cv2.cvtColor(buf.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_YV12)
    off_w, off_h = (bgr.shape[1] - crop) // 2, (bgr.shape[0] - crop) // 2
    dst = cv2.resize(bgr[off_h:off_h + crop, off_w:off_w + crop], (mid, mid))
    ih, iw = dst.shape[:2]
    if ih > iw:
        scale = np.float32(size) / np.float32(ih); rh, rw = size, int(np.float32(iw) * scale)
    else:
        scale = np.float32(size) / np.float32(iw); rh, rw = int(np.float32(ih) * scale), size
    dst = cv2.resize(dst, (rw, rh)).astype(np.float32)                                   # ConvertTo(CV_32F)
    dst = cv2.divide(dst, (255.0, 255.0, 255.0, 0.0))
    dst = cv2.subtract(dst, (float(np.float32(0.485)), float(np.float32(0.456)), float(np.float32(0.406)), 0.0))
    dst = cv2.divide(dst, (float(np.float32(0.229)), float(np.float32(0.224)), float(np.float32(0.225)), 0.0))
    dst = cv2.copyMakeBorder(dst, 0, size - rh, 0, size - rw, cv2.BORDER_CONSTANT, value=(0, 0, 0, 0))
    return dst, float(scale)


@pytest.mark.parametrize("h,w,size", [(504, 896, 256), (504, 896, 512), (480, 640, 256), (720, 1280, 256)])
def test_receiver_frame_path_matches_opencv(h, w, size):
    """I420 -> YV12-trick BGR -> centre crop 256 -> 512 x 512 -> ResizeAndNormalizeMat: bit-exact against the OpenCV
    calls of Program.cs, including the float32 scalar arithmetic of Cv2.Divide / Cv2.Subtract."""
    buf = _i420(h, w, seed=7 + size)
    ref, ref_scale = _cv2_receiver_pipeline(buf, h, w, size)
    got, scale = pr.csharp_frame_to_tensor(buf, h, w, size)
    assert got.shape == ref.shape == (size, size, 3) and scale == ref_scale
    assert np.array_equal(got, ref)
