"""8f-1 through the C ABI: uint8 frames -> hmdpose_preprocess / hmdpose_run_detect_u8 / hmdpose_run_best_u8 against
oracle/preprocess_ref.py (bit-exact vs cv2.resize and vs the reference method, tests/test_oracle_preprocess.py)."""
import numpy as np
import pytest

from oracle import preprocess_ref as pr

pytestmark = pytest.mark.gpu
CAM = np.array([[480, 480, 128, 128, 1000, 1]], np.float32)


def _frames(b, h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (b, h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((1, 8, 8, 1), np.uint8))[:, :h, :w]
    return (img.astype(np.int32) + rng.integers(-20, 21, (b, h, w, 3))).clip(0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def sess(synth_sd):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=3, precision="parity")
    yield s
    s.close()


@pytest.mark.parametrize("h,w", [(480, 640), (504, 896), (300, 200), (256, 256), (720, 1280)])
def test_preprocess_is_bit_exact(sess, h, w):
    f = _frames(2, h, w, seed=h)
    got, scale = sess.preprocess_host(f)
    for b in range(2):
        ref, ref_scale = pr.preprocess_image(f[b], 256)
        assert np.array_equal(got[b], ref)                       # resize, /255, mean/std, zero padding: every bit
        assert scale == pytest.approx(ref_scale, rel=1e-7)


def test_detect_from_uint8_frames_equals_detect_on_the_preprocessed_tensor(sess):
    f = _frames(3, 360, 640, seed=9)
    cam = np.repeat(CAM, 3, axis=0)
    x = np.stack([pr.preprocess_image(f[b], 256)[0] for b in range(3)]).transpose(0, 3, 1, 2)   # NCHW like common.py:397
    want = sess.detect_host(np.ascontiguousarray(x), cam)
    got = sess.detect_u8_host(f, cam)
    for k in ("boxes", "scores", "labels", "rotation", "translation", "hand", "anchor_idx"):
        assert np.array_equal(got[k], want[k]), k
    assert got["scale"] == pytest.approx(256 / 640)
    best, _ = sess.best_u8_host(f[1], CAM[0])
    assert np.array_equal(best, sess.best_host(np.ascontiguousarray(x[1]), CAM[0]))


def _i420_frames(n, h, w, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, h * w * 3 // 2), dtype=np.uint8)        # every byte value: saturation paths included


@pytest.mark.parametrize("h,w", [(128, 128), (200, 120), (96, 250), (504, 896)])
def test_uint8_preprocess_up_scaling_border_rows(sess, h, w):
    """frames smaller than the network input are UP-scaled: OpenCV's clamped-row rule of the vertical pass (border rows
    read the same source row twice with separately truncated products) -- bit-exact too"""
    f = _frames(1, h, w, seed=w)
    got, _ = sess.preprocess_host(f)
    assert np.array_equal(got[0], pr.preprocess_image(f[0], 256)[0])


@pytest.mark.parametrize("h,w", [(504, 896), (480, 640), (720, 1280)])
def test_i420_receiver_frame_path_is_bit_exact(sess, h, w):
    """WebRTCNetCoreSandbox/Program.cs:137-200 in one kernel: I420 -> YUV2BGR_YV12 -> centre crop 256 -> 512 x 512 ->
    ResizeAndNormalizeMat(256): every bit equal to the oracle, which is itself bit-exact against the same OpenCV calls
    (tests/test_oracle_preprocess.py)."""
    f = _i420_frames(2, h, w, seed=h + 1)
    got, scale = sess.preprocess_i420_host(f, h, w)
    for b in range(2):
        ref, ref_scale = pr.csharp_frame_to_tensor(f[b], h, w, 256)
        assert np.array_equal(got[b], ref)
        assert scale == ref_scale == 0.5
    # other crop / rescale sizes (no rescale at all: crop == rescaled == network size)
    got2, scale2 = sess.preprocess_i420_host(f[:1], h, w, crop_size=256, rescaled_size=256)
    ref2, _ = pr.csharp_frame_to_tensor(f[0], h, w, 256, crop=256, mid=256)
    assert np.array_equal(got2[0], ref2) and scale2 == 1.0


def test_best_pose_from_an_i420_frame_equals_best_pose_on_the_preprocessed_tensor(sess):
    h, w = 504, 896
    f = _i420_frames(1, h, w, seed=77)[0]
    x = pr.csharp_frame_to_tensor(f, h, w, 256)[0].transpose(2, 0, 1)          # CvDnn.BlobFromImage: HWC -> CHW
    best, scale = sess.best_i420_host(f, h, w, CAM[0])
    assert np.array_equal(best, sess.best_host(np.ascontiguousarray(x), CAM[0])) and scale == 0.5
