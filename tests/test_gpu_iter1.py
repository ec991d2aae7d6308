"""--iter 1 checkpoints (evaluate.py:26 default): one refinement step of the rotation / translation / hand
sub-nets (hmdegopose/model.py:232-346) through the C ABI against oracle/net_ref.py (bit-identical to the reference
module built with params['iter'] = 1, tests/test_oracle_pins.py)."""
import numpy as np
import pytest
import torch

from oracle import net_ref, postprocess_ref as pp, synth_weights as sw

pytestmark = pytest.mark.gpu
CAM = np.array([[480, 480, 128, 128, 1000, 1]], np.float32)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.fixture(scope="module")
def sd1():
    return sw.synthetic_weights(3, 256, iters=1)


@pytest.fixture(scope="module")
def x3():
    return torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(77))


@pytest.fixture(scope="module")
def oracle1(sd1, x3):
    assert net_ref.has_iterative(sd1)
    return [t.numpy() for t in net_ref.forward(sd1, x3)[1:]]


def test_iter1_parity_mode_matches_oracle(sd1, x3, oracle1):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(sd1, image_size=256, max_batch=3, precision="parity")
    got = s.raw_host(x3.numpy())
    for name, g, r in zip(("regression", "classification", "rotation", "translation_raw", "hand"), got, oracle1):
        assert relerr(g, r) < 1e-3, name
    # the refinement really ran: without it the estimates differ by O(1)
    sd0 = {k: v for k, v in sd1.items() if ".iterative_submodel." not in k}
    rot0 = net_ref.forward(sd0, x3)[3].numpy()
    assert relerr(rot0, oracle1[2]) > 1e-2
    # detections through the full path: same kept anchors as the oracle post-processing on the oracle heads
    cam = np.repeat(CAM, 3, axis=0)
    ref = pp.detect(*oracle1, cam, 256)
    det = s.detect_host(x3.numpy(), cam)
    for b in range(3):
        assert np.array_equal(det["anchor_idx"][b], ref[b]["anchor_idx"])
        k = int(ref[b]["count"])
        if k:
            assert np.abs(det["rotation"][b][:k] - ref[b]["rotation"][:k]).max() < 2e-3
            assert np.abs(det["translation"][b][:k] - ref[b]["translation"][:k]).max() < 0.1      # mm
            assert relerr(det["hand"][b][:k], ref[b]["hand"][:k]) < 1e-3
    best = s.best_host(x3[1].numpy(), CAM[0])
    want = pp.csharp_best(oracle1[0][1], oracle1[1][1], oracle1[2][1], oracle1[3][1], CAM[0], 256)
    assert best[0] == pytest.approx(want[0], rel=1e-4) and np.allclose(best[5:], want[5:], rtol=2e-3, atol=2e-3)
    s.close()


def test_iter1_fast_mode(sd1, x3, oracle1):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(sd1, image_size=256, max_batch=3, precision="fast")
    got = s.raw_host(x3.numpy())
    for name, g, r in zip(("regression", "classification", "rotation", "translation_raw", "hand"), got, oracle1):
        assert relerr(g, r) < 0.1, name
    cam = np.repeat(CAM, 3, axis=0)
    det = s.detect_host(x3.numpy(), cam)                                   # staged check: oracle post on the GPU heads
    ref = pp.detect(*got, cam, 256)
    for b in range(3):
        assert np.array_equal(det["anchor_idx"][b], ref[b]["anchor_idx"])
        assert np.array_equal(det["hand"][b], ref[b]["hand"])
    s.close()
