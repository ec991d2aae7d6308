"""GPU parity tests (run on the B200 box): CUDA path through the C-ABI vs the oracle on the same seeded
inputs, and vs the committed golden vectors produced by the unmodified reference.

Tolerances (BASELINE.json north_star): kept-box indices and labels bit-exact; logits / box offsets within
rel 1e-3; rotation within 0.1 degree; translation within 0.1 mm.  End-to-end assertions hold in the fp32
`parity` mode; the fp16 `fast` mode is asserted per kernel and its end-to-end deltas are bounded loosely
(the random-weight network amplifies rounding ~100x)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import net_ref, postprocess_ref as pp, synth_weights as sw

pytestmark = pytest.mark.gpu

CAM = np.array([[480, 480, 128, 128, 1000, 1], [687.7084, 688.8967, 435.8758, 242.4822, 1000, 1]], np.float32)
KEYS = ("boxes", "scores", "labels", "rotation", "translation", "hand", "anchor_idx")


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.fixture(scope="module")
def frames():
    x0 = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "input_256.npy")))
    return torch.cat([x0, torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(1234))], 0)


@pytest.fixture(scope="module")
def oracle_out(synth_sd, frames):
    _, reg, cls, rot, tr, hand = net_ref.forward(synth_sd, frames)
    return [t.numpy() for t in (reg, cls, rot, tr, hand)]


@pytest.fixture(scope="module")
def parity_sess(synth_sd):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision="parity")
    yield s
    s.close()


@pytest.fixture(scope="module")
def fast_sess(synth_sd):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision="fast")
    yield s
    s.close()


def cam_rows(b):
    return np.stack([CAM[i % 2] for i in range(b)])


# ------------------------------------------------------------------------------------------------
# pointwise GEMM kernels (per-kernel parity for the tcgen05 path)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,gate,res,act", [
    (128, 64, 64, False, False, 0), (300, 144, 24, False, False, 1), (4096, 1152, 192, False, False, 1),
    (640, 320, 1152, True, True, 0), (4, 64, 64, False, False, 0), (1000, 40, 240, True, False, 0),
    (16384, 96, 16, False, False, 1), (777, 112, 672, True, True, 0), (1364, 64, 320, False, False, 2)])
def test_pointwise_gemm_kernels(M, N, K, gate, res, act):
    from hmd_ego_pose_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = (rng.standard_normal(N) * 0.1).astype(np.float32)
    rpi = 64 if gate else M
    g = (rng.random(((M + rpi - 1) // rpi, K)) + 0.25).astype(np.float32) if gate else None
    R = rng.standard_normal((M, N)).astype(np.float32) if res else None

    def ref(fp16, round_gated):
        a, w = (A.astype(np.float16), W.astype(np.float16)) if fp16 else (A, W)
        a, w = a.astype(np.float64), w.astype(np.float64)
        if gate:
            a = a * g[np.arange(M) // rpi]
            if round_gated:
                a = a.astype(np.float16).astype(np.float64)
        y = a @ w.T + bias
        y = y / (1 + np.exp(-y)) if act == 1 else (1 / (1 + np.exp(-y)) if act == 2 else y)
        if res:
            y = y + (R.astype(np.float16).astype(np.float64) if fp16 else R)
        return y

    # impl 0 = FFMA cross-check kernel, 1 = tcgen05 kind::f16 (fast mode), 3 = tcgen05 3xTF32 split precision (parity mode)
    for impl, prec, tol in ((0, 0, 2e-5), (3, 0, 2e-5), (0, 1, 2e-3), (1, 1, 2e-3)):
        D = np.zeros((M, N), np.float32)
        ms = ctypes.c_float()
        rc = lib.hmdpose_test_gemm(0, impl, prec, M, N, K, A.ctypes.data, W.ctypes.data, bias.ctypes.data,
                                   g.ctypes.data if gate else None, rpi, R.ctypes.data if res else None, act,
                                   D.ctypes.data, ctypes.byref(ms))
        assert rc == 0, lib.hmdpose_last_error(None)
        err = relerr(D, ref(prec == 1, impl == 1))
        print(f"gemm M={M} N={N} K={K} impl={impl} prec={prec}: relerr {err:.2e} ({ms.value * 1e3:.1f} us)")
        assert err < tol, (impl, prec)


# ------------------------------------------------------------------------------------------------
# network
# ------------------------------------------------------------------------------------------------
def test_network_parity_mode_vs_oracle(parity_sess, frames, oracle_out):
    got = parity_sess.raw_host(frames.numpy())
    for name, g, r in zip(("regression", "classification", "rotation", "translation_raw", "hand"), got, oracle_out):
        assert g.shape == r.shape
        assert relerr(g, r) < 1e-3, name          # logits / box offsets within rel 1e-3 (north_star)
    assert np.abs(got[2] - oracle_out[2]).max() * 180.0 < 0.1   # rotation, degrees (values are in units of pi)


def test_network_parity_mode_vs_reference_golden(parity_sess, frames, gold_dir):
    g = np.load(os.path.join(gold_dir, "net_golden_256.npz"))   # outputs of the UNMODIFIED reference module
    got = parity_sess.raw_host(frames[:2].numpy())
    hs = int(g["hand_stride"])
    for name, arr in (("regression", got[0]), ("classification", got[1]), ("rotation", got[2]),
                      ("translation_raw", got[3]), ("hand_sub", got[4][:, ::hs])):
        assert relerr(arr, g[name]) < 1e-3, name


def test_parity_mode_runs_every_pointwise_conv_on_the_tensor_cores(parity_sess, frames):
    """the mode that carries the end-to-end assertions is a tcgen05 mode -- split-precision (3xTF32)
    GEMMs on fp32 activations -- not the FFMA cross-check kernel."""
    parity_sess.detect_host(frames.numpy(), cam_rows(4))
    kernels = [k for _, k, *_ in parity_sess.profile_steps(4, mode=1, reps=1)]
    assert "gemm_simt_kernel" not in kernels, sorted(set(kernels))
    assert kernels.count("gemm_tf32_kernel") >= 32 + 3, sorted(set(kernels))


def test_network_device_api_accepts_permuted_nhwc_view(parity_sess, frames, oracle_out):
    # the reference call site passes a permuted NHWC tensor (eval/common.py:397): consumed without a copy
    x = frames.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    assert not x.is_contiguous()
    got = parity_sess.forward_raw(x)
    torch.cuda.synchronize()
    assert relerr(got[0].cpu().numpy(), oracle_out[0]) < 1e-3
    assert relerr(got[4].cpu().numpy(), oracle_out[4]) < 1e-3


def test_network_fast_mode_bounded(fast_sess, frames, oracle_out):
    got = fast_sess.raw_host(frames.numpy())
    errs = {n: relerr(g, r) for n, g, r in zip(("reg", "cls", "rot", "traw", "hand"), got, oracle_out)}
    print("fast-mode end-to-end deltas vs fp32 oracle:", errs)
    # measured 1.3e-2 .. 3.3e-2 on this random-weight net (fp16 storage through ~80 layers, each kernel within 2e-3 of
    # the reference layer on identical inputs: tests/test_gpu_stages.py); bound = measured + margin
    assert max(errs.values()) < 5e-2
    flips = int(((got[1] > 0.5) != (oracle_out[1] > 0.5)).sum())
    nthr = int((oracle_out[1] > 0.5).sum())
    print("threshold-set flips:", flips, "of", nthr)
    assert flips < 0.08 * max(1, nthr)


def test_micro_batching_is_transparent(synth_sd, frames, parity_sess):
    from hmd_ego_pose_b200 import HmdPoseSession
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision="parity", micro_batch=3)
    a = s.raw_host(frames.numpy())
    b = parity_sess.raw_host(frames.numpy())
    for x, y in zip(a, b):
        # only the grouping of the squeeze partial sums depends on the launch geometry
        assert relerr(x, y) < 2e-5
    c = s.raw_host(frames.numpy())
    for x, y in zip(a, c):
        assert np.array_equal(x, y)      # run-to-run: deterministic kernels, same bits
    s.close()


def test_512_parity(synth_sd):
    from hmd_ego_pose_b200 import HmdPoseSession
    x = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(99))
    ref = net_ref.forward(synth_sd, x)[1:]
    s = HmdPoseSession(synth_sd, image_size=512, max_batch=1, precision="parity")
    assert s.num_anchors == 49104
    got = s.raw_host(x.numpy())
    for g, r in zip(got, ref):
        assert relerr(g, r.numpy()) < 1e-3
    s.close()


def test_non_power_of_two_size_parity_mode_and_fast_mode_rejection(synth_sd):
    """image_size 384 (any multiple of 128, e.g. the reference's phi1 size 640): parity mode runs it; the fast mode's
    fused kernels need power-of-two maps and say so at create time."""
    from hmd_ego_pose_b200 import HmdPoseSession
    from hmd_ego_pose_b200._native import HmdPoseError
    x = torch.randn(1, 3, 384, 384, generator=torch.Generator().manual_seed(384))
    ref = net_ref.forward(synth_sd, x)[1:]
    s = HmdPoseSession(synth_sd, image_size=384, max_batch=1, precision="parity")
    assert s.num_anchors == 9 * (48 * 48 + 24 * 24 + 12 * 12 + 6 * 6 + 3 * 3)
    for g, r in zip(s.raw_host(x.numpy()), ref):
        assert relerr(g, r.numpy()) < 1e-3
    s.close()
    with pytest.raises(HmdPoseError, match="power-of-two"):
        HmdPoseSession(synth_sd, image_size=384, max_batch=1, precision="fast")


def test_multiclass_heads(synth_sd):
    from hmd_ego_pose_b200 import HmdPoseSession
    sd = sw.synthetic_weights(1, 256, num_classes=3)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    _, reg, cls, rot, tr, hand = [t.numpy() if hasattr(t, "numpy") else t for t in net_ref.forward(sd, x, 3)]
    s = HmdPoseSession(sd, image_size=256, max_batch=2, precision="parity")
    assert s.num_classes == 3
    got = s.raw_host(x.numpy())
    assert got[1].shape == (2, 12276, 3) and relerr(got[1], cls) < 1e-3
    cam = cam_rows(2)
    ref = pp.detect(reg, cls, rot, tr, hand, cam, 256)
    det = s.postprocess_host(reg, cls, rot, tr, hand, cam)
    for b in range(2):
        assert np.array_equal(det["anchor_idx"][b], ref[b]["anchor_idx"])
        assert np.array_equal(det["labels"][b], ref[b]["labels"])
    s.close()


# ------------------------------------------------------------------------------------------------
# post-processing: bit-exact on identical inputs
# ------------------------------------------------------------------------------------------------
def test_postprocess_on_oracle_heads(parity_sess, oracle_out):
    reg, cls, rot, tr, hand = oracle_out
    cam = cam_rows(4)
    ref = pp.detect(reg, cls, rot, tr, hand, cam, 256)
    got = parity_sess.postprocess_host(reg, cls, rot, tr, hand, cam)
    assert int(ref[0]["count"]) == 0 and (got["boxes"][0] == -1).all() and (got["labels"][0] == -1).all()
    for b in range(4):
        r = ref[b]
        for k in ("anchor_idx", "labels", "scores", "rotation", "hand"):
            assert np.array_equal(got[k][b], r[k]), (b, k)                 # bit-exact
        assert np.abs(got["boxes"][b] - r["boxes"]).max() <= 1e-4          # exp() is the only libm-dependent step
        assert np.allclose(got["translation"][b], r["translation"], rtol=1e-6, atol=1e-5)


def test_filter_boxes_bit_exact_on_identical_inputs(parity_sess, oracle_out):
    reg, cls, rot, tr, hand = oracle_out
    a, t = pp.anchors_for_shape((256, 256))
    boxes = pp.decode_boxes(a, reg, 256, 256)
    trans = pp.decode_translation(t, tr, cam_rows(4))
    got = parity_sess.filter_boxes_host(boxes, cls, rot, trans, hand)
    for b in range(4):
        r = pp.filter_detections(boxes[b], cls[b], rot[b], trans[b], hand[b])
        for k in KEYS:
            assert np.array_equal(got[k][b], r[k]), (b, k)


def test_nms_stress_ties_and_many_candidates(parity_sess):
    # > 4096 candidates (global-memory sort path), exact score ties, zero-area boxes, the 100-detection cap
    rng = np.random.default_rng(7)
    N = 12276
    c = rng.random((1, N, 2)) * 256
    wh = rng.random((1, N, 2)) * 40 + 1
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 2).clip(0, 255).astype(np.float32)
    cls = (np.round(rng.random((1, N, 1)) * 50) / 50 * 0.6 + 0.38).astype(np.float32)   # heavy ties, ~80% pass
    boxes[0, 100:200, 2] = boxes[0, 100:200, 0]
    rot, tr = rng.random((1, N, 3), np.float32), rng.random((1, N, 3), np.float32)
    hand = rng.random((1, N, 63), np.float32)
    assert (cls > 0.5).sum() > 4096
    got = parity_sess.filter_boxes_host(boxes, cls, rot, tr, hand)
    r = pp.filter_detections(boxes[0], cls[0], rot[0], tr[0], hand[0])
    for k in KEYS:
        assert np.array_equal(got[k][0], r[k]), k


def test_csharp_best_pose(parity_sess, oracle_out, frames):
    reg, cls, rot, tr, hand = oracle_out
    for b in range(4):
        ref = pp.csharp_best(reg[b], cls[b], rot[b], tr[b], CAM[0], 256)
        got = parity_sess.best_from_raw_host(reg[b], cls[b], rot[b], tr[b], CAM[0])
        assert np.array_equal(got[:5], ref[:5]) and np.allclose(got[5:], ref[5:], rtol=1e-6, atol=1e-7), b
    e2e = parity_sess.best_host(frames[1].numpy(), CAM[0])
    ref = pp.csharp_best(reg[1], cls[1], rot[1], tr[1], CAM[0], 256)
    assert e2e[0] > 0.5 and abs(e2e[0] - ref[0]) < 1e-3
    assert np.abs(e2e[5:8] - ref[5:8]).max() * 180 / np.pi < 0.1 and np.abs(e2e[8:] - ref[8:]).max() < 1e-4
    empty = parity_sess.best_host(frames[0].numpy(), CAM[0])     # onnx-models/input.npy: nothing passes
    assert (empty == 0).all()


# ------------------------------------------------------------------------------------------------
# end to end: TrainModelWithLoss boundary
# ------------------------------------------------------------------------------------------------
def test_end_to_end_detect_parity_mode(parity_sess, frames, oracle_out):
    reg, cls, rot, tr, hand = oracle_out
    cam = cam_rows(4)
    ref = pp.detect(reg, cls, rot, tr, hand, cam, 256)
    det = parity_sess.detect_host(frames.numpy(), cam)
    for b in range(4):
        r, k = ref[b], int(ref[b]["count"])
        if not np.array_equal(det["anchor_idx"][b], r["anchor_idx"]):
            s = np.sort(cls[b, :, 0])
            pytest.fail(f"image {b}: kept indices differ; min |score-0.5| = {np.abs(s - 0.5).min():.3e}")
        assert np.array_equal(det["labels"][b], r["labels"])
        assert np.abs(det["rotation"][b][:k] - r["rotation"][:k]).max() * 180.0 < 0.1 if k else True   # degrees
        assert np.abs(det["translation"][b][:k] - r["translation"][:k]).max() < 0.1 if k else True     # mm
        assert np.abs(det["boxes"][b][:k] - r["boxes"][:k]).max() < 0.05 if k else True                # px
        assert (det["boxes"][b][k:] == -1).all() and (det["hand"][b][k:] == -1).all()


def test_b16_parity_mode_end_to_end_bit_exact_indices(synth_sd):
    """BASELINE.json configs[1]: batch 16 (the benched launch plan: tile counts, chain CTAs per image, persistent grids)
    forward + NMS + pose recovery, kept indices / labels bit-exact against the fp32 oracle, rotation within 0.1 degree,
    translation within 0.1 mm -- through the host API and through the device API used by bench.py."""
    from hmd_ego_pose_b200 import HmdPoseSession
    x = torch.randn(16, 3, 256, 256, generator=torch.Generator().manual_seed(1616))
    cam = cam_rows(16)
    reg, cls, rot, tr, hand = [t.numpy() for t in net_ref.forward(synth_sd, x)[1:]]
    ref = pp.detect(reg, cls, rot, tr, hand, cam, 256)
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=16, precision="parity")
    raw = s.raw_host(x.numpy())
    for name, g, r in zip(("regression", "classification", "rotation", "translation_raw", "hand"), raw, (reg, cls, rot, tr, hand)):
        assert relerr(g, r) < 1e-3, name
    det_host = s.detect_host(x.numpy(), cam)
    dev = s.detect(x.cuda(), torch.from_numpy(cam).cuda())
    torch.cuda.synchronize()
    det_dev = dict(zip(KEYS, [t.cpu().numpy() for t in dev]))
    total = 0
    for det in (det_host, det_dev):
        for b in range(16):
            r, k = ref[b], int(ref[b]["count"])
            total += k
            assert np.array_equal(det["anchor_idx"][b], r["anchor_idx"]), b
            assert np.array_equal(det["labels"][b], r["labels"]), b
            if k:
                assert np.abs(det["rotation"][b][:k] - r["rotation"][:k]).max() * 180.0 < 0.1      # degrees
                assert np.abs(det["translation"][b][:k] - r["translation"][:k]).max() < 0.1        # mm
                assert relerr(det["boxes"][b][:k], r["boxes"][:k]) < 1e-3
    assert total > 16 * 20      # the synthetic net fires on every frame: the comparison is not vacuous
    s.close()


def test_b16_fast_mode_detections_match_oracle_postprocessing_of_own_heads(synth_sd):
    """The benched configuration itself (fast mode, batch 16): the kept indices / labels / rotation are bit-exact
    against the oracle post-processing applied to the GPU's own head tensors (the network half of the fast mode is
    pinned kernel by kernel at this batch size in tests/test_gpu_stages.py)."""
    from hmd_ego_pose_b200 import HmdPoseSession
    x = torch.randn(16, 3, 256, 256, generator=torch.Generator().manual_seed(1617)).numpy()
    cam = cam_rows(16)
    s = HmdPoseSession(synth_sd, image_size=256, max_batch=16, precision="fast")
    raw = s.raw_host(x)
    det = s.detect_host(x, cam)
    ref = pp.detect(*raw, cam, 256)
    for b in range(16):
        assert np.array_equal(det["anchor_idx"][b], ref[b]["anchor_idx"])
        assert np.array_equal(det["labels"][b], ref[b]["labels"])
        assert np.array_equal(det["scores"][b], ref[b]["scores"])
        k = int(ref[b]["count"])
        assert k > 0
        # pose rows come from pose_gather_kernel (headers evaluated at the kept anchors only, fp32 CUDA cores) while
        # `raw` holds the dense tcgen05 headers: same values up to the fp16 rounding of the header input
        assert np.abs(det["rotation"][b][:k] - ref[b]["rotation"][:k]).max() < 2e-3 * np.abs(ref[b]["rotation"][:k]).max()
        assert np.abs(det["translation"][b][:k] - ref[b]["translation"][:k]).max() < 2e-3 * np.abs(ref[b]["translation"][:k]).max()
        assert np.abs(det["hand"][b][:k] - ref[b]["hand"][:k]).max() < 2e-3 * np.abs(ref[b]["hand"][:k]).max()
        assert (det["rotation"][b][k:] == -1).all() and (det["translation"][b][k:] == -1).all()
    s.close()


def test_pose_headers_gathered_at_kept_anchors_match_the_dense_headers(synth_sd, frames):
    """on the detection path the rotation / translation / hand headers are evaluated only at the <= 100
    kept anchors (pose_gather_kernel) -- legal because filter_detections discards every other row
    (hmdegopose/layers.py:369-374).  Same detections as the dense headers (HMDPOSE_DENSE_POSE=1), fewer header launches'
    work: the dense rotation / translation header tiles are gone from the plan."""
    from hmd_ego_pose_b200 import HmdPoseSession
    cam = cam_rows(4)
    for precision, tol in (("parity", 2e-5), ("fast", 2e-3)):
        os.environ["HMDPOSE_DENSE_POSE"] = "1"
        try:
            dense = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision=precision)
        finally:
            del os.environ["HMDPOSE_DENSE_POSE"]
        gath = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision=precision)
        d0, d1 = dense.detect_host(frames.numpy(), cam), gath.detect_host(frames.numpy(), cam)
        for key in ("anchor_idx", "labels", "scores", "boxes"):
            assert np.array_equal(d0[key], d1[key]), key
        for key in ("rotation", "translation", "hand"):
            m = d0["anchor_idx"] >= 0
            assert np.abs(d0[key][m] - d1[key][m]).max() <= tol * np.abs(d0[key][m]).max(), (precision, key)
            assert np.array_equal(d0[key][~m], d1[key][~m])            # -1 padding
        k0 = [k for _, k, *_ in dense.profile_steps(4, mode=1, reps=1)]
        k1 = [k for _, k, *_ in gath.profile_steps(4, mode=1, reps=1)]
        assert "pose_gather_kernel" in k1 and "pose_gather_kernel" not in k0 and "hand_gather_kernel" not in k1
        b0 = sum(by for *_, by, _ in dense.profile_steps(4, mode=1, reps=1))
        b1 = sum(by for *_, by, _ in gath.profile_steps(4, mode=1, reps=1))
        assert b1 < b0
        dense.close()
        gath.close()


def test_train_model_with_loss_dropin(synth_sd, frames, oracle_out):
    from hmd_ego_pose_b200 import TrainModelWithLoss
    reg, cls, rot, tr, hand = oracle_out
    cam = cam_rows(4)
    ref = pp.detect(reg, cls, rot, tr, hand, cam, 256)
    m = TrainModelWithLoss(synth_sd, max_batch=4, precision="parity", compat_last_only=True, compat_cpu=True).eval()
    out = m(frames.cuda(), torch.from_numpy(cam), params={"img_size": (256, 256)})
    assert len(out) == 6 and out[0].shape == (100, 4) and out[2].dtype == torch.int32 and out[5].shape == (100, 63)
    assert not out[0].is_cuda
    k = int(ref[3]["count"])
    assert np.array_equal(out[2].numpy(), ref[3]["labels"])
    assert np.allclose(out[1].numpy()[:k], ref[3]["scores"][:k], atol=1e-4)
    with pytest.raises(NotImplementedError):
        m(frames.cuda(), torch.from_numpy(cam), is_losses=True, params={"img_size": (256, 256)})
    batched = TrainModelWithLoss(synth_sd, max_batch=4, precision="parity")(frames.cuda(), torch.from_numpy(cam))
    assert batched[0].shape == (4, 100, 4) and batched[0].is_cuda


def test_argument_errors(parity_sess, frames):
    from hmd_ego_pose_b200 import _native
    big = np.zeros((5, 3, 256, 256), np.float32)
    with pytest.raises(_native.HmdPoseError, match="batch out of range"):
        parity_sess.detect_host(big, cam_rows(5))
    with pytest.raises(ValueError):
        parity_sess.forward_raw(torch.zeros(1, 3, 128, 128))


def test_profile_steps_reports_every_launch(fast_sess, frames):
    fast_sess.detect_host(frames.numpy(), cam_rows(4))
    prof = fast_sess.profile_steps(4, mode=1, reps=2)
    kernels = {k for _, k, *_ in prof}
    for want in ("stem_kernel", "dw3_kernel", "gemm_tc2_kernel", "se3_kernel", "sepconv_kernel", "sepconv3_kernel",
                 "filter_fused_kernel"):
        assert want in kernels, (want, kernels)
    assert all(ms > 0 for _, _, ms, _, _ in prof) and len(prof) == fast_sess.last_launch_count


# ------------------------------------------------------------------------------------------------
# launch-plan variants of the fast mode must agree with each other
# ------------------------------------------------------------------------------------------------
def _fast_raw(synth_sd, frames, env, precision="fast"):
    from hmd_ego_pose_b200 import HmdPoseSession
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:   # the switches are read while the launch plan is built (first call)
        s = HmdPoseSession(synth_sd, image_size=256, max_batch=4, precision=precision)
        out = s.raw_host(frames.numpy())
        n = s.last_launch_count
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out, n


def test_bifpn_chain_launch_is_bit_identical_to_one_launch_per_node(synth_sd, frames):
    chained, n_chain = _fast_raw(synth_sd, frames, {})
    single, n_single = _fast_raw(synth_sd, frames, {"HMDPOSE_NO_CHAIN": "1"})
    assert n_single - n_chain == 11                      # 24 node launches -> 13
    for a, b in zip(chained, single):
        assert np.array_equal(a, b)                      # same per-pixel arithmetic, only the launch structure differs


def test_implicit_gemm_heads_agree_with_stencil_path(synth_sd, frames, oracle_out):
    """sepconv3 (nine shifted-descriptor tcgen05 taps, folded tap matrices) vs sepconv (CUDA-core stencil + one GEMM):
    different rounding points of the same fp16 pipeline, so they agree to the fast-mode error level."""
    implicit, _ = _fast_raw(synth_sd, frames, {})
    stencil, _ = _fast_raw(synth_sd, frames, {"HMDPOSE_NO_SEP3": "1"})
    for a, b, ref in zip(implicit, stencil, oracle_out):
        assert relerr(a, b) < 3e-2
        assert relerr(a, ref) < 0.1 and relerr(b, ref) < 0.1


def test_pose_packet_end_to_end(parity_sess, frames):
    """hmdpose_run_packet = hmdpose_run_best + the 24-byte data-channel message (Program.cs:208-292)."""
    cam = CAM[0]
    for i in (1, 2):
        best = parity_sess.best_host(frames[i].numpy(), cam)
        packet, score = parity_sess.packet_host(frames[i].numpy(), cam)
        assert packet == pp.csharp_pose_packet(best) and score == best[0]


def test_squeeze_excite_folded_into_depthwise_tail_matches_the_separate_launch(synth_sd, frames):
    """HMDPOSE_SE_FOLD=1 (optional): last-block-per-image gate inside dw3_kernel vs se3_kernel, fp32 parity mode: only the
    summation order of the squeeze / FC reductions differs."""
    base, n0 = _fast_raw(synth_sd, frames, {}, precision="parity")
    fold, n1 = _fast_raw(synth_sd, frames, {"HMDPOSE_SE_FOLD": "1"}, precision="parity")
    assert n0 - n1 == 6                                   # blocks 0..5 lose their SE launch (the fold covers the tiny FC layers)
    for a, b in zip(fold, base):
        assert relerr(a, b) < 2e-5


def test_fast_mode_512_detect_and_ungraphed_launches(synth_sd):
    """BASELINE.json configs[3] shape (512x512) in the bench's fast mode, staged check: the detections equal the oracle
    post-processing applied to the GPU's own head tensors; the same with CUDA graphs disabled (plain launches)."""
    from hmd_ego_pose_b200 import HmdPoseSession
    x = torch.randn(3, 3, 512, 512, generator=torch.Generator().manual_seed(21)).numpy()
    cam = np.tile(np.array([[960, 960, 256, 256, 1000, 1]], np.float32), (3, 1))
    outs = []
    for use_graph in (True, False):
        s = HmdPoseSession(synth_sd, image_size=512, max_batch=3, precision="fast", micro_batch=2, use_graph=use_graph)
        raw = s.raw_host(x)
        assert raw[0].shape == (3, 49104, 4)
        det = s.detect_host(x, cam)
        ref = pp.detect(*raw, cam, 512)
        for b in range(3):
            assert np.array_equal(det["anchor_idx"][b], ref[b]["anchor_idx"])
            assert np.array_equal(det["labels"][b], ref[b]["labels"])
            k = int(ref[b]["count"])
            assert np.abs(det["rotation"][b][:k] - ref[b]["rotation"][:k]).max() < 2e-3 * np.abs(ref[b]["rotation"][:k]).max()
            assert (det["rotation"][b][k:] == -1).all()
        outs.append(raw)
        s.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)          # graph replay and plain launches run the same kernels


def test_fast_mode_rejects_128(synth_sd):
    """128 x 128 is a parity-mode size only: its coarsest pyramid level (1 x 1) does not fit the staging tile of the fused
    BiFPN / head kernel, so fast-mode handles refuse it at create time instead of faulting at the first run."""
    from hmd_ego_pose_b200 import HmdPoseSession
    from hmd_ego_pose_b200._native import HmdPoseError
    with pytest.raises(HmdPoseError, match="power-of-two"):
        HmdPoseSession(synth_sd, image_size=128, max_batch=2, precision="fast")
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(77))
    ref = net_ref.forward(synth_sd, x)[1:]
    s = HmdPoseSession(synth_sd, image_size=128, max_batch=1, precision="parity")
    for g, r in zip(s.raw_host(x.numpy()), ref):
        assert relerr(g, r.numpy()) < 1e-3
    s.close()
