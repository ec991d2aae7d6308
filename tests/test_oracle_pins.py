"""Pin the oracle against the reference's own code (build container only: needs /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import net_ref, postprocess_ref as pp, ref_import, synth_weights as sw

pytestmark = [pytest.mark.refpin,
              pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")]


def test_state_dict_layout():
    m = ref_import.reference_model()
    rsd = m.state_dict()
    shapes = sw.param_shapes()
    assert set(shapes) == set(rsd) and len(rsd) == 1048
    for k, v in rsd.items():
        assert tuple(v.shape) == shapes[k], k
    assert sum(p.numel() for p in m.parameters()) == 3919055


@pytest.mark.parametrize("size,batch", [(256, 2), (512, 1)])
def test_network_matches_reference(synth_sd, size, batch):
    m = ref_import.reference_model()
    m.load_state_dict(synth_sd)
    x = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(7))
    # the reference call site passes a permuted NHWC view (eval/common.py:397)
    x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    with torch.no_grad():
        ref = m(x)
    out = net_ref.forward(synth_sd, x)
    for a, b in zip(ref[0], out[0]):
        assert torch.equal(a, b)
    for a, b in zip(ref[1:], out[1:]):
        assert a.shape == b.shape and torch.equal(a, b)


def test_network_multiclass_matches_reference():
    sd = sw.synthetic_weights(1, 256, num_classes=3)
    m = ref_import.reference_model(num_classes=3)
    m.load_state_dict(sd)
    x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = m(x)
    out = net_ref.forward(sd, x, num_classes=3)
    assert torch.equal(ref[2], out[2]) and ref[2].shape == (1, 12276, 3)


def test_network_iter1_matches_reference():
    """--iter 1 (evaluate.py:26 default): state-dict layout and all five head tensors bit-identical, including the
    reference's zip() quirk that runs only conv_list[0] of every iterative sub-net (hmdegopose/model.py:252-258)."""
    sd = sw.synthetic_weights(3, 256, iters=1)
    m = ref_import.reference_model(iters=1)
    assert set(m.state_dict()) == set(sd) == set(sw.param_shapes(1, 1))
    m.load_state_dict(sd)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = m(x)
    out = net_ref.forward(sd, x)
    assert net_ref.has_iterative(sd)
    for a, b in zip(ref[1:], out[1:]):
        assert a.shape == b.shape and torch.equal(a, b)


@pytest.mark.parametrize("size", [256, 512, 320])
def test_anchors_match_reference(size):
    ra = ref_import.reference_anchor_functions()
    a, t = pp.anchors_for_shape((size, size))
    ar, tr = ra.anchors_for_shape((size, size))
    assert np.array_equal(a, ar) and np.array_equal(t, tr)


def test_decode_matches_reference_layers():
    L = ref_import.reference_layers()
    rng = np.random.default_rng(0)
    B, S = 3, 256
    a, t = pp.anchors_for_shape((S, S))
    reg = (rng.standard_normal((B, len(a), 4)) * 0.7).astype(np.float32)
    raw = (rng.standard_normal((B, len(a), 3)) * 0.7).astype(np.float32)
    cam = np.array([[480, 480, 128, 128, 1000, 1], [687.7084, 688.8967, 435.8758, 242.4822, 1000, 1],
                    [572.4114, 573.57043, 325.2611, 242.04899, 1000, 1.6666666]], np.float32)
    img = torch.zeros(B, 3, S, S)
    with torch.no_grad():
        rb = L.ClipBoxes()(img, L.RegressBoxes()(torch.tensor(a)[None], torch.tensor(reg))).numpy()
        c = torch.tensor(cam)
        rt = L.CalculateTxTy()(L.RegressTranslation()(torch.tensor(t)[None], torch.tensor(raw)),
                               fx=c[:, 0], fy=c[:, 1], px=c[:, 2], py=c[:, 3], tz_scale=c[:, 4],
                               image_scale=c[:, 5]).numpy()
    ob = pp.decode_boxes(a, reg, S, S)
    ot = pp.decode_translation(t, raw, cam)
    # identical fp32 op order; numpy exp vs torch exp may differ by 1 ulp
    assert np.abs(ob - rb).max() <= 2e-4 * max(1.0, np.abs(rb).max() / 255)
    assert np.allclose(ot, rt, rtol=2e-7, atol=0)
