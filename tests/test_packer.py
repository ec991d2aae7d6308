"""Host logic of the product side on the CPU: BN folding / blob layout / prefix stripping."""
import struct

import numpy as np
import pytest
import torch

from hmd_ego_pose_b200 import packer
from oracle import net_ref
from tests import folded_ref


def test_fold_reproduces_oracle(synth_sd):
    folded = packer.fold(synth_sd)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = net_ref.forward(synth_sd, x)
        got = folded_ref.forward(folded, x)
    for name, a, b in zip(("reg", "cls", "rot", "trans", "hand"), ref[1:], got[1:]):
        err = float((a - b).abs().max() / a.abs().max())
        assert err < 2e-4, (name, err)
    for a, b in zip(ref[0], got[0]):
        assert float((a - b).abs().max() / a.abs().max()) < 2e-4


def test_blob_layout_and_prefix(synth_sd):
    blob = packer.pack(synth_sd)
    magic, ver, n, ncls, _ = struct.unpack_from("<8sIIII", blob, 0)
    assert magic == b"HMDPOSEW" and ver == 1 and ncls == 1
    folded = packer.fold(synth_sd)
    assert n == len(folded)
    data0 = (24 + n * 136 + 63) // 64 * 64
    name, ndim, d0, d1, d2, d3, _, off, cnt = struct.unpack_from("<96sI4IIQQ", blob, 24)
    name = name.rstrip(b"\0").decode()
    arr = np.frombuffer(blob, np.float32, cnt, data0 + off)
    assert np.array_equal(arr, folded[name].ravel()) and name == "stem.w" and (d0, d1, d2, d3) == (32, 3, 3, 3)
    # training checkpoints carry a "model." / "model.module." prefix (evaluate.py:105-116)
    pre = {"model." + k: v for k, v in synth_sd.items()}
    assert packer.pack(pre) == blob
    pre2 = {"model.module." + k: v for k, v in synth_sd.items()}
    assert packer.pack(pre2) == blob


def test_fusion_weights_are_normalised(synth_sd):
    folded = packer.fold(synth_sd)
    w = folded["bifpn1.fw.p4_w2"]
    ref = torch.relu(synth_sd["bifpn.1.p4_w2"])
    ref = ref / (ref.sum() + 1e-4)
    assert np.array_equal(w, ref.numpy())
    assert (folded["bifpn0.fw.p6_w1"][2] == 0)


def test_iter1_checkpoint_packs_the_single_executed_refinement_layer():
    from oracle import synth_weights as sw
    from hmd_ego_pose_b200 import packer
    sd = sw.synthetic_weights(3, 256, iters=1)
    t = packer.fold(sd)
    assert t["head.rot.it.dw.w"].shape == (91, 3, 3) and t["head.rot.it.pw.w"].shape == (64, 91)
    assert t["head.hand.it.dw.w"].shape == (631, 3, 3) and t["head.hand.it.hdr0.pw.w"].shape == (567, 64)
    assert t["head.trans.it.hdr0.pw.w"].shape == (18, 64) and t["head.trans.it.hdr1.pw.w"].shape == (9, 64)
    assert not any("conv_list.1" in k for k in t)
    assert len(packer.pack(sd)) > len(packer.pack({k: v for k, v in sd.items() if ".iterative_submodel." not in k}))


def test_product_side_synthetic_weights_equal_the_oracle_recipe(gold_dir):
    """bench.py's GPU arm builds its weights with hmd_ego_pose_b200.synthetic (no oracle import); the CPU baseline uses
    the same function, and it must reproduce oracle.synth_weights bit for bit (same network on both arms)."""
    import os
    import torch
    from hmd_ego_pose_b200 import synthetic
    from oracle import synth_weights as sw
    path = os.path.join(gold_dir, "bn_stats_seed0.npz")
    a = synthetic.synthetic_state_dict(0, bn_stats_path=path)
    b = sw.synthetic_weights(0, 256, bn_stats=sw.load_bn_stats(path))
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    a1, b1 = synthetic.raw_weights(3, 2, 1), sw.raw_weights(3, 2, 1)
    assert set(a1) == set(b1) and all(torch.equal(a1[k], b1[k]) for k in a1)
