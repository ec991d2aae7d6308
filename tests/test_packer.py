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
