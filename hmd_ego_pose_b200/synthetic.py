"""Seeded random-init weights of the EfficientPose-phi0 architecture for benchmarks and demos (no checkpoint is
shipped with the reference: SURVEY.md 0, `.MISSING_LARGE_BLOBS`).

Product-side utility: it only builds a ``state_dict`` with the reference's parameter names --
conv weights ~ N(0, 1/fan_in), BN gamma ~ U(0.7, 1.3), beta ~ N(0, 0.3), BiFPN fusion parameters ~ U(-0.2, 1.5),
regression-type headers x0.1, classifier header x0.5 with bias -2 -- and installs BatchNorm running statistics from a
file (``tests/golden/bn_stats_seed0.npz``: statistics calibrated once so that activations neither vanish nor explode).
It does not import ``oracle/``; ``tests/test_packer.py`` checks that it reproduces ``oracle.synth_weights`` bit for bit,
so the GPU arm and the CPU baseline of bench.py run the same network.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import packer

CLS_BIAS = -2.0

HEADS = ("regressor", "classifier", "rotation_net", "translation_net", "hand_net")


def header_specs(num_classes: int = 1):
    """(state-dict prefix, out channels, row width) of every header sepconv."""
    return [
        ("regressor.header", 36, 4),
        ("classifier.header", 9 * num_classes, num_classes),
        ("rotation_net.initial_rotation", 27, 3),
        ("translation_net.initial_translation_xy", 18, 2),
        ("translation_net.initial_translation_z", 9, 1),
        ("hand_net.initial_hand_coords", 567, 63),
    ]


def param_shapes(num_classes: int = 1, iters: int = 0) -> "OrderedDict[str, tuple]":
    """Names and shapes of the reference state_dict (appendix A.6), in module order; ``iters`` = params["iter"]."""
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def bn(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)
        s[p + ".running_mean"] = (c,)
        s[p + ".running_var"] = (c,)
        s[p + ".num_batches_tracked"] = ()

    def sep(p, cin, cout, norm):
        s[p + ".depthwise_conv.conv.weight"] = (cin, 1, 3, 3)
        s[p + ".pointwise_conv.conv.weight"] = (cout, cin, 1, 1)
        s[p + ".pointwise_conv.conv.bias"] = (cout,)
        if norm:
            bn(p + ".bn", cout)

    for c in range(3):
        p = f"bifpn.{c}"
        for n in ("conv6_up", "conv5_up", "conv4_up", "conv3_up",
                  "conv4_down", "conv5_down", "conv6_down", "conv7_down"):
            sep(f"{p}.{n}", 64, 64, True)
        if c == 0:
            for n, cin in (("p5_down_channel", 320), ("p4_down_channel", 112), ("p3_down_channel", 40),
                           ("p5_to_p6", 320), ("p4_down_channel_2", 112), ("p5_down_channel_2", 320)):
                s[f"{p}.{n}.0.conv.weight"] = (64, cin, 1, 1)
                s[f"{p}.{n}.0.conv.bias"] = (64,)
                bn(f"{p}.{n}.1", 64)
        for n, k in (("p6_w1", 2), ("p5_w1", 2), ("p4_w1", 2), ("p3_w1", 2),
                     ("p4_w2", 3), ("p5_w2", 3), ("p6_w2", 3), ("p7_w2", 2)):
            s[f"{p}.{n}"] = (k,)

    def head(p, headers):
        for i in range(3):
            sep(f"{p}.conv_list.{i}", 64, 64, False)
        for lvl in range(5):
            for i in range(3):
                bn(f"{p}.bn_list.{lvl}.{i}", 64)
        for hp, cout in headers:
            sep(hp, 64, cout, False)

    head("regressor", [("regressor.header", 36)])
    head("classifier", [("classifier.header", 9 * num_classes)])

    p = "backbone_net.model"
    s[p + "._conv_stem.conv.weight"] = (32, 3, 3, 3)
    bn(p + "._bn0", 32)
    for i, (k, st, e, cin, cout, skip) in enumerate(packer.B0_BLOCKS):
        b = f"{p}._blocks.{i}"
        cexp = cin * e
        if e != 1:
            s[b + "._expand_conv.conv.weight"] = (cexp, cin, 1, 1)
            bn(b + "._bn0", cexp)
        s[b + "._depthwise_conv.conv.weight"] = (cexp, 1, k, k)
        bn(b + "._bn1", cexp)
        cse = max(1, int(cin * 0.25))
        s[b + "._se_reduce.conv.weight"] = (cse, cexp, 1, 1)
        s[b + "._se_reduce.conv.bias"] = (cse,)
        s[b + "._se_expand.conv.weight"] = (cexp, cse, 1, 1)
        s[b + "._se_expand.conv.bias"] = (cexp,)
        s[b + "._project_conv.conv.weight"] = (cout, cexp, 1, 1)
        bn(b + "._bn2", cout)

    head("rotation_net", [("rotation_net.initial_rotation", 27)])
    head("translation_net", [("translation_net.initial_translation_xy", 18),
                             ("translation_net.initial_translation_z", 9)])
    head("hand_net", [("hand_net.initial_hand_coords", 567)])

    # iterative refinement sub-nets (hmdegopose/model.py:232-346): conv_list takes the 64 features + the current estimate
    for net, cin, hs in (("rotation_net", 64 + 27, [("head", 27)]),
                         ("translation_net", 64 + 27, [("head_xy", 18), ("head_z", 9)]),
                         ("hand_net", 64 + 567, [("head", 567)])):
        if iters < 1:
            break
        q = f"{net}.iterative_submodel"
        for i in range(3):
            sep(f"{q}.conv_list.{i}", cin, 64, False)
        for k in range(iters):
            for i in range(3):
                bn(f"{q}.norm_layer.{k}.{i}", 64)
        for hn, cout in hs:
            sep(f"{q}.{hn}", 64, cout, False)
    return s


def raw_weights(seed: int = 0, num_classes: int = 1, iters: int = 0) -> Dict[str, torch.Tensor]:
    """Step (i) and (iv): everything that does not depend on data.  BN running stats are
    initialised to (0, 1)."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    # the iter-0 parameters are drawn first and in the same order as before, so seeds keep their meaning
    names = sorted(param_shapes(num_classes).items())
    names += sorted((k, v) for k, v in param_shapes(num_classes, iters).items() if ".iterative_submodel." in k)
    for name, shp in names:
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.int64)
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shp)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shp)
        elif len(shp) == 4:  # conv weight
            fan_in = shp[1] * shp[2] * shp[3]
            sd[name] = torch.randn(shp, generator=g) * (fan_in ** -0.5)
        elif ".conv.bias" in name:
            sd[name] = torch.randn(shp, generator=g) * 0.1
        elif name.endswith(("_w1", "_w2")):  # BiFPN fusion parameters (negatives exercise the ReLU)
            sd[name] = torch.rand(shp, generator=g) * 1.7 - 0.2
        elif name.endswith(".weight"):  # BN gamma
            sd[name] = torch.rand(shp, generator=g) * 0.6 + 0.7
        elif name.endswith(".bias"):  # BN beta
            sd[name] = torch.randn(shp, generator=g) * 0.3
        else:
            raise KeyError(name)
    for hp, _, _ in header_specs(num_classes):
        scale = 0.5 if hp.startswith("classifier") else 0.1
        sd[hp + ".pointwise_conv.conv.weight"] = sd[hp + ".pointwise_conv.conv.weight"] * scale
    sd["classifier.header.pointwise_conv.conv.bias"] = torch.full_like(
        sd["classifier.header.pointwise_conv.conv.bias"], CLS_BIAS)
    for k in list(sd):   # regression-type refinement heads x0.1 as well
        if ".iterative_submodel.head" in k and k.endswith("pointwise_conv.conv.weight"):
            sd[k] = sd[k] * 0.1
    return sd


def synthetic_state_dict(seed: int = 0, num_classes: int = 1, bn_stats_path: Optional[str] = None,
                         iters: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init ``state_dict``; with ``bn_stats_path`` the BatchNorm running statistics are taken from that .npz."""
    sd = raw_weights(seed, num_classes, iters)
    if bn_stats_path is not None:
        with np.load(bn_stats_path) as z:
            for k in z.files:
                v = torch.from_numpy(z[k].copy())
                assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
                sd[k] = v.to(torch.float32).clone()
    return sd
