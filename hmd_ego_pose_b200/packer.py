"""Weight packer: reference ``state_dict`` -> BN-folded blob consumed by libhmdpose.so.

Input names are those of ``backbone.HMDEgoPose.state_dict()`` (pytorch-sandbox/backbone.py:13,
layout in SURVEY.md appendix A.6), optionally carrying the ``model.`` / ``model.module.`` prefix that
training checkpoints have (pytorch-sandbox/evaluate.py:105-116).

Folding rules (eval-mode BatchNorm, eps = 1e-3, done in float64, stored as float32):
  conv(no bias) -> BN :  W' = W * g/sqrt(v+eps),           b' = beta - mean * g/sqrt(v+eps)
  conv(+bias)  -> BN :  W' = W * g/sqrt(v+eps),           b' = (b - mean) * g/sqrt(v+eps) + beta
  head trunks: the pointwise conv is shared by the 5 pyramid levels but each level has its own BN
  (efficientdet/model.py:353-357) -> five folded copies.
  BiFPN fast-attention weights: relu(p) / (sum(relu(p)) + 1e-4) in float32, exactly as
  efficientdet/model.py:212-213 evaluates them at run time.

Blob layout (little endian):
  header   : 8s magic "HMDPOSEW", u32 version=1, u32 n_tensors, u32 num_classes, u32 reserved
  table    : n_tensors x { 96s name, u32 ndim, 4 x u32 dims, u32 pad, u64 byte offset into data, u64 count }
  data     : float32 tensors, each 64-byte aligned, starting at the next 64-byte boundary
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np
import torch

BN_EPS = 1e-3
# (kernel, stride, expand, cin, cout, skip): EfficientNet-B0 as instantiated by the reference
# (efficientnet/utils.py:235-241).  Kept here (product code must not import oracle/).
B0_BLOCKS = [
    (3, 1, 1, 32, 16, False), (3, 2, 6, 16, 24, False), (3, 1, 6, 24, 24, True), (5, 2, 6, 24, 40, False),
    (5, 1, 6, 40, 40, True), (3, 2, 6, 40, 80, False), (3, 1, 6, 80, 80, True), (3, 1, 6, 80, 80, True),
    (5, 1, 6, 80, 112, False), (5, 1, 6, 112, 112, True), (5, 1, 6, 112, 112, True), (5, 2, 6, 112, 192, False),
    (5, 1, 6, 192, 192, True), (5, 1, 6, 192, 192, True), (5, 1, 6, 192, 192, True), (3, 1, 6, 192, 320, False),
]
NODES = ("conv6_up", "conv5_up", "conv4_up", "conv3_up", "conv4_down", "conv5_down", "conv6_down", "conv7_down")
FUSE = ("p6_w1", "p5_w1", "p4_w1", "p3_w1", "p4_w2", "p5_w2", "p6_w2", "p7_w2")
PROJ = (("p3_dc", "p3_down_channel"), ("p4_dc", "p4_down_channel"), ("p5_dc", "p5_down_channel"),
        ("p5_to_p6", "p5_to_p6"), ("p4_dc2", "p4_down_channel_2"), ("p5_dc2", "p5_down_channel_2"))
HEADS = (("box", "regressor", ("header",)), ("cls", "classifier", ("header",)),
         ("rot", "rotation_net", ("initial_rotation",)),
         ("trans", "translation_net", ("initial_translation_xy", "initial_translation_z")),
         ("hand", "hand_net", ("initial_hand_coords",)))


def strip_prefix(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Drop the ``model.`` / ``model.module.`` checkpoint prefix (evaluate.py:105-116)."""
    out = OrderedDict()
    for k, v in sd.items():
        for pre in ("model.module.", "model."):
            if k.startswith(pre):
                k = k[len(pre):]
                break
        out[k] = v
    return out


def _np(t) -> np.ndarray:
    return t.detach().cpu().to(torch.float64).numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float64)


def fold(sd: Mapping[str, torch.Tensor]) -> "OrderedDict[str, np.ndarray]":
    """Return the folded fp32 tensors keyed by the names libhmdpose expects."""
    sd = strip_prefix(sd)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def bn_scale_shift(p):
        g, beta = _np(sd[p + ".weight"]), _np(sd[p + ".bias"])
        mean, var = _np(sd[p + ".running_mean"]), _np(sd[p + ".running_var"])
        s = g / np.sqrt(var + BN_EPS)
        return s, beta - mean * s

    def put(name, arr):
        out[name] = np.ascontiguousarray(arr, dtype=np.float32)

    def conv_bn(name, wkey, bkey, bnp):
        w = _np(sd[wkey])
        if w.ndim == 4 and w.shape[2:] == (1, 1):
            w = w[:, :, 0, 0]
        s, sh = bn_scale_shift(bnp)
        b = _np(sd[bkey]) if bkey else 0.0
        put(name + ".w", w * s.reshape((-1,) + (1,) * (w.ndim - 1)))
        put(name + ".b", b * s + sh)

    bb = "backbone_net.model"
    conv_bn("stem", bb + "._conv_stem.conv.weight", None, bb + "._bn0")
    for i, (k, s, e, cin, cout, skip) in enumerate(B0_BLOCKS):
        p, q = f"{bb}._blocks.{i}", f"blk{i}"
        if e != 1:
            conv_bn(q + ".exp", p + "._expand_conv.conv.weight", None, p + "._bn0")
        w = _np(sd[p + "._depthwise_conv.conv.weight"])[:, 0]              # [C,k,k]
        sc, sh = bn_scale_shift(p + "._bn1")
        put(q + ".dw.w", w * sc[:, None, None])
        put(q + ".dw.b", sh)
        put(q + ".se_r.w", _np(sd[p + "._se_reduce.conv.weight"])[:, :, 0, 0])
        put(q + ".se_r.b", _np(sd[p + "._se_reduce.conv.bias"]))
        put(q + ".se_e.w", _np(sd[p + "._se_expand.conv.weight"])[:, :, 0, 0])
        put(q + ".se_e.b", _np(sd[p + "._se_expand.conv.bias"]))
        conv_bn(q + ".proj", p + "._project_conv.conv.weight", None, p + "._bn2")

    for c in range(3):
        p, q = f"bifpn.{c}", f"bifpn{c}"
        for n in NODES:
            put(f"{q}.{n}.dw.w", _np(sd[f"{p}.{n}.depthwise_conv.conv.weight"])[:, 0])
            conv_bn(f"{q}.{n}.pw", f"{p}.{n}.pointwise_conv.conv.weight", f"{p}.{n}.pointwise_conv.conv.bias",
                    f"{p}.{n}.bn")
        for n in FUSE:
            w = torch.relu(sd[f"{p}.{n}"].detach().cpu().float())
            w = w / (torch.sum(w, dim=0) + 1e-4)                           # float32, as the reference
            pad = np.zeros(3, np.float32)
            pad[: w.numel()] = w.numpy()
            put(f"{q}.fw.{n}", pad)
        if c == 0:
            for short, long in PROJ:
                conv_bn(f"{q}.{short}", f"{p}.{long}.0.conv.weight", f"{p}.{long}.0.conv.bias", f"{p}.{long}.1")

    for short, long, headers in HEADS:
        if short in ("rot", "trans", "hand") and not any(k.startswith(long + ".") for k in sd):
            continue  # EfficientDet checkpoint (efficientdet/model.py:420-...): detector-only blob for the D0 variant
        for i in range(3):
            put(f"head.{short}.l{i}.dw.w", _np(sd[f"{long}.conv_list.{i}.depthwise_conv.conv.weight"])[:, 0])
            for lvl in range(5):
                conv_bn(f"head.{short}.l{i}.lvl{lvl}.pw", f"{long}.conv_list.{i}.pointwise_conv.conv.weight",
                        f"{long}.conv_list.{i}.pointwise_conv.conv.bias", f"{long}.bn_list.{lvl}.{i}")
                # the same fold kept factored (shared weights x per-level scale): the implicit-GEMM kernel keeps ONE set
                # of tap matrices resident for the five levels and applies the scale in its epilogue
                put(f"head.{short}.l{i}.lvl{lvl}.pw.scale", bn_scale_shift(f"{long}.bn_list.{lvl}.{i}")[0])
            put(f"head.{short}.l{i}.pw.raw", _np(sd[f"{long}.conv_list.{i}.pointwise_conv.conv.weight"])[:, :, 0, 0])
        for j, hn in enumerate(headers):
            put(f"head.{short}.hdr{j}.dw.w", _np(sd[f"{long}.{hn}.depthwise_conv.conv.weight"])[:, 0])
            put(f"head.{short}.hdr{j}.pw.w", _np(sd[f"{long}.{hn}.pointwise_conv.conv.weight"])[:, :, 0, 0])
            put(f"head.{short}.hdr{j}.pw.b", _np(sd[f"{long}.{hn}.pointwise_conv.conv.bias"]))
    # --iter 1 refinement sub-nets (hmdegopose/model.py:232-346).  The reference's zip() over norm_layer (one entry per
    # iteration step) stops after conv_list[0], so exactly one (64 + P) -> 64 separable conv + norm_layer[0][0] + swish
    # feeds the refinement head(s); conv_list[1:], norm_layer[0][1:] never run and are not packed.
    for short, long, hnames in (("rot", "rotation_net", ("head",)), ("trans", "translation_net", ("head_xy", "head_z")),
                                ("hand", "hand_net", ("head",))):
        q = f"{long}.iterative_submodel"
        if f"{q}.conv_list.0.depthwise_conv.conv.weight" not in sd:
            continue
        if f"{q}.norm_layer.1.0.weight" in sd:
            raise NotImplementedError("more than one refinement iteration does not run in the reference either "
                                      "(conv_list[1] takes 91/631 channels but receives 64)")
        put(f"head.{short}.it.dw.w", _np(sd[f"{q}.conv_list.0.depthwise_conv.conv.weight"])[:, 0])
        conv_bn(f"head.{short}.it.pw", f"{q}.conv_list.0.pointwise_conv.conv.weight",
                f"{q}.conv_list.0.pointwise_conv.conv.bias", f"{q}.norm_layer.0.0")
        for j, hn in enumerate(hnames):
            put(f"head.{short}.it.hdr{j}.dw.w", _np(sd[f"{q}.{hn}.depthwise_conv.conv.weight"])[:, 0])
            put(f"head.{short}.it.hdr{j}.pw.w", _np(sd[f"{q}.{hn}.pointwise_conv.conv.weight"])[:, :, 0, 0])
            put(f"head.{short}.it.hdr{j}.pw.b", _np(sd[f"{q}.{hn}.pointwise_conv.conv.bias"]))
    return out


def num_classes_of(sd: Mapping[str, torch.Tensor]) -> int:
    sd = strip_prefix(sd)
    return int(sd["classifier.header.pointwise_conv.conv.weight"].shape[0]) // 9


def pack(sd: Mapping[str, torch.Tensor]) -> bytes:
    """state_dict -> blob bytes."""
    tensors = fold(sd)
    ncls = num_classes_of(sd)
    table = b""
    data = bytearray()
    for name, arr in tensors.items():
        assert len(name) < 96 and arr.ndim <= 4, name
        off = len(data)
        dims = list(arr.shape) + [0] * (4 - arr.ndim)
        table += struct.pack("<96sI4IIQQ", name.encode(), arr.ndim, *dims, 0, off, arr.size)
        data += arr.tobytes()
        data += b"\0" * (-len(data) % 64)
    header = struct.pack("<8sIIII", b"HMDPOSEW", 1, len(tensors), ncls, 0)
    head = header + table
    head += b"\0" * (-len(head) % 64)
    return bytes(head) + bytes(data)


def pack_to_file(sd: Mapping[str, torch.Tensor], path: str) -> str:
    with open(path, "wb") as f:
        f.write(pack(sd))
    return path
