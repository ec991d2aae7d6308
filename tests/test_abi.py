"""The C-ABI library loads on a CPU-only box and exports every symbol include/hmdpose.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from hmd_ego_pose_b200 import _native, anchors_for_shape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(headers=("hmdpose.h", "hmdpose_internal.h")):
    """every entry point declared in include/: the drop-in surface (hmdpose.h) and the test / profiling hooks
    (hmdpose_internal.h, kept out of the product header)"""
    names = set()
    for hname in headers:
        src = open(os.path.join(ROOT, "include", hname)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hmdpose_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_test_hooks_are_not_in_the_product_header():
    pub = declared_symbols(("hmdpose.h",))
    for n in ("hmdpose_test_gemm", "hmdpose_debug_read", "hmdpose_profile_steps"):
        assert n not in pub and n in declared_symbols(("hmdpose_internal.h",))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _native.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hmdpose.h but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes binding"
    assert set(_native.SYMBOLS) == set(names)
    assert lib.hmdpose_version().startswith(b"hmdpose-b200")


def test_default_config_matches_reference_constants():
    lib = _native.load()
    cfg = _native.Config()
    lib.hmdpose_default_config(ctypes.byref(cfg))
    # train.py:78-81, layers.py:414
    assert cfg.score_threshold == 0.5 and cfg.iou_threshold == 0.5 and cfg.max_detections == 100
    assert cfg.abi_version == _native.ABI_VERSION


def test_host_anchor_arithmetic_matches_golden(gold_dir):
    a, t = anchors_for_shape((256, 256))
    assert np.array_equal(a, np.load(os.path.join(gold_dir, "anchors_256.npy")))
    assert np.array_equal(t, np.load(os.path.join(gold_dir, "translation_anchors_256.npy")))
    a5, t5 = anchors_for_shape((512, 512))
    assert a5.shape == (49104, 4)
    assert np.array_equal(t5, np.load(os.path.join(gold_dir, "translation_anchors_512.npy")))


def test_create_fails_loudly_without_gpu(synth_sd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hmd_ego_pose_b200 import HmdPoseSession
    with pytest.raises(_native.HmdPoseError, match="no CUDA device"):
        HmdPoseSession(synth_sd)


def test_bad_blob_is_rejected():
    lib = _native.load()
    cfg = _native.Config()
    lib.hmdpose_default_config(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    junk = ctypes.create_string_buffer(b"not a blob" * 10)
    rc = lib.hmdpose_create_from_memory(ctypes.byref(cfg), junk, 100, ctypes.byref(h))
    assert rc == -2 and b"magic" in lib.hmdpose_last_error(None)
    assert lib.hmdpose_run_best(None, None, None, None) == -1


def test_pose_packet_bytes_match_the_receiver():
    """hmdpose_pose_packet is host-only: 24 bytes = little-endian fp32 {rvec (rad), t (m)} (Program.cs:279-292)."""
    import struct
    import numpy as np
    from hmd_ego_pose_b200.model import pose_packet
    from oracle import postprocess_ref as pp
    rng = np.random.default_rng(0)
    for _ in range(20):
        best = rng.standard_normal(11).astype(np.float32)
        got = pose_packet(best)
        assert len(got) == 24 and got == pp.csharp_pose_packet(best)
        assert np.array_equal(np.frombuffer(got, "<f4"), best[5:11])      # PoseDataChannel.cs:80-108 reads them back
    assert pose_packet(np.zeros(11, np.float32)) == struct.pack("<6f", *([0.0] * 6))
