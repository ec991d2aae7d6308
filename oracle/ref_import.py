"""Import the UNMODIFIED reference network from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: /root/reference does not exist on the GPU box, so
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  Used by
oracle/make_golden.py and tests/test_oracle_pins.py to pin oracle/net_ref.py and
oracle/postprocess_ref.py against the reference's own code.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF_ROOT = "/root/reference/pytorch-sandbox"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "backbone.py"))


def reference_model(num_classes: int = 1, iters: int = 0):
    """backbone.HMDEgoPose(params, num_classes, compound_coef=0, onnx_export=True) in eval mode
    (constructed as evaluate.py:84 does)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from backbone import HMDEgoPose  # type: ignore
        m = HMDEgoPose({"iter": iters}, num_classes=num_classes, compound_coef=0, onnx_export=True)
    return m.eval()


def reference_layers() -> types.ModuleType:
    """The TF-free half of hmdegopose/layers.py (lines 1-259, everything before
    ``import tensorflow`` at :260) executed as a module: RegressBoxes, ClipBoxes,
    RegressTranslation, CalculateTxTy, bbox_transform_inv, translation_transform_inv."""
    src = open(os.path.join(REF_ROOT, "hmdegopose", "layers.py")).read().split("\n")
    cut = next(i for i, l in enumerate(src) if l.strip() == "import tensorflow as tf")
    mod = types.ModuleType("ref_layers_tf_free")
    exec(compile("\n".join(src[:cut]), "layers.py[:%d]" % cut, "exec"), mod.__dict__)
    return mod


def reference_anchor_functions() -> types.ModuleType:
    """generators/utils/anchors.py imports a compiled Cython module at :28 that is only needed
    for training targets; execute the file with that import stubbed out to reach
    ``anchors_for_shape`` (:273-318) and its helpers unchanged."""
    stub = types.ModuleType("generators.utils.compute_overlap")
    stub.compute_overlap = None
    saved = {k: sys.modules.get(k) for k in ("generators", "generators.utils", "generators.utils.compute_overlap")}
    sys.modules["generators"] = types.ModuleType("generators")
    sys.modules["generators.utils"] = types.ModuleType("generators.utils")
    sys.modules["generators.utils.compute_overlap"] = stub
    try:
        mod = types.ModuleType("ref_anchors")
        path = os.path.join(REF_ROOT, "generators", "utils", "anchors.py")
        exec(compile(open(path).read(), path, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
