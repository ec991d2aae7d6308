"""Regenerate tests/golden/ from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Everything written here comes from the reference's own code or data files:
  anchors_256.npy, translation_anchors_{256,512}.npy   <- onnx-models/*.txt (shipped goldens)
  camera_params.json                                    <- onnx-models/camera_params*.txt
  input_256.npy                                         <- onnx-models/input.npy
  bn_stats_seed0.npz     calibrated BN statistics of oracle.synth_weights (seed 0, S=256)
  net_golden_256.npz     backbone.HMDEgoPose (reference module) outputs on input_256.npy and on a
                         seeded randn frame with synthetic_weights(seed=0)
  post_golden_256.npz    reference hmdegopose/layers.py (TF-free half) decode of those outputs
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import postprocess_ref as pp
from . import ref_import, synth_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ONNX = os.path.join(ref_import.REF_ROOT, "onnx-models")
HAND_STRIDE = 16


def main() -> None:
    assert ref_import.available(), "reference tree not mounted"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # --- shipped golden data files ---------------------------------------------------
    a256 = np.loadtxt(os.path.join(ONNX, "anchors_256.txt"), dtype=np.float64).astype(np.float32).reshape(-1, 4)
    t256 = np.loadtxt(os.path.join(ONNX, "translation_anchors_256.txt"), dtype=np.float64).astype(np.float32).reshape(-1, 3)
    t512 = np.loadtxt(os.path.join(ONNX, "translation_anchors_512.txt"), dtype=np.float64).astype(np.float32).reshape(-1, 3)
    np.save(os.path.join(GOLD, "anchors_256.npy"), a256)
    np.save(os.path.join(GOLD, "translation_anchors_256.npy"), t256)
    np.save(os.path.join(GOLD, "translation_anchors_512.npy"), t512)
    cams = {}
    for name in ("camera_params", "camera_params_hololens2", "camera_params_webcam"):
        cams[name] = [float(v) for v in open(os.path.join(ONNX, name + ".txt")).read().split()]
    json.dump(cams, open(os.path.join(GOLD, "camera_params.json"), "w"), indent=1)
    x_fixed = np.load(os.path.join(ONNX, "input.npy")).astype(np.float32)
    assert x_fixed.shape == (1, 3, 256, 256)
    np.save(os.path.join(GOLD, "input_256.npy"), x_fixed)

    # --- synthetic weights + calibrated BN stats -----------------------------------------
    sd = synth_weights.raw_weights(0)
    stats = synth_weights.calibrate(sd, 256, 0)
    np.savez_compressed(os.path.join(GOLD, "bn_stats_seed0.npz"), **{k: v.numpy() for k, v in stats.items()})

    # --- reference network outputs ------------------------------------------------------
    model = ref_import.reference_model()
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(1234)
    x = torch.cat([torch.from_numpy(x_fixed), torch.randn(1, 3, 256, 256, generator=g)], 0)
    with torch.no_grad():
        feats, reg, cls, rot, tr, hand = model(x)
    net = {"regression": reg.numpy(), "classification": cls.numpy(), "rotation": rot.numpy(),
           "translation_raw": tr.numpy(), "hand_sub": hand[:, ::HAND_STRIDE].numpy().copy(),
           "hand_stride": np.int32(HAND_STRIDE)}
    for i, f in enumerate(feats):
        net[f"feat{i + 3}"] = f.numpy()
    np.savez_compressed(os.path.join(GOLD, "net_golden_256.npz"), **net)

    # --- reference decode (layers.py TF-free half) ----------------------------------------
    L = ref_import.reference_layers()
    RA = ref_import.reference_anchor_functions()
    anchors, tanchors = RA.anchors_for_shape((256, 256))
    cam = torch.tensor([cams["camera_params"], cams["camera_params_hololens2"]], dtype=torch.float32)
    with torch.no_grad():
        boxes = L.ClipBoxes()(x, L.RegressBoxes()(torch.tensor(anchors)[None], reg[..., :4]))
        txy = L.RegressTranslation()(torch.tensor(tanchors)[None], tr)
        trans = L.CalculateTxTy()(txy, fx=cam[:, 0], fy=cam[:, 1], px=cam[:, 2], py=cam[:, 3],
                                  tz_scale=cam[:, 4], image_scale=cam[:, 5])
    post = {"cam": cam.numpy(), "boxes": boxes.numpy(), "translation": trans.numpy()}
    # restated (UNPINNED) filter outputs, kept as a regression fixture of the oracle itself
    det = pp.detect(reg.numpy(), cls.numpy(), rot.numpy(), tr.numpy(), hand.numpy(), cam.numpy(), 256)
    for b, d in enumerate(det):
        for k, v in d.items():
            post[f"det{b}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, "post_golden_256.npz"), **post)
    sizes = {f: os.path.getsize(os.path.join(GOLD, f)) for f in sorted(os.listdir(GOLD))}
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
