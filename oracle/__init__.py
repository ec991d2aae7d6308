"""CPU oracle for the EfficientPose-phi0 inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as a fallback for the CUDA path.

Contents
--------
net_ref.py          fp32 torch-CPU functional restatement of the reference network
                    (backbone.py / efficientnet / efficientdet / hmdegopose/model.py).
                    Pinned against the real reference module (tests/test_oracle_pins.py,
                    run in the build container) and against tests/golden/net_golden_256.npz.
postprocess_ref.py  numpy restatement of anchors, decode, TF filter_detections and the
                    C# receiver's filter.  Anchors + decode are pinned (golden anchor files,
                    reference layers.py lines 1-259 executed directly).  The TensorFlow /
                    OpenCV halves are "parity unpinned": TensorFlow and OpenCvSharp are not
                    in the reference tree nor in this image; their published algorithms are
                    restated and cross-checked against torchvision.ops.nms.
synth_weights.py    seeded, BN-calibrated synthetic weights.
ref_import.py       imports the unmodified reference from /root/reference (build container
                    only; never at GPU-box run time).
make_golden.py      regenerates tests/golden/ from the real reference.
"""
