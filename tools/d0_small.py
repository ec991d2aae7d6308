"""One small EfficientDet-d0 detection call (for compute-sanitizer): python tools/d0_small.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sd = dict(synthetic.synthetic_state_dict(0, num_classes=90, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz")))
sd["classifier.header.pointwise_conv.conv.weight"] = sd["classifier.header.pointwise_conv.conv.weight"] * 0.03
sd["regressor.header.pointwise_conv.conv.weight"] = sd["regressor.header.pointwise_conv.conv.weight"] * 0.1
x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(3)).numpy()
s = HmdPoseSession(sd, image_size=256, max_batch=2, precision="fast", use_graph=False)
det = s.d0_detect_host(x, float(os.environ.get("D0_THR", "0.2")), 0.2, max_out=4096, allow_truncation=True)
print("launches/step", s.last_launch_count, "detections", [len(d["scores"]) for d in det])
s.close()
