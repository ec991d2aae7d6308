"""A few detection steps through the host API (for ncu / compute-sanitizer): python tools/one_step.py [precision] [B] [steps] [use_graph]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prec = sys.argv[1] if len(sys.argv) > 1 else "fast"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
use_graph = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
sd = synthetic.synthetic_state_dict(0, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz"))
s = HmdPoseSession(sd, image_size=256, max_batch=B, precision=prec, use_graph=use_graph)
rng = np.random.default_rng(0)
cam = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))
for i in range(steps):
    x = rng.standard_normal((B, 3, 256, 256)).astype(np.float32)
    det = s.detect_host(x, cam)
print("launches/step", s.last_launch_count, "detections", int((det["scores"] > 0).sum()), "gpu ms", round(s.last_gpu_ms, 3))
s.close()
