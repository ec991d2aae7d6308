"""In-situ per-launch times of one step: python tools/steps_table.py [precision] [B]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prec = sys.argv[1] if len(sys.argv) > 1 else "fast"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sd = synthetic.synthetic_state_dict(0, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz"))
s = HmdPoseSession(sd, image_size=256, max_batch=B, precision=prec)
x = np.random.default_rng(0).standard_normal((B, 3, 256, 256)).astype(np.float32)
cam = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))
s.detect_host(x, cam)
ins = s.profile_steps(B, mode=1 | 0x100, reps=20)
print(f"{prec} B={B}: in-situ total {sum(p[2] for p in ins):.3f} ms, {len(ins)} launches")
for name, kern, ms, by, fl in ins:
    print(f"{name[:44]:44s} {kern:22s} {ms * 1e3:7.1f} us {by / 1e6:8.2f} MB {by / max(ms, 1e-9) / 1e6:8.1f} GB/s {fl / max(ms, 1e-9) / 1e9:7.2f} TF/s")
