"""BASELINE.json configs[3] sanity: EfficientPose-phi0 512x512, batch 64 per GPU, fast mode, host API (4 micro-batches of 16)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmd_ego_pose_b200 import HmdPoseSession, synthetic  # noqa: E402

sd = synthetic.synthetic_state_dict(0, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz"))
B, S = 64, 512
sess = HmdPoseSession(sd, image_size=S, max_batch=B, precision="fast")
rng = np.random.default_rng(0)
x = rng.standard_normal((B, 3, S, S)).astype(np.float32)
cam = np.tile(np.array([[960, 960, 256, 256, 1000, 1]], np.float32), (B, 1))
out = sess.detect_host(x, cam)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    out = sess.detect_host(x, cam)
dt = (time.perf_counter() - t0) / n
kept = (out["scores"] > 0).sum(axis=1)
print(f"512x512 B={B}: {dt * 1e3:.2f} ms per batch through hmdpose_run_detect (host frames) = {B / dt:.0f} frames/s; "
      f"gpu {sess.last_gpu_ms:.2f} ms; launches {sess.last_launch_count}; detections per frame min/max {kept.min()}/{kept.max()}")
