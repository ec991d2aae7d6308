"""Timeline of CTA 0 of the last fused-MBConv launch (blk15 at 256x256): python tools/mb_timeline.py [B]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = synthetic.synthetic_state_dict(0, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz"))
s = HmdPoseSession(sd, image_size=256, max_batch=B, precision="fast")
x = np.random.default_rng(0).standard_normal((B, 3, 256, 256)).astype(np.float32)
for _ in range(3):
    s.raw_host(x)
tl = s.debug_read("__mb_timeline")
names = {0: "entry", 1: "tmem", 2: "prefetch issued", 3: "pdl_wait done", 16: "SE start", 17: "FC1 allreduce done", 18: "gate applied",
         19: "project MMA done", 20: "cluster sync", 21: "push done", 22: "cluster sync", 23: "end"}
for li in range(3):
    names.update({4 + 4 * li: f"s{li} operands landed", 5 + 4 * li: f"s{li} expand MMA done", 6 + 4 * li: f"s{li} epilogue done", 7 + 4 * li: f"s{li} stencil done"})
print(", ".join(f"{names.get(i, i)}={tl[i]:.2f}" for i in sorted(names) if tl[i] >= 0))
