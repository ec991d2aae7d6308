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
names = {0: "worker entry", 1: "constants in smem + pdl_wait", 2: "s0 expand done", 3: "s0 epilogue done", 4: "s0 stencil done",
         5: "all slices done", 6: "SE all-reduce done", 7: "gate applied", 8: "project MMA done", 9: "partials stored",
         10: "cluster barrier", 11: "end", 24: "issuer: pdl_wait done", 25: "issuer: x landed", 26: "issuer: expand MMAs done",
         27: "issuer: W_proj + A2 ready"}
t0 = min(tl[i] for i in names if tl[i] >= 0)
print(", ".join(f"{names[i]}={tl[i] - t0:.2f}" for i in sorted(names, key=lambda i: tl[i]) if tl[i] >= 0))
