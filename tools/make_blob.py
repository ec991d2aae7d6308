"""Write the synthetic-weights blob for the C harness: python tools/make_blob.py out.blob"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmd_ego_pose_b200 import packer, synthetic  # noqa: E402

sd = synthetic.synthetic_state_dict(0, bn_stats_path=os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz"))
packer.pack_to_file(sd, sys.argv[1])
print(sys.argv[1], os.path.getsize(sys.argv[1]))
