import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
sd = dict(synthetic.synthetic_state_dict(0, num_classes=90, bn_stats_path='tests/golden/bn_stats_seed0.npz'))
sd["classifier.header.pointwise_conv.conv.weight"] = sd["classifier.header.pointwise_conv.conv.weight"] * 0.03
sd["regressor.header.pointwise_conv.conv.weight"] = sd["regressor.header.pointwise_conv.conv.weight"] * 0.1
x = torch.randn(32, 3, 512, 512, generator=torch.Generator().manual_seed(3)).pin_memory().numpy()
s = HmdPoseSession(sd, image_size=512, max_batch=32, precision="fast")
for _ in range(3): det = s.d0_detect_host(x, 0.2, 0.2, max_out=4096, allow_truncation=True)
ms = []
for _ in range(5):
    det = s.d0_detect_host(x, 0.2, 0.2, max_out=4096, allow_truncation=True); ms.append(s.last_gpu_ms)
print(os.environ.get("HMDPOSE_D0_DENSE", "fused"), "gpu ms per batch of 32:", np.round(ms, 3), "launches", s.last_launch_count, "dets/frame", np.mean([len(d["scores"]) for d in det]))
