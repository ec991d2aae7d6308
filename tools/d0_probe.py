import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from hmd_ego_pose_b200 import HmdPoseSession, synthetic
sd0 = synthetic.synthetic_state_dict(0, num_classes=90, bn_stats_path='tests/golden/bn_stats_seed0.npz')
x = torch.randn(4, 3, 512, 512, generator=torch.Generator().manual_seed(3)).numpy()
for sc in (0.03, 0.06, 0.1, 0.2):
    sd = dict(sd0)
    k = "classifier.header.pointwise_conv.conv.weight"; sd[k] = sd[k] * sc
    k = "regressor.header.pointwise_conv.conv.weight"; sd[k] = sd[k] * 0.1
    s = HmdPoseSession(sd, image_size=512, max_batch=4, precision="fast")
    raw = s.raw_host(x)
    det = s.d0_detect_host(x, 0.2, 0.2, max_out=4096, allow_truncation=True)
    print("scale", sc, "anchors over 0.2:", int((raw[1].max(axis=2) > 0.2).sum(axis=1).mean()), "kept per frame:", [len(d["scores"]) for d in det], "gpu ms", round(s.last_gpu_ms, 2))
    s.close()
