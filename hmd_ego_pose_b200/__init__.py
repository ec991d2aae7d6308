"""hmd_ego_pose_b200 -- B200-native (sm_100a) EfficientPose-phi0 inference hot path of
doughtmw/hmd-ego-pose behind the reference's own call sites.

Python boundary : ``TrainModelWithLoss`` (mirror of pytorch-sandbox/train.py:18-85, inference branch)
C# / C boundary : ``include/hmdpose.h`` (libhmdpose.so), see INTEGRATION.md

Everything computes in hand-written CUDA inside ``lib/libhmdpose.so``; there is no CPU or PyTorch
fallback -- importing works without a GPU (so the ABI can be inspected), running does not.
"""
from . import _native, packer, sharding, synthetic  # noqa: F401
from .model import HmdPoseSession, TrainModelWithLoss, anchors_for_shape

__all__ = ["HmdPoseSession", "TrainModelWithLoss", "anchors_for_shape", "_native"]
