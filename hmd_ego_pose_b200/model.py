"""Host-side mirror of the reference's Python interface for the inference hot path.

``TrainModelWithLoss`` has the constructor shape, ``.eval()`` and ``.forward(imgs, camera_params,
is_losses=False, params=...)`` of pytorch-sandbox/train.py:18-85 and returns the same six tensors
(boxes, scores, labels, rotation, translation, hand), -1 padded to ``max_detections`` rows, but
  * for EVERY image of the batch (``[B,100,...]``); ``compat_last_only=True`` reproduces the
    reference, which keeps only the last image (hmdegopose/layers.py:466-482);
  * on the GPU (the reference moves all five head tensors to the CPU for TensorFlow,
    layers.py:448-452); ``compat_cpu=True`` returns CPU tensors like the reference.
All arithmetic happens in libhmdpose.so (hand-written CUDA); PyTorch only owns memory and streams.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _native, packer
from ._native import Config, check


def anchors_for_shape(image_shape: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """generators/utils/anchors.py:273-318 for square inputs: (N,4) x1,y1,x2,y2 boxes and (N,3)
    cx,cy,stride translation anchors, float32.  Host arithmetic of libhmdpose (no GPU needed)."""
    lib = _native.load()
    size = int(image_shape[0])
    if len(image_shape) > 1 and int(image_shape[1]) != size:
        raise ValueError("only square inputs are on the hot path (params['img_size'] = (S, S))")
    n = lib.hmdpose_compute_anchors(size, None, None, 0)
    a = np.empty((n, 4), np.float32)
    t = np.empty((n, 3), np.float32)
    rc = lib.hmdpose_compute_anchors(size, a.ctypes.data, t.ctypes.data, n)
    if rc < 0:
        raise _native.HmdPoseError(f"hmdpose_compute_anchors failed: {rc}")
    return a, t


def d0_anchors(image_size: int) -> np.ndarray:
    """efficientdet/utils.py:76-139 ``Anchors.forward`` for a square input: (N,4) float32 (y1,x1,y2,x2)."""
    lib = _native.load()
    n = lib.hmdpose_compute_anchors_d0(int(image_size), None, 0)
    a = np.empty((n, 4), np.float32)
    rc = lib.hmdpose_compute_anchors_d0(int(image_size), a.ctypes.data, n)
    if rc < 0:
        raise _native.HmdPoseError(f"hmdpose_compute_anchors_d0 failed: {rc}")
    return a


def pose_packet(best11: np.ndarray) -> bytes:
    """Program.cs:279-292: the 24-byte "pose" data-channel message (rvec rad, t m; little-endian fp32) of one
    ``best_host`` result.  Host arithmetic of libhmdpose (no GPU needed)."""
    lib = _native.load()
    b = np.ascontiguousarray(best11, np.float32).reshape(_native.BEST_LEN)
    out = np.empty(24, np.uint8)
    rc = lib.hmdpose_pose_packet(b.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise _native.HmdPoseError(f"hmdpose_pose_packet failed: {rc}")
    return out.tobytes()


class HmdPoseSession:
    """Owns one libhmdpose handle (the ``InferenceSession`` / loaded-model analogue)."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], image_size: int = 256, max_batch: int = 1,
                 device: int = 0, precision: str = "fast", score_threshold: float = 0.5,
                 iou_threshold: float = 0.5, max_detections: int = 100, micro_batch: int = 0,
                 use_graph: bool = True):
        self.lib = _native.load()
        cfg = Config()
        self.lib.hmdpose_default_config(ctypes.byref(cfg))
        cfg.image_size, cfg.max_batch, cfg.device = int(image_size), int(max_batch), int(device)
        cfg.precision = {"parity": _native.PRECISION_PARITY, "fast": _native.PRECISION_FAST}[precision]
        cfg.score_threshold, cfg.iou_threshold = float(score_threshold), float(iou_threshold)
        cfg.max_detections, cfg.micro_batch, cfg.use_graph = int(max_detections), int(micro_batch), int(use_graph)
        blob = packer.pack(state_dict)
        cfg.num_classes = packer.num_classes_of(state_dict)
        self.cfg = cfg
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self.lib.hmdpose_create_from_memory(ctypes.byref(cfg), buf, len(blob), ctypes.byref(self.handle))
        if rc != 0:
            msg = self.lib.hmdpose_last_error(None)
            raise _native.HmdPoseError(f"hmdpose_create failed ({rc}): {msg.decode() if msg else ''}")
        self.image_size = int(image_size)
        self.max_batch = int(max_batch)
        self.max_detections = int(max_detections)
        self.num_classes = int(cfg.num_classes)
        self.device = torch.device("cuda", int(device))
        self.num_anchors = self.lib.hmdpose_num_anchors(self.handle)

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.hmdpose_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----
    def _check_imgs(self, imgs: torch.Tensor) -> torch.Tensor:
        if imgs.dim() != 4 or imgs.shape[1] != 3 or imgs.shape[2] != self.image_size or imgs.shape[3] != self.image_size:
            raise ValueError(f"imgs must be (B,3,{self.image_size},{self.image_size}) NCHW (any strides), got {tuple(imgs.shape)}")
        if imgs.shape[0] > self.max_batch:
            raise ValueError(f"batch {imgs.shape[0]} > max_batch {self.max_batch}")
        if imgs.dtype != torch.float32:
            imgs = imgs.float()
        if imgs.device != self.device:
            imgs = imgs.to(self.device)
        return imgs

    def _stream(self) -> int:
        # NULL means "the handle's own stream" in the C ABI, so torch's default stream (handle 0) is passed
        # as cudaStreamLegacy (0x1), which names the same stream explicitly.
        s = torch.cuda.current_stream(self.device).cuda_stream
        return s if s != 0 else 1

    # ---- device-resident calls (inputs/outputs are CUDA tensors; asynchronous) ----
    def forward_raw(self, imgs: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        """HMDEgoPose.forward contract minus the feature maps (backbone.py:104-125):
        regression (B,N,4), classification (B,N,C), rotation (B,N,3), translation_raw (B,N,3), hand (B,N,63)."""
        imgs = self._check_imgs(imgs)
        B, N, C = imgs.shape[0], self.num_anchors, self.num_classes
        o = [torch.empty((B, N, w), dtype=torch.float32, device=self.device) for w in (4, C, 3, 3, _native.NUM_HAND)]
        sb, sc, sh, sw = imgs.stride()
        check(self.lib.hmdpose_run_raw_device(self.handle, imgs.data_ptr(), sb, sc, sh, sw, B,
                                              *[t.data_ptr() for t in o], self._stream()), self.handle)
        return tuple(o)

    def detect(self, imgs: torch.Tensor, camera_params: torch.Tensor) -> List[torch.Tensor]:
        """[boxes (B,D,4), scores (B,D), labels (B,D) int32, rotation (B,D,3), translation (B,D,3) mm,
        hand (B,D,63), kept_anchor_idx (B,D) int32], -1 padded."""
        imgs = self._check_imgs(imgs)
        B, D = imgs.shape[0], self.max_detections
        cam = camera_params.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(cam.shape) != (B, 6):
            raise ValueError("camera_params must be (B,6) = [fx,fy,px,py,tz_scale,image_scale]")
        f32 = dict(dtype=torch.float32, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        boxes, scores, labels = torch.empty((B, D, 4), **f32), torch.empty((B, D), **f32), torch.empty((B, D), **i32)
        rot, trans = torch.empty((B, D, 3), **f32), torch.empty((B, D, 3), **f32)
        hand, idx = torch.empty((B, D, _native.NUM_HAND), **f32), torch.empty((B, D), **i32)
        sb, sc, sh, sw = imgs.stride()
        check(self.lib.hmdpose_run_detect_device(self.handle, imgs.data_ptr(), sb, sc, sh, sw, cam.data_ptr(), B,
                                                 boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                                 rot.data_ptr(), trans.data_ptr(), hand.data_ptr(), idx.data_ptr(),
                                                 self._stream()), self.handle)
        return [boxes, scores, labels, rot, trans, hand, idx]

    # ---- host-buffer calls (numpy in / numpy out; H2D + D2H inside the call) ----
    def detect_host(self, imgs: np.ndarray, cam: np.ndarray) -> Dict[str, np.ndarray]:
        imgs = np.ascontiguousarray(imgs, np.float32)
        cam = np.ascontiguousarray(cam, np.float32)
        B, D = imgs.shape[0], self.max_detections
        out = {"boxes": np.empty((B, D, 4), np.float32), "scores": np.empty((B, D), np.float32),
               "labels": np.empty((B, D), np.int32), "rotation": np.empty((B, D, 3), np.float32),
               "translation": np.empty((B, D, 3), np.float32), "hand": np.empty((B, D, _native.NUM_HAND), np.float32),
               "anchor_idx": np.empty((B, D), np.int32)}
        check(self.lib.hmdpose_run_detect(self.handle, imgs.ctypes.data, cam.ctypes.data, B,
                                          *[out[k].ctypes.data for k in ("boxes", "scores", "labels", "rotation",
                                                                         "translation", "hand", "anchor_idx")]),
              self.handle)
        return out

    def raw_host(self, imgs: np.ndarray) -> Tuple[np.ndarray, ...]:
        imgs = np.ascontiguousarray(imgs, np.float32)
        B, N, C = imgs.shape[0], self.num_anchors, self.num_classes
        o = [np.empty((B, N, w), np.float32) for w in (4, C, 3, 3, _native.NUM_HAND)]
        check(self.lib.hmdpose_run_raw(self.handle, imgs.ctypes.data, B, *[t.ctypes.data for t in o]), self.handle)
        return tuple(o)

    def best_host(self, img: np.ndarray, cam: np.ndarray) -> np.ndarray:
        """C# receiver result for one frame (Program.cs:208-276): 11 floats, see hmdpose.h."""
        img = np.ascontiguousarray(img, np.float32)
        cam = np.ascontiguousarray(cam, np.float32).reshape(6)
        out = np.empty(_native.BEST_LEN, np.float32)
        check(self.lib.hmdpose_run_best(self.handle, img.ctypes.data, cam.ctypes.data, out.ctypes.data), self.handle)
        return out

    # ---- EfficientDet-d0 detection variant (utils/utils.py:90-128) ----
    def _d0_unpack(self, B, max_out, call, allow_truncation=False) -> List[Dict[str, np.ndarray]]:
        rois = np.empty((B, max_out, 4), np.float32)
        cls = np.empty((B, max_out), np.int32)
        scores = np.empty((B, max_out), np.float32)
        idx = np.empty((B, max_out), np.int32)
        cnt = np.empty((B,), np.int32)
        check(call(rois.ctypes.data, cls.ctypes.data, scores.ctypes.data, idx.ctypes.data, cnt.ctypes.data), self.handle)
        if (cnt < 0).any() and not allow_truncation:
            raise _native.HmdPoseError(f"frames {np.nonzero(cnt < 0)[0].tolist()} have more than max_out={max_out} NMS "
                                       "survivors (the reference keeps all): raise max_out (<= 4096)")
        self.last_d0_truncated = (cnt < 0)
        cnt = np.abs(cnt)
        # the reference returns one dict per image with variable-length arrays (empty arrays when nothing passes)
        return [{"rois": rois[b, :cnt[b]].copy(), "class_ids": cls[b, :cnt[b]].astype(np.int64),
                 "scores": scores[b, :cnt[b]].copy(), "anchor_idx": idx[b, :cnt[b]].copy()} for b in range(B)]

    def d0_detect_host(self, imgs: np.ndarray, threshold: float, iou_threshold: float,
                       max_out: int = 512, allow_truncation: bool = False) -> List[Dict[str, np.ndarray]]:
        """EfficientDet forward + ``postprocess`` for host frames (B,3,S,S)."""
        imgs = np.ascontiguousarray(imgs, np.float32)
        B = imgs.shape[0]
        return self._d0_unpack(B, max_out, lambda *o: self.lib.hmdpose_run_d0(
            self.handle, imgs.ctypes.data, B, float(threshold), float(iou_threshold), int(max_out), *o), allow_truncation)

    def d0_postprocess_host(self, regression: np.ndarray, classification: np.ndarray, threshold: float,
                            iou_threshold: float, max_out: int = 512, allow_truncation: bool = False) -> List[Dict[str, np.ndarray]]:
        """``postprocess`` alone on host head tensors (B,N,4) / (B,N,C)."""
        reg = np.ascontiguousarray(regression, np.float32)
        cls = np.ascontiguousarray(classification, np.float32)
        B = reg.shape[0]
        return self._d0_unpack(B, max_out, lambda *o: self.lib.hmdpose_d0_postprocess(
            self.handle, reg.ctypes.data, cls.ctypes.data, B, float(threshold), float(iou_threshold), int(max_out), *o),
            allow_truncation)

    # ---- uint8 frames: pre-processing on the device (generators/colibri_common.py:622-656) ----
    def preprocess_host(self, frames: np.ndarray) -> Tuple[np.ndarray, float]:
        """uint8 RGB frames (B, H, W, 3) -> (float32 (B, S, S, 3) NHWC, scale) exactly as ``preprocess_image`` builds it."""
        f = np.ascontiguousarray(frames, np.uint8)
        B, H, W, _ = f.shape
        out = np.empty((B, self.image_size, self.image_size, 3), np.float32)
        scale = ctypes.c_float(0)
        check(self.lib.hmdpose_preprocess(self.handle, f.ctypes.data, B, H, W, out.ctypes.data, ctypes.byref(scale)),
              self.handle)
        return out, float(scale.value)

    def detect_u8_host(self, frames: np.ndarray, cam: np.ndarray) -> Dict[str, np.ndarray]:
        """Pre-processing + network + post-processing for uint8 RGB frames (B, H, W, 3); the tensor stays on the device."""
        f = np.ascontiguousarray(frames, np.uint8)
        cam = np.ascontiguousarray(cam, np.float32)
        B, H, W, _ = f.shape
        D = self.max_detections
        out = {"boxes": np.empty((B, D, 4), np.float32), "scores": np.empty((B, D), np.float32),
               "labels": np.empty((B, D), np.int32), "rotation": np.empty((B, D, 3), np.float32),
               "translation": np.empty((B, D, 3), np.float32), "hand": np.empty((B, D, _native.NUM_HAND), np.float32),
               "anchor_idx": np.empty((B, D), np.int32)}
        scale = ctypes.c_float(0)
        check(self.lib.hmdpose_run_detect_u8(self.handle, f.ctypes.data, B, H, W, cam.ctypes.data,
                                             *[out[k].ctypes.data for k in ("boxes", "scores", "labels", "rotation",
                                                                            "translation", "hand", "anchor_idx")],
                                             ctypes.byref(scale)), self.handle)
        out["scale"] = float(scale.value)
        return out

    def best_u8_host(self, frame: np.ndarray, cam: np.ndarray) -> Tuple[np.ndarray, float]:
        f = np.ascontiguousarray(frame, np.uint8)
        cam = np.ascontiguousarray(cam, np.float32).reshape(6)
        out = np.empty(_native.BEST_LEN, np.float32)
        scale = ctypes.c_float(0)
        check(self.lib.hmdpose_run_best_u8(self.handle, f.ctypes.data, f.shape[0], f.shape[1], cam.ctypes.data,
                                           out.ctypes.data, ctypes.byref(scale)), self.handle)
        return out, float(scale.value)

    # ---- raw I420 video frames: the C# receiver's frame path (WebRTCNetCoreSandbox/Program.cs:137-200) ----
    def preprocess_i420_host(self, frames: np.ndarray, height: int, width: int, crop_size: int = 256,
                             rescaled_size: int = 512) -> Tuple[np.ndarray, float]:
        """I420 frames (B, height * width * 3 / 2) uint8 -> (float32 (B, S, S, 3) in the Mat's channel order, scale):
        YUV2BGR_YV12 on the I420 buffer, centre crop, rescale, ResizeAndNormalizeMat -- bit-exact against OpenCV."""
        f = np.ascontiguousarray(frames, np.uint8).reshape(-1, height * width * 3 // 2)
        out = np.empty((f.shape[0], self.image_size, self.image_size, 3), np.float32)
        scale = ctypes.c_float(0)
        check(self.lib.hmdpose_preprocess_i420(self.handle, f.ctypes.data, f.shape[0], int(height), int(width),
                                               int(crop_size), int(rescaled_size), out.ctypes.data, ctypes.byref(scale)),
              self.handle)
        return out, float(scale.value)

    def best_i420_host(self, frame: np.ndarray, height: int, width: int, cam: np.ndarray, crop_size: int = 256,
                       rescaled_size: int = 512) -> Tuple[np.ndarray, float]:
        """One I420 frame -> the receiver's result (Program.cs:137-276): 11 floats, see hmdpose.h."""
        f = np.ascontiguousarray(frame, np.uint8).reshape(height * width * 3 // 2)
        cam = np.ascontiguousarray(cam, np.float32).reshape(6)
        out = np.empty(_native.BEST_LEN, np.float32)
        scale = ctypes.c_float(0)
        check(self.lib.hmdpose_run_best_i420(self.handle, f.ctypes.data, int(height), int(width), int(crop_size),
                                             int(rescaled_size), cam.ctypes.data, out.ctypes.data, ctypes.byref(scale)),
              self.handle)
        return out, float(scale.value)

    def packet_host(self, img: np.ndarray, cam: np.ndarray) -> Tuple[bytes, float]:
        """One frame -> (24-byte pose packet, score): hmdpose_run_packet (Program.cs:208-292)."""
        img = np.ascontiguousarray(img, np.float32)
        cam = np.ascontiguousarray(cam, np.float32).reshape(6)
        out = np.empty(24, np.uint8)
        score = ctypes.c_float(0)
        check(self.lib.hmdpose_run_packet(self.handle, img.ctypes.data, cam.ctypes.data, out.ctypes.data,
                                          ctypes.byref(score)), self.handle)
        return out.tobytes(), float(score.value)

    def postprocess_host(self, regression, classification, rotation, translation_raw, hand, cam) -> Dict[str, np.ndarray]:
        arrs = [np.ascontiguousarray(a, np.float32) for a in (regression, classification, rotation, translation_raw, hand, cam)]
        B, D = arrs[0].shape[0], self.max_detections
        out = {"boxes": np.empty((B, D, 4), np.float32), "scores": np.empty((B, D), np.float32),
               "labels": np.empty((B, D), np.int32), "rotation": np.empty((B, D, 3), np.float32),
               "translation": np.empty((B, D, 3), np.float32), "hand": np.empty((B, D, _native.NUM_HAND), np.float32),
               "anchor_idx": np.empty((B, D), np.int32)}
        check(self.lib.hmdpose_postprocess(self.handle, *[a.ctypes.data for a in arrs], B,
                                           *[out[k].ctypes.data for k in ("boxes", "scores", "labels", "rotation",
                                                                          "translation", "hand", "anchor_idx")]),
              self.handle)
        return out

    def filter_boxes_host(self, boxes, classification, rotation, translation, hand) -> Dict[str, np.ndarray]:
        arrs = [np.ascontiguousarray(a, np.float32) for a in (boxes, classification, rotation, translation, hand)]
        B, D = arrs[0].shape[0], self.max_detections
        out = {"boxes": np.empty((B, D, 4), np.float32), "scores": np.empty((B, D), np.float32),
               "labels": np.empty((B, D), np.int32), "rotation": np.empty((B, D, 3), np.float32),
               "translation": np.empty((B, D, 3), np.float32), "hand": np.empty((B, D, _native.NUM_HAND), np.float32),
               "anchor_idx": np.empty((B, D), np.int32)}
        check(self.lib.hmdpose_filter_boxes(self.handle, *[a.ctypes.data for a in arrs], B,
                                            *[out[k].ctypes.data for k in ("boxes", "scores", "labels", "rotation",
                                                                           "translation", "hand", "anchor_idx")]),
              self.handle)
        return out

    def best_from_raw_host(self, regression, classification, rotation, translation_raw, cam) -> np.ndarray:
        arrs = [np.ascontiguousarray(a, np.float32) for a in (regression, classification, rotation, translation_raw, cam)]
        out = np.empty(_native.BEST_LEN, np.float32)
        check(self.lib.hmdpose_best_from_raw(self.handle, *[a.ctypes.data for a in arrs], out.ctypes.data), self.handle)
        return out

    def debug_read(self, name: str) -> np.ndarray:
        n = self.lib.hmdpose_debug_read(self.handle, name.encode(), None, 0)
        if n < 0:
            check(int(n), self.handle)
        out = np.empty(int(n), np.float32)
        n2 = self.lib.hmdpose_debug_read(self.handle, name.encode(), out.ctypes.data, int(n))
        if n2 < 0:
            check(int(n2), self.handle)
        return out

    def profile_steps(self, batch: int, mode: int = 1, reps: int = 5):
        """[(step name, kernel, ms, algorithmic bytes, flops)] for one pass over ``batch`` frames."""
        n = self.lib.hmdpose_profile_steps(self.handle, batch, mode, reps, None, None, None, None, None, 0)
        if n < 0:
            check(int(n), self.handle)
        names, kernels = ctypes.create_string_buffer(64 * n), ctypes.create_string_buffer(64 * n)
        ms, by, fl = np.zeros(n, np.float32), np.zeros(n, np.float64), np.zeros(n, np.float64)
        rc = self.lib.hmdpose_profile_steps(self.handle, batch, mode, reps, names, kernels, ms.ctypes.data,
                                            by.ctypes.data, fl.ctypes.data, n)
        if rc < 0:
            check(int(rc), self.handle)
        dec = lambda buf, i: buf.raw[64 * i:64 * i + 64].split(b"\0")[0].decode()
        return [(dec(names, i), dec(kernels, i), float(ms[i]), float(by[i]), float(fl[i])) for i in range(n)]

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.hmdpose_last_launch_count(self.handle))

    @property
    def last_gpu_ms(self) -> float:
        return float(self.lib.hmdpose_last_gpu_ms(self.handle))


class TrainModelWithLoss(nn.Module):
    """Drop-in for pytorch-sandbox/train.py:18-85 (inference branch only).

    ``model`` may be the reference ``HMDEgoPose`` module (its ``state_dict()`` is packed), or a
    ``state_dict`` itself.  ``params['img_size']`` selects the input side exactly like train.py:35."""

    def __init__(self, model: Union[nn.Module, Mapping[str, torch.Tensor]], max_batch: int = 16, device: int = 0,
                 precision: str = "fast", compat_last_only: bool = False, compat_cpu: bool = False,
                 score_threshold: float = 0.5, max_detections: int = 100, micro_batch: int = 0):
        super().__init__()
        self._state_dict = model.state_dict() if isinstance(model, nn.Module) else dict(model)
        self._opts = dict(max_batch=max_batch, device=device, precision=precision, score_threshold=score_threshold,
                          max_detections=max_detections, micro_batch=micro_batch)
        self.compat_last_only = compat_last_only
        self.compat_cpu = compat_cpu
        self._sessions: Dict[int, HmdPoseSession] = {}

    def session(self, image_size: int) -> HmdPoseSession:
        if image_size not in self._sessions:
            self._sessions[image_size] = HmdPoseSession(self._state_dict, image_size=image_size, **self._opts)
        return self._sessions[image_size]

    def forward(self, imgs, camera_params, is_losses=False, model_3d_points=None, classification_gt=None,
                regression_gt=None, transformation_gt=None, coords_3d_gt=None, params=None):
        if is_losses:
            raise NotImplementedError("training losses are out of scope of the B200 inference path (train.py:41-70)")
        size = int(params["img_size"][0]) if params and "img_size" in params else int(imgs.shape[-1])
        out = self.session(size).detect(imgs, camera_params)[:6]
        if self.compat_last_only:
            out = [t[-1] for t in out]
        if self.compat_cpu:
            out = [t.cpu() for t in out]
        return out

    def forward_raw(self, imgs):
        return self.session(int(imgs.shape[-1])).forward_raw(imgs)
