/*
 * hmdpose.h -- C ABI of libhmdpose.so: the B200-native EfficientPose-phi0 inference hot path
 * (EfficientNet-B0 backbone + 3x BiFPN + box/class/rotation/translation/hand heads + post-processing).
 *
 * This is the ONLY boundary.  Plain pointers and sizes, cdecl, no C++/torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree):
 *
 *   Python side  pytorch-sandbox/train.py:23-85       TrainModelWithLoss.forward (inference branch),
 *                                                      called at pytorch-sandbox/eval/common.py:400
 *   C# side      unity-sandbox/WebRTCNetCoreSandbox/Program.cs:76   new InferenceSession(...)
 *                unity-sandbox/WebRTCNetCoreSandbox/Program.cs:219  Session.Run(inputOnnxValues)
 *                unity-sandbox/WebRTCNetCoreSandbox/Program.cs:247-270  format_translation /
 *                                                      format_bboxes / filter_detections
 *                (verbatim twin: unity-sandbox/OpenCVDNNSandboxNetCore/Program.cs:103-153)
 *
 * Conventions: every function returns 0 on success or a negative HMDPOSE_E_* code and never
 * throws across the ABI; hmdpose_last_error() gives the message.  The library owns all device
 * memory, pinned staging and CUDA graphs; the caller owns every buffer it passes.  A handle may be
 * used from any thread but not concurrently (internal mutex).  There is no CPU fallback: without a
 * CUDA device hmdpose_create fails with HMDPOSE_E_CUDA.
 */
#ifndef HMDPOSE_H_
#define HMDPOSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMDPOSE_ABI_VERSION 1

#define HMDPOSE_OK 0
#define HMDPOSE_E_ARG -1      /* bad argument (null pointer, batch > max_batch, ...) */
#define HMDPOSE_E_WEIGHTS -2  /* weight blob missing / malformed / wrong architecture */
#define HMDPOSE_E_CUDA -3     /* CUDA error or no CUDA device */
#define HMDPOSE_E_STATE -4    /* call not valid for this handle (e.g. out-of-scope feature) */

/* arithmetic modes behind the same ABI */
#define HMDPOSE_PRECISION_PARITY 0 /* fp32 activations + fp32 FFMA everywhere: end-to-end parity mode */
#define HMDPOSE_PRECISION_FAST 1   /* fp16 activations, tcgen05 kind::f16 GEMMs with fp32 TMEM accumulate */

#define HMDPOSE_NUM_HAND 63  /* hand-joint parameters per anchor (hmdegopose/model.py:113) */
#define HMDPOSE_BEST_LEN 11  /* floats written by hmdpose_run_best */

typedef struct hmdpose hmdpose_t;

typedef struct hmdpose_config {
  int abi_version;        /* = HMDPOSE_ABI_VERSION */
  int image_size;         /* S: 256 or 512 -- params['img_size'], train.py:35.  parity mode: any multiple of 128;
                             fast mode: a power of two >= 256 (anything else fails at create with HMDPOSE_E_ARG) */
  int max_batch;          /* largest B accepted by the run_* calls */
  int device;             /* CUDA ordinal */
  int precision;          /* HMDPOSE_PRECISION_* */
  int num_classes;        /* 1 for HMD-EgoPose (evaluate.py:84) */
  float score_threshold;  /* 0.5, train.py:80 */
  float iou_threshold;    /* 0.5, layers.py:414 */
  int max_detections;     /* 100, train.py:81 */
  int micro_batch;        /* 0 = library default; frames per internal pass (sized for the 126 MB L2) */
  int use_graph;          /* 1 = replay the per-batch launch sequence as a CUDA graph */
} hmdpose_config_t;

/* Fill *cfg with the reference's hard-coded defaults (train.py:78-81, layers.py:408-416). */
void hmdpose_default_config(hmdpose_config_t* cfg);

/*
 * Replaces: new InferenceSession(model.onnx, options) (Program.cs:57-78) and
 * HMDEgoPose(...).load_state_dict(...) + TrainModelWithLoss(model).eval() (evaluate.py:84-124).
 * weights_path: blob written by hmd_ego_pose_b200.packer (BN-folded tensors).
 */
int hmdpose_create(const char* weights_path, int image_size, int max_batch, int device,
                   float score_threshold, float iou_threshold, int max_detections, hmdpose_t** out);
int hmdpose_create_ex(const hmdpose_config_t* cfg, const char* weights_path, hmdpose_t** out);
/* Same, from a blob already in host memory (the PyTorch-side wrapper packs a state_dict in memory). */
int hmdpose_create_from_memory(const hmdpose_config_t* cfg, const void* blob, size_t blob_bytes,
                               hmdpose_t** out);
void hmdpose_destroy(hmdpose_t* h);
const char* hmdpose_last_error(const hmdpose_t* h); /* h may be NULL: last create error */

/* N = 9 * sum_l ceil(S/2^l)^2, l = 3..7  (generators/utils/anchors.py:273-318): 12276 @256, 49104 @512 */
int hmdpose_num_anchors(const hmdpose_t* h);
int hmdpose_num_classes(const hmdpose_t* h);
/* Host copy of the constant anchors the library precomputes at create time instead of on every
 * forward (train.py:36): boxes (N,4) x1,y1,x2,y2 and translation anchors (N,3) cx,cy,stride. */
int hmdpose_get_anchors(const hmdpose_t* h, float* anchors_n4, float* translation_anchors_n3);
/* The same anchor arithmetic without a handle or a GPU (host-only; used by CPU tests). */
int hmdpose_compute_anchors(int image_size, float* anchors_n4, float* translation_anchors_n3, int capacity_n);

/*
 * Session.Run twin (Program.cs:219; tensor contract hmdegopose/misc_utils.py:77-83):
 * input (B,3,S,S) fp32 NCHW contiguous HOST memory -> the five head tensors, HOST memory, fp32:
 * regression (B,N,4), classification (B,N,C) after sigmoid, rotation (B,N,3),
 * translation_raw (B,N,3), hand (B,N,63).  Any output pointer may be NULL (skipped).
 */
int hmdpose_run_raw(hmdpose_t* h, const float* input_nchw, int batch, float* regression,
                    float* classification, float* rotation, float* translation_raw, float* hand);

/*
 * TrainModelWithLoss.forward(imgs, camera_params, is_losses=False) twin (train.py:72-85) for EVERY
 * image of the batch (the reference returns only the last one, layers.py:466-482):
 * cam6 (B,6) = [fx,fy,px,py,tz_scale,image_scale] (generators/colibri_common.py:658-678).
 * Outputs, padded with -1 exactly like layers.py:377-384, max_detections = D rows per image:
 * boxes (B,D,4) x1,y1,x2,y2; scores (B,D); labels (B,D) int32; rotation (B,D,3) in units of pi;
 * translation (B,D,3) mm; hand (B,D,63); kept_anchor_idx (B,D) int32 (anchor row of each kept box).
 * Any output pointer may be NULL.
 */
int hmdpose_run_detect(hmdpose_t* h, const float* input_nchw, const float* cam6, int batch,
                       float* boxes, float* scores, int32_t* labels, float* rotation,
                       float* translation, float* hand, int32_t* kept_anchor_idx);

/*
 * C# receiver twin: Program.cs:208-276 for one frame (batch 1).  out11 =
 * [score, rect.X, rect.Y, rect.Width, rect.Height, rvec.x, rvec.y, rvec.z (rad), t.x, t.y, t.z (m)]
 * with the Rect fields exactly as Program.cs:840-845 builds them; zeros if no score > threshold
 * (Program.cs:929-932).  out11[5..10] is the 24-byte pose packet of Program.cs:279-292.
 */
int hmdpose_run_best(hmdpose_t* h, const float* input_nchw, const float* cam6, float* out11);

/*
 * Post-processing alone on caller-supplied head tensors (HOST memory): format_translation +
 * format_bboxes + FilterDetections (loss.py:12-51, layers.py:264-400).  Same outputs as run_detect.
 */
int hmdpose_postprocess(hmdpose_t* h, const float* regression, const float* classification,
                        const float* rotation, const float* translation_raw, const float* hand,
                        const float* cam6, int batch, float* boxes, float* scores, int32_t* labels,
                        float* rotation_out, float* translation_out, float* hand_out,
                        int32_t* kept_anchor_idx);
/* filter_detections alone (layers.py:264-400) on already-decoded boxes (B,N,4) and translations
 * (B,N,3): the entry point for bit-exact NMS / top-k checks on identical inputs. */
int hmdpose_filter_boxes(hmdpose_t* h, const float* boxes_in, const float* classification,
                         const float* rotation, const float* translation, const float* hand, int batch,
                         float* boxes, float* scores, int32_t* labels, float* rotation_out,
                         float* translation_out, float* hand_out, int32_t* kept_anchor_idx);
/* C# post-processing alone (Program.cs:247-270) on one frame's head tensors. */
int hmdpose_best_from_raw(hmdpose_t* h, const float* regression, const float* classification,
                          const float* rotation, const float* translation_raw, const float* cam6,
                          float* out11);

/*
 * Frame pre-processing on the device: generators/colibri_common.py:622-656 `preprocess_image`
 * (C# twin ResizeAndNormalizeMat, Program.cs:397-445).  images: `batch` uint8 RGB frames of height x width x 3, HOST
 * memory.  The long side is resized to the network size with OpenCV's INTER_LINEAR arithmetic for 8-bit data, then
 * /255, ImageNet mean / std, zero padding bottom / right.  *scale (may be NULL) receives the resize factor the caller
 * puts into camera_params[5] (colibri_common.py:658-678).
 *   hmdpose_preprocess      -> the float32 tensor itself, (B, S, S, 3) NHWC, HOST memory (what the reference builds)
 *   hmdpose_run_detect_u8   = pre-processing + hmdpose_run_detect without the tensor ever leaving the device
 *   hmdpose_run_best_u8     = pre-processing + hmdpose_run_best for one frame (the receiver's per-frame path)
 */
int hmdpose_preprocess(hmdpose_t* h, const uint8_t* images, int batch, int height, int width, float* out_nhwc,
                       float* scale);
int hmdpose_run_detect_u8(hmdpose_t* h, const uint8_t* images, int batch, int height, int width, const float* cam6,
                          float* boxes, float* scores, int32_t* labels, float* rotation, float* translation,
                          float* hand, int32_t* kept_anchor_idx, float* scale);
int hmdpose_run_best_u8(hmdpose_t* h, const uint8_t* image, int height, int width, const float* cam6, float* out11,
                        float* scale);

/*
 * The C# receiver's frame path on the device (WebRTCNetCoreSandbox/Program.cs:137-200, 381-445): `frames` are raw I420
 * video frames as I420AVideoFrame.CopyTo delivers them (Y, U, V planes, height * width * 3 / 2 bytes each, HOST memory).
 *   Cv2.CvtColor(YUV2BGR_YV12) on that buffer (Program.cs:146-160: the chroma planes are read swapped),
 *   CenterCropAndRescaleMat(crop_size -> rescaled_size x rescaled_size)          (Program.cs:170-173, 381-395: 256 -> 512),
 *   ResizeAndNormalizeMat to the network size                                    (Program.cs:397-445)
 * in one kernel, bit-exact against the OpenCV calls (every resize stage rounds to uint8, the normalisation runs in
 * float32 as OpenCV evaluates it on CV_32F data).  The tensor is in the Mat's channel order, exactly what
 * CvDnn.BlobFromImage(swapRB = false) hands to the network (Program.cs:192-200).
 *   hmdpose_preprocess_i420 -> the float32 tensor, (B, S, S, 3) NHWC, HOST memory
 *   hmdpose_run_best_i420   = this + hmdpose_run_best: the receiver's whole per-frame region Program.cs:137-276
 */
int hmdpose_preprocess_i420(hmdpose_t* h, const uint8_t* frames, int batch, int height, int width, int crop_size,
                            int rescaled_size, float* out_nhwc, float* scale);
int hmdpose_run_best_i420(hmdpose_t* h, const uint8_t* frame, int height, int width, int crop_size, int rescaled_size,
                          const float* cam6, float* out11, float* scale);

/*
 * Pose packet of the WebRTC "pose" data channel: the six floats the receiver sends after
 * post-processing, Program.cs:279-292 -- { rvec.x, rvec.y, rvec.z (axis-angle, rad), t.x, t.y, t.z (m) } copied with
 * Buffer.BlockCopy into 24 bytes (little-endian fp32), read back the same way by PoseDataChannel.cs:80-108.
 * hmdpose_pose_packet is host-only arithmetic on an out11 of hmdpose_run_best / hmdpose_best_from_raw;
 * hmdpose_run_packet = hmdpose_run_best + hmdpose_pose_packet (score is returned separately, may be NULL).
 */
#define HMDPOSE_PACKET_BYTES 24
int hmdpose_pose_packet(const float* out11, uint8_t* packet24);
int hmdpose_run_packet(hmdpose_t* h, const float* input_nchw, const float* cam6, uint8_t* packet24, float* score);

/*
 * EfficientDet-d0 detection variant (BASELINE.json configs[4]).  The handle holds an
 * EfficientDet checkpoint (backbone_net + bifpn + regressor + classifier, e.g. 90 COCO classes) or a full
 * HMDEgoPose one; only the backbone, BiFPN and the box / class sub-nets run.  Post-processing is the EfficientDet
 * one, not filter_detections:
 *   anchors      efficientdet/utils.py:76-139   (y1,x1,y2,x2), anchor_scale 4, ratios (1,1),(1.4,.7),(.7,1.4)
 *   decode+clip  efficientdet/utils.py:7-52     -> (x1,y1,x2,y2), x1,y1 >= 0, x2 <= S-1, y2 <= S-1
 *   postprocess  utils/utils.py:90-128          score = max_c, keep score > threshold, torchvision batched_nms
 *                (boxes offset by class_id * (max_coordinate + 1), IoU > iou_threshold suppresses), keep order
 * Outputs, HOST memory, max_out <= 4096 rows per frame: rois (B,max_out,4), class_ids (B,max_out), scores (B,max_out),
 * kept_anchor_idx (B,max_out), counts (B).  Rows >= |counts[b]| are -1.  Any output pointer may be NULL.
 * The reference keeps EVERY NMS survivor: when a frame has more survivors than max_out rows, counts[b] is NEGATIVE
 * (-counts[b] rows were written, in reference order, and more exist) -- never a silent truncation.
 */
int hmdpose_run_d0(hmdpose_t* h, const float* input_nchw, int batch, float threshold, float iou_threshold,
                   int max_out, float* rois, int32_t* class_ids, float* scores, int32_t* kept_anchor_idx,
                   int32_t* counts);
/* utils/utils.py:90-128 alone on host head tensors: regression (B,N,4) = dy,dx,dh,dw; classification (B,N,C). */
int hmdpose_d0_postprocess(hmdpose_t* h, const float* regression, const float* classification, int batch,
                           float threshold, float iou_threshold, int max_out, float* rois, int32_t* class_ids,
                           float* scores, int32_t* kept_anchor_idx, int32_t* counts);
/* EfficientDet anchors without a handle or a GPU (host-only): (N,4) y1,x1,y2,x2.  Returns N. */
int hmdpose_compute_anchors_d0(int image_size, float* anchors_yxyx_n4, int capacity_n);

/*
 * Device-resident variants used by the PyTorch-side wrapper (no host round trip; the reference
 * instead copies all five head tensors to the CPU, layers.py:448-452).  All pointers are DEVICE
 * pointers on the handle's device.  input strides are in ELEMENTS so the reference's permuted NHWC
 * view (eval/common.py:397) is consumed without a copy.  stream: a cudaStream_t cast to void*
 * (NULL = the handle's own stream); the call is asynchronous with respect to the host.
 */
int hmdpose_run_raw_device(hmdpose_t* h, const float* d_input, int64_t stride_b, int64_t stride_c,
                           int64_t stride_h, int64_t stride_w, int batch, float* d_regression,
                           float* d_classification, float* d_rotation, float* d_translation_raw,
                           float* d_hand, void* stream);
int hmdpose_run_detect_device(hmdpose_t* h, const float* d_input, int64_t stride_b, int64_t stride_c,
                              int64_t stride_h, int64_t stride_w, const float* d_cam6, int batch,
                              float* d_boxes, float* d_scores, int32_t* d_labels, float* d_rotation,
                              float* d_translation, float* d_hand, int32_t* d_kept_anchor_idx,
                              void* stream);

/* ---- run statistics (not part of the reference surface) ---- */
/* Number of kernels of this library launched by the last run_* call (graph nodes when replayed). */
int hmdpose_last_launch_count(const hmdpose_t* h);
/* Device time in ms of the last run_* call's GPU work (CUDA events on the stream the call used; waits for the
 * call if it is still in flight).  0 when the last call was a *_device call on a stream that has been destroyed. */
float hmdpose_last_gpu_ms(const hmdpose_t* h);
/* Test / profiling hooks (hmdpose_debug_read, hmdpose_profile_steps, hmdpose_test_gemm) are exported too but
 * declared in hmdpose_internal.h: they are not part of the drop-in surface. */
const char* hmdpose_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HMDPOSE_H_ */
