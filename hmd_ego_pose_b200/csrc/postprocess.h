// Host-side launchers of the post-processing kernels (postprocess.cu, compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hp {

constexpr int MAX_DET_CAP = 256;  // upper bound accepted for max_detections (reference: 100)

struct PostBuffers {
  unsigned long long* keys = nullptr;  // [B*C][cap] sort keys: (~sortable(score) << 32) | anchor
  int* kept_idx = nullptr;             // [B*C][max_det]
  float* kept_score = nullptr;         // [B*C][max_det]
  int* kept_count = nullptr;           // [B*C]
  int cap = 0;                         // power of two >= N
};

// Single-class fast path: score threshold + compaction + on-the-fly box decode of the candidates + NMS +
// translation recovery + gather/pad in ONE kernel (one block per image).
struct FilterArgs {
  const float* boxes;        // (B,N,4) already decoded, or null: decode from anchors + regression
  const float* anchors; const float* reg; float wmax, hmax;
  const float* scores;       // (B,N,1)
  const float* rotation;     // (B,N,3)
  const float* translation;  // (B,N,3) already decoded, or null: decode from tanchors + traw + cam
  const float* tanchors; const float* traw; const float* cam;
  const float* hand;         // (B,N,H) or null (hand rows are produced elsewhere)
  int N, H, cap, max_det;
  float score_thr, iou_thr;
  unsigned long long* keys;  // [B][cap] scratch
  float* o_boxes; float* o_scores; int* o_labels; float* o_rot; float* o_trans; float* o_hand; int* o_idx;
};
void launch_filter_fused(const FilterArgs& a, int B, cudaStream_t st);

// EfficientDet-d0 detection variant (efficientdet/utils.py:7-139, utils/utils.py:90-128)
constexpr int D0_MAX_OUT = 4096;  // device capacity of detections per image (the reference keeps all NMS survivors;
                                  // a frame with more is flagged: o_count = -kept)
constexpr int D0_SEL_SMEM = 512;  // selected boxes held in shared memory; later ones live in sel_scratch
struct D0Args {
  const float* anchors_yxyx;  // (N,4) y1,x1,y2,x2
  const float* reg;           // (B,N,4) dy,dx,dh,dw
  const float* cls;           // (B,N,C) after sigmoid
  int N, C, cap, max_out;
  float wmax, hmax, threshold, iou_thr;
  unsigned long long* keys;   // [B][cap] scratch
  int* cand_cls;              // [B][N] arg-max class of every anchor above the threshold
  int* cand_count;            // [B]
  float* box_scratch;         // [B][N][4] offset boxes when a candidate set does not fit shared memory
  float* o_rois; int* o_cls; float* o_scores; int* o_idx; int* o_count;   // [B][max_out][...], [B]
  float* sel_scratch;   // [B][max_out] float4: class-offset boxes of the selections beyond D0_SEL_SMEM
};
void launch_d0(const D0Args& a, int B, cudaStream_t st);
// the same when the candidate keys / classes / counts were already produced by the classifier header's epilogue
// (sepconv_kernel out_mode 2): only the per-image sort + decode + class-offset NMS
void launch_d0_nms(const D0Args& a, int B, cudaStream_t st);

// Frame pre-processing (generators/colibri_common.py:622-656; C# twin Program.cs:397-445), SURVEY.md 8f-1
struct PreArgs {
  const uint8_t* img;   // [B][h][w][3] uint8 RGB
  float* out;           // [B][S][S][3] float32 NHWC (the layout eval/common.py:397 permutes into an NCHW view)
  int B, h, w, rh, rw, S;
};
void launch_preprocess(const PreArgs& a, cudaStream_t st);
// the C# receiver's frame path (WebRTCNetCoreSandbox/Program.cs:137-200, 381-445): I420 frame -> YV12-trick BGR ->
// centre crop -> rescale -> ResizeAndNormalizeMat, one kernel
struct I420Args {
  const uint8_t* img;   // [B][h*w*3/2] I420 frames (Y, U, V planes; h, w even)
  float* out;           // [B][S][S][3] float32 NHWC in the Mat's channel order
  int B, h, w, crop, mid, rh, rw, S;
};
void launch_preprocess_i420(const I420Args& a, cudaStream_t st);

void launch_decode_boxes(const float* anchors, const float* reg, int B, int N, int width, int height, float* boxes,
                         cudaStream_t st);
void launch_decode_translation(const float* tanchors, const float* raw, const float* cam, int B, int N, float* out,
                               cudaStream_t st);
void launch_filter_nms(const PostBuffers& pb, const float* boxes, const float* scores, int B, int N, int C,
                       float score_thr, float iou_thr, int max_det, cudaStream_t st);
void launch_topk_gather(const PostBuffers& pb, const float* boxes, const float* rotation, const float* translation,
                        const float* hand, int B, int N, int C, int H, int max_det, float* o_boxes, float* o_scores,
                        int* o_labels, float* o_rot, float* o_trans, float* o_hand, int* o_idx, cudaStream_t st);
void launch_best(const float* anchors, const float* tanchors, const float* reg, const float* scores, const float* rot,
                 const float* traw, const float* cam, int B, int N, int C, float score_thr, int width, int height,
                 float* out11, cudaStream_t st);

}  // namespace hp
