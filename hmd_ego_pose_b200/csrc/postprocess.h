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
