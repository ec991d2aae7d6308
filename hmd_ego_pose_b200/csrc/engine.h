// Host-side engine of libhmdpose: weight blob, anchors, buffer planning, launch plans, CUDA graphs.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hmdpose.h"
#include "../../include/hmdpose_internal.h"
#include "common.cuh"
#include "postprocess.h"

namespace hp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define HP_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw hp::Error(HMDPOSE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                          std::to_string(__LINE__) + ")");                              \
  } while (0)

// ---- weight blob (written by hmd_ego_pose_b200/packer.py) -------------------------------------
struct HostTensor {
  std::vector<int> dims;
  const float* data = nullptr;
  size_t count = 0;
};
struct WeightBlob {
  std::vector<uint8_t> storage;
  std::map<std::string, HostTensor> tensors;
  int num_classes = 1;
  void parse(const void* blob, size_t bytes);
  const HostTensor& get(const std::string& name) const;
  bool has(const std::string& name) const { return tensors.count(name) != 0; }
};

// ---- anchors (generators/utils/anchors.py:273-419), host arithmetic in double -----------------
int anchors_count(int S);
void compute_anchors(int S, std::vector<float>& boxes, std::vector<float>& tanchors);
// EfficientDet-native anchors (efficientdet/utils.py:76-139): (N,4) y1,x1,y2,x2
void compute_anchors_d0(int S, std::vector<float>& yxyx);

struct BlockSpec { int k, s, e, cin, cout; bool skip; };
extern const BlockSpec kB0Blocks[16];
void same_pad(int n, int k, int s, int* lo, int* hi);

struct Tens {
  void* p = nullptr;
  int H = 0, W = 0, C = 0;
  size_t elems(int b) const { return (size_t)b * H * W * C; }
};

struct Step {
  std::string name;
  std::function<void(cudaStream_t)> launch;
  const char* kernel = "";  // kernel function this step launches
  double bytes = 0;         // algorithmic (compulsory) HBM bytes of the step as fused
  double flops = 0;         // 2 * MACs
};

enum PlanMode { PLAN_RAW = 0, PLAN_DET = 1, PLAN_BEST = 2, PLAN_D0 = 4 };  // D0: box + class heads, EfficientDet-native post-processing

struct Plan {
  int b = 0;
  int mode = 0;
  std::vector<Step> steps;
  std::vector<void*> owned;  // device tables owned by this plan
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
  ~Plan();
};

class Engine {
 public:
  Engine(const hmdpose_config_t& cfg, const void* blob, size_t bytes);
  ~Engine();
  void release();   // frees every device / host resource (also used when the constructor throws)

  // network + post-processing over `batch` frames already on the device
  void run_device(const float* d_in, long long sb, long long sc, long long sh, long long sw, const float* d_cam,
                  int batch, bool want_raw, float* raw[5], bool want_det, float* d_boxes, float* d_scores,
                  int32_t* d_labels, float* d_rot, float* d_trans, float* d_hand, int32_t* d_idx, bool want_best,
                  float* d_best, cudaStream_t st);
  void run_raw_host(const float* in, int batch, float* outs[5]);
  void run_detect_host(const float* in, const float* cam, int batch, float* boxes, float* scores, int32_t* labels,
                       float* rot, float* trans, float* hand, int32_t* idx);
  void run_best_host(const float* in, const float* cam, float* out11);
  void postprocess_host(const float* reg, const float* cls, const float* rot, const float* traw, const float* hand,
                        const float* cam, const float* boxes_in, const float* trans_in, int batch, float* boxes,
                        float* scores, int32_t* labels, float* rot_o, float* trans_o, float* hand_o, int32_t* idx);
  void best_from_raw_host(const float* reg, const float* cls, const float* rot, const float* traw, const float* cam,
                          float* out11);
  // EfficientDet-d0 detection variant (utils/utils.py:90-128): backbone + BiFPN + box/class heads + class-offset NMS
  void run_d0_host(const float* in, int batch, float thr, float iou, int max_out, float* rois, int32_t* class_ids,
                   float* scores, int32_t* idx, int32_t* counts);
  void d0_postprocess_host(const float* reg, const float* cls, int batch, float thr, float iou, int max_out, float* rois,
                           int32_t* class_ids, float* scores, int32_t* idx, int32_t* counts);
  // uint8 RGB frames (all h x w) -> pre-processed network input (colibri_common.py:622-656) and the paths behind it
  void preprocess_host(const uint8_t* imgs, int batch, int h, int w, float* out_nhwc, float* scale);
  void run_detect_u8_host(const uint8_t* imgs, int batch, int h, int w, const float* cam, float* boxes, float* scores,
                          int32_t* labels, float* rot, float* trans, float* hand, int32_t* idx, float* scale);
  void run_best_u8_host(const uint8_t* img, int h, int w, const float* cam, float* out11, float* scale);
  // I420 frames through the C# receiver's frame path (Program.cs:137-200): YV12-trick BGR, centre crop, rescale, normalise
  void preprocess_i420_host(const uint8_t* frames, int batch, int h, int w, int crop, int mid, float* out_nhwc, float* scale);
  void run_best_i420_host(const uint8_t* frame, int h, int w, int crop, int mid, const float* cam, float* out11, float* scale);
  long long debug_read(const std::string& name, float* out, long long cap);
  // per-step device times (CUDA events on the handle's stream, un-graphed), averaged over reps
  int profile_steps(int batch, int mode, int reps, char* names, char* kernels, float* ms, double* bytes, double* flops,
                    int capacity);

  hmdpose_config_t cfg;
  int N = 0;  // anchors
  std::vector<float> h_anchors, h_tanchors;
  std::string last_error;
  std::mutex mu;
  int last_launches = 0;
  float last_ms = 0.f;
  float last_gpu_ms();          // device time of the last run_* call (waits for it if it is still in flight)
  cudaStream_t stream = nullptr;

 private:
  template <typename T> void upload_weights();
  template <typename T> void alloc_buffers();
  template <typename T> Plan* get_plan(int b, int mode);
  template <typename T> std::unique_ptr<Plan> build_plan(int b, int mode);
  template <typename T> Step stem_step(const float* d_in, long long sb, long long sc, long long sh, long long sw, int b);
  void run_plan(Plan* p, cudaStream_t st);
  // Cross-stream ordering of a handle's internal buffers: every entry point calls enter(st) before it enqueues work
  // that touches them and leave(st) after its last enqueue.  When consecutive calls use different streams (a torch
  // stream through the *_device entry points, the handle's own stream through the host API) the later stream waits
  // for the earlier call's completion event, so mixing the two on one handle is ordered instead of racing.
  void enter(cudaStream_t st);
  void leave(cudaStream_t st);
  void wait_stream();   // host wait on the handle's stream (spin, or sleep with HMDPOSE_BLOCKING_SYNC=1)
  void* dalloc(size_t bytes);
  const void* w9_for(const std::string& dw_name, const std::string& pw_name);
  float* upload_f32(const float* src, size_t n);
  template <typename T> void* upload_as(const float* src, size_t n);
  void reg_debug(const std::string& name, const Tens& t, bool is_t = true);
  void ensure_host_staging(int batch, bool need_raw = false);
  void add_post_steps(std::vector<Step>& steps, int b, int mode, bool decode_boxes, bool decode_trans, bool hand_from_raw,
                      bool pose_from_gather = false);
  void ensure_d0(float thr, float iou);
  D0Args d0_args() const;
  void d0_download(int f0, int b, uint8_t* h_out);
  void d0_scatter(const uint8_t* h_out, int batch, int max_out, float* rois, int32_t* class_ids, float* scores,
                  int32_t* idx, int32_t* counts);
  // single-class detection plans evaluate the rotation / translation / hand HEADERS only at the kept anchors
  // (pose_gather_kernel); HMDPOSE_DENSE_POSE=1 keeps the dense rotation / translation headers
  bool gather_pose_for(int mode) const {
    return mode == PLAN_DET && !full_hand_for(mode) && cfg.num_classes == 1 && !post_v1_ && !dense_pose_;
  }
  bool full_hand_for(int mode) const { return mode == PLAN_RAW || gather_hand_off_ || (iter1_ && (mode & PLAN_DET)); }

  WeightBlob blob_;
  bool fast_ = false;
  int mb_ = 1;  // micro-batch (frames per internal pass)
  std::vector<void*> allocs_;
  std::map<std::string, void*> wdev_;          // device weights by name (fp32 unless suffixed ".T")
  std::map<std::string, std::pair<Tens, bool>> debug_;  // name -> (tensor, stored as T?)
  int last_b_ = 0;
  std::map<int, std::unique_ptr<Plan>> plans_;  // key = b * 8 + mode
  std::map<int, std::unique_ptr<Plan>> post_plans_;  // post-processing only (hmdpose_postprocess & co)

  // activations (sized for mb_)
  Tens stem_out_;
  struct BlockBufs { Tens exp, dw, out; float* se_partial = nullptr; float* gate = nullptr; int tiles = 0;
                     void* wgated = nullptr; };   // wgated: per-image project weights with the gate folded in (fast mode)
  BlockBufs blk_[16];
  void* scratch_exp_ = nullptr; void* scratch_dw_ = nullptr;
  struct CellBufs { Tens in[5], in2[2], p6_pre, up[5], out[5], fused[5], dwb[5]; };
  CellBufs cell_[3];
  Tens trunk_[5][5][3], hdw_[5][5], hdrdw_[6][5];   // trunk: ping-pong [0]/[1]; HMDPOSE_KEEP_ALL keeps layer i in [i]
  int trunk_slot(int layer) const { return keep_all_ ? layer : (layer & 1); }
  int trunk_final() const { return keep_all_ ? 2 : 0; }   // where the output of the third trunk layer sits
  // --iter 1 refinement sub-nets (rotation, translation, hand): concat input, its depthwise output, the 64-channel
  // refinement feature, and (non-fused paths) the depthwise output of the refinement heads
  bool iter1_ = false;
  Tens it_in_[3][5], it_dw_[3][5], it_y_[3][5], it_hdw_[4][5];
  int lvl_hw_[5], lvl_side_[5], lvl_off_[5];
  // micro-batch-local head outputs + post-processing buffers
  float *o_reg_ = nullptr, *o_cls_ = nullptr, *o_rot_ = nullptr, *o_traw_ = nullptr, *o_hand_ = nullptr;
  float *p_boxes_ = nullptr, *p_trans_ = nullptr;
  PostBuffers pb_;
  float *det_boxes_ = nullptr, *det_scores_ = nullptr, *det_rot_ = nullptr, *det_trans_ = nullptr, *det_hand_ = nullptr;
  int32_t *det_labels_ = nullptr, *det_idx_ = nullptr;
  float* d_best_ = nullptr;
  uint8_t* d_u8_ = nullptr; size_t d_u8_bytes_ = 0;   // device copy of the uint8 frames (pre-processing entry points)
  float stage_u8(const uint8_t* imgs, int batch, int h, int w);
  float stage_i420(const uint8_t* frames, int batch, int h, int w, int crop, int mid);   // H2D + preprocess_kernel into d_in_stage_ (NHWC)
  StemW stem_wb_;   // stem weights + bias as passed to stem_kernel (kernel parameter)
  float* mb_part_ = nullptr; size_t mb_part_bytes_ = 0;   // split-K scratch of the fused MBConv kernel
  int* se_counters_ = nullptr;   // [16 blocks][mb]: dw3 blocks finished per image (squeeze-excite folded into dw3)
  // D0 variant
  int num_heads_ = 5;  // 2 for a detector-only blob (regressor + classifier)
  float d0_thr_ = -1.f, d0_iou_ = -1.f;
  float* d_anchors_d0_ = nullptr;
  int *d0_cand_cls_ = nullptr, *d0_count_ = nullptr, *d0_ocls_ = nullptr, *d0_oidx_ = nullptr, *d0_ocount_ = nullptr;
  float *d0_orois_ = nullptr, *d0_oscores_ = nullptr, *d0_sel_ = nullptr;
  float* d_cam_local_ = nullptr;  // [mb][6] camera rows of the current micro-batch (fixed address for graphs)
  float *d_anchors_ = nullptr, *d_tanchors_ = nullptr;
  // full-batch staging for the host API
  float* d_in_stage_ = nullptr; float* d_cam_stage_ = nullptr;
  float *d_full_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float *df_boxes_ = nullptr, *df_scores_ = nullptr, *df_rot_ = nullptr, *df_trans_ = nullptr, *df_hand_ = nullptr;
  int32_t *df_labels_ = nullptr, *df_idx_ = nullptr;
  uint8_t* h_pinned_ = nullptr; size_t h_pinned_bytes_ = 0;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, ev_block_ = nullptr, ev_done_ = nullptr;
  cudaStream_t last_stream_ = nullptr;
  bool has_last_ = false;
  bool timing_valid_ = false;   // ev0_/ev1_ bracket the last run_* call
  bool dense_pose_ = false;
  bool wgate_ = true;     // squeeze-excite gate folded into per-image project weights by se3_kernel (HMDPOSE_NO_WGATE=1: off)
  bool projk_ = true;     // split-K cluster GEMM for the deep-K project convolutions of the small maps (HMDPOSE_NO_PROJK=1: off)
  int projk_max_batch_ = 4;   // ... on plans of at most this many frames (HMDPOSE_PROJK_MAX_BATCH)
  bool expdw_ = false;    // HMDPOSE_EXPDW at create time: fused expand + depthwise kernel for the large maps (opt-in)
  bool mbfuse_ = false;   // HMDPOSE_MBFUSE at create time: fused MBConv cluster kernel for the small maps
  bool keep_all_ = false, force_simt_ = false, v1_ = false, gather_hand_off_ = false, post_v1_ = false;
};

// device watchdog record (tma.cuh): one host-mapped TrapInfo per process, installed into every translation unit that
// waits on mbarriers, for the current device.  trap_info_describe() is "" unless a kernel recorded a time-out.
void trap_info_setup();
void trap_info_install_gemm_tu(void* mapped);   // engine_gemm.cu's copy of the device pointer
std::string trap_info_describe();

// pointwise GEMM dispatch (engine_gemm.cu)
void init_gemm_kernels();
// builds a device table for the problems (tcgen05 path encodes the TMA descriptors) and returns a launcher
std::function<void(cudaStream_t)> make_gemm_launcher(std::vector<GemmProb> probs, bool fast, bool force_simt,
                                                     std::vector<void*>& owned, const char** kernel_name = nullptr,
                                                     bool v1 = false);
int gemm_choose_bn(int N, int* n_tiles, int cap = 128);
// fused expand + depthwise of the large maps (expdw_tc.cuh); empty when the block does not fit the kernel
std::function<void(cudaStream_t)> make_expdw_launcher(EdSpec sp, std::vector<void*>& owned, int* tiles_per_img);
// split-K project GEMM of the small maps (projk_tc.cuh); empty when the problem does not fit the kernel
bool projk_fits(PkSpec sp, size_t part_bytes);
std::function<void(cudaStream_t)> make_projk_launcher(PkSpec sp, const void* w, std::vector<void*>& owned, float* part,
                                                      size_t part_bytes);
// fused MBConv block for small maps (mbconv_tc.cuh); empty when the block does not fit the kernel
std::function<void(cudaStream_t)> make_mbconv_launcher(MbSpec sp, int batch, std::vector<void*>& owned, float* part,
                                                       size_t part_bytes);
int sep3_debug_timeline(float* out, int cap);
int mb_debug_timeline(float* out, int cap);
void encode_act_4d(CUtensorMap* tm, const void* base, bool is_half, int C, int W, int H, int B, int box_c, int box_w,
                   int box_h, bool swizzle128 = false);
// fused depthwise-separable conv on the 64-channel pyramid (fast mode)
struct SepSpec {
  GemmProb p;  // bias, W (pointwise weights, fp16 [N][64]), out, N, ldo, act, out_mode & head mapping
  const void* in = nullptr; const void* fb = nullptr; const void* fc = nullptr;
  const float* dw_w = nullptr;
  const float* scale = nullptr;  // per-output-channel scale on the accumulator (per-level BN of a shared head conv)
  const void* w9 = nullptr;  // folded tap matrices [9][N][64] fp16 (W_pw . diag(w_dw[:, tap])), implicit-GEMM path
  int H = 0, W = 0, Bn = 0, fused = 0, mode_b = 0, mode_c = 0;
  float w0 = 0, w1 = 0, w2 = 0;
};
std::function<void(cudaStream_t)> make_sepconv_launcher(std::vector<SepSpec> specs, std::vector<void*>& owned,
                                                        const char** kernel_name = nullptr);
// several small-level BiFPN nodes in ONE launch: CTA g runs them in order for images [g*nb, (g+1)*nb)
std::function<void(cudaStream_t)> make_sepconv_chain_launcher(std::vector<SepSpec> specs, int nb, std::vector<void*>& owned);

}  // namespace hp
