// extern "C" boundary of libhmdpose.so (declared in include/hmdpose.h).  Nothing throws across it.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>

#include "engine.h"

struct hmdpose {
  hp::Engine* eng = nullptr;
};

static thread_local std::string g_create_error;

template <typename F>
static int guarded(hmdpose_t* h, F&& f) {
  if (!h || !h->eng) return HMDPOSE_E_ARG;
  std::lock_guard<std::mutex> lock(h->eng->mu);
  try {
    f(*h->eng);
    return HMDPOSE_OK;
  } catch (const hp::Error& e) {
    h->eng->last_error = e.what() + (e.code == HMDPOSE_E_CUDA ? hp::trap_info_describe() : std::string());
    return e.code;
  } catch (const std::exception& e) {
    h->eng->last_error = e.what();
    return HMDPOSE_E_STATE;
  }
}

extern "C" {

const char* hmdpose_version(void) { return "hmdpose-b200 0.1 (sm_100a)"; }

void hmdpose_default_config(hmdpose_config_t* cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = HMDPOSE_ABI_VERSION;
  cfg->image_size = 256;
  cfg->max_batch = 1;
  cfg->device = 0;
  cfg->precision = HMDPOSE_PRECISION_FAST;
  cfg->num_classes = 0;          // take it from the weight blob
  cfg->score_threshold = 0.5f;   // train.py:80
  cfg->iou_threshold = 0.5f;     // layers.py:414
  cfg->max_detections = 100;     // train.py:81
  cfg->micro_batch = 0;
  cfg->use_graph = 1;
}

int hmdpose_create_from_memory(const hmdpose_config_t* cfg, const void* blob, size_t blob_bytes, hmdpose_t** out) {
  if (!cfg || !out || cfg->abi_version != HMDPOSE_ABI_VERSION) {
    g_create_error = "bad config / ABI version";
    return HMDPOSE_E_ARG;
  }
  *out = nullptr;
  try {
    std::unique_ptr<hp::Engine> eng(new hp::Engine(*cfg, blob, blob_bytes));   // a throwing constructor frees its own device memory
    std::unique_ptr<hmdpose> h(new hmdpose());
    h->eng = eng.release();
    *out = h.release();
    return HMDPOSE_OK;
  } catch (const hp::Error& e) {
    g_create_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return HMDPOSE_E_STATE;
  } catch (...) {
    g_create_error = "unknown error";
    return HMDPOSE_E_STATE;
  }
}

int hmdpose_create_ex(const hmdpose_config_t* cfg, const char* weights_path, hmdpose_t** out) {
  if (!weights_path) { g_create_error = "null weights path"; return HMDPOSE_E_ARG; }
  try {   // nothing may throw across the C boundary (directories / unseekable paths report tellg() == -1)
    std::ifstream f(weights_path, std::ios::binary | std::ios::ate);
    if (!f) { g_create_error = std::string("cannot open ") + weights_path; return HMDPOSE_E_WEIGHTS; }
    const std::streamoff n = f.tellg();
    if (n <= 0 || n > (std::streamoff)1 << 32) {
      g_create_error = std::string("not a weight blob (unreadable or empty): ") + weights_path;
      return HMDPOSE_E_WEIGHTS;
    }
    f.seekg(0);
    std::vector<char> buf((size_t)n);
    if (!f.read(buf.data(), n)) { g_create_error = "short read on weight blob"; return HMDPOSE_E_WEIGHTS; }
    return hmdpose_create_from_memory(cfg, buf.data(), (size_t)n, out);
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return HMDPOSE_E_WEIGHTS;
  } catch (...) {
    g_create_error = "unknown error while reading the weight blob";
    return HMDPOSE_E_WEIGHTS;
  }
}

int hmdpose_create(const char* weights_path, int image_size, int max_batch, int device, float score_threshold,
                   float iou_threshold, int max_detections, hmdpose_t** out) {
  hmdpose_config_t cfg;
  hmdpose_default_config(&cfg);
  cfg.image_size = image_size; cfg.max_batch = max_batch; cfg.device = device;
  cfg.score_threshold = score_threshold; cfg.iou_threshold = iou_threshold; cfg.max_detections = max_detections;
  return hmdpose_create_ex(&cfg, weights_path, out);
}

void hmdpose_destroy(hmdpose_t* h) {
  if (!h) return;
  delete h->eng;
  delete h;
}

const char* hmdpose_last_error(const hmdpose_t* h) {
  if (h && h->eng) return h->eng->last_error.c_str();
  return g_create_error.c_str();
}

int hmdpose_num_anchors(const hmdpose_t* h) { return (h && h->eng) ? h->eng->N : HMDPOSE_E_ARG; }
int hmdpose_num_classes(const hmdpose_t* h) { return (h && h->eng) ? h->eng->cfg.num_classes : HMDPOSE_E_ARG; }

int hmdpose_get_anchors(const hmdpose_t* h, float* anchors_n4, float* translation_anchors_n3) {
  if (!h || !h->eng) return HMDPOSE_E_ARG;
  if (anchors_n4) std::memcpy(anchors_n4, h->eng->h_anchors.data(), h->eng->h_anchors.size() * 4);
  if (translation_anchors_n3) std::memcpy(translation_anchors_n3, h->eng->h_tanchors.data(), h->eng->h_tanchors.size() * 4);
  return HMDPOSE_OK;
}

int hmdpose_compute_anchors(int image_size, float* anchors_n4, float* translation_anchors_n3, int capacity_n) {
  if (image_size < 8) return HMDPOSE_E_ARG;
  const int n = hp::anchors_count(image_size);
  if (!anchors_n4 && !translation_anchors_n3) return n;
  if (capacity_n < n) return HMDPOSE_E_ARG;
  std::vector<float> a, t;
  hp::compute_anchors(image_size, a, t);
  if (anchors_n4) std::memcpy(anchors_n4, a.data(), a.size() * 4);
  if (translation_anchors_n3) std::memcpy(translation_anchors_n3, t.data(), t.size() * 4);
  return n;
}

int hmdpose_preprocess(hmdpose_t* h, const uint8_t* images, int batch, int height, int width, float* out_nhwc,
                       float* scale) {
  return guarded(h, [&](hp::Engine& e) { e.preprocess_host(images, batch, height, width, out_nhwc, scale); });
}

int hmdpose_run_detect_u8(hmdpose_t* h, const uint8_t* images, int batch, int height, int width, const float* cam6,
                          float* boxes, float* scores, int32_t* labels, float* rotation, float* translation, float* hand,
                          int32_t* kept_anchor_idx, float* scale) {
  return guarded(h, [&](hp::Engine& e) {
    e.run_detect_u8_host(images, batch, height, width, cam6, boxes, scores, labels, rotation, translation, hand,
                         kept_anchor_idx, scale);
  });
}

int hmdpose_preprocess_i420(hmdpose_t* h, const uint8_t* frames, int batch, int height, int width, int crop_size,
                            int rescaled_size, float* out_nhwc, float* scale) {
  return guarded(h, [&](hp::Engine& e) { e.preprocess_i420_host(frames, batch, height, width, crop_size, rescaled_size, out_nhwc, scale); });
}

int hmdpose_run_best_i420(hmdpose_t* h, const uint8_t* frame, int height, int width, int crop_size, int rescaled_size,
                          const float* cam6, float* out11, float* scale) {
  return guarded(h, [&](hp::Engine& e) { e.run_best_i420_host(frame, height, width, crop_size, rescaled_size, cam6, out11, scale); });
}

int hmdpose_run_best_u8(hmdpose_t* h, const uint8_t* image, int height, int width, const float* cam6, float* out11,
                        float* scale) {
  return guarded(h, [&](hp::Engine& e) { e.run_best_u8_host(image, height, width, cam6, out11, scale); });
}

int hmdpose_pose_packet(const float* out11, uint8_t* packet24) {
  if (!out11 || !packet24) return HMDPOSE_E_ARG;
  for (int i = 0; i < 6; ++i) {   // little-endian fp32, independent of the host byte order
    uint32_t u;
    std::memcpy(&u, out11 + 5 + i, 4);
    packet24[4 * i + 0] = (uint8_t)(u & 0xff);
    packet24[4 * i + 1] = (uint8_t)((u >> 8) & 0xff);
    packet24[4 * i + 2] = (uint8_t)((u >> 16) & 0xff);
    packet24[4 * i + 3] = (uint8_t)((u >> 24) & 0xff);
  }
  return 0;
}

int hmdpose_run_packet(hmdpose_t* h, const float* input_nchw, const float* cam6, uint8_t* packet24, float* score) {
  float out11[HMDPOSE_BEST_LEN];
  const int rc = hmdpose_run_best(h, input_nchw, cam6, out11);
  if (rc != 0) return rc;
  if (score) *score = out11[0];
  return hmdpose_pose_packet(out11, packet24);
}

int hmdpose_compute_anchors_d0(int image_size, float* anchors_yxyx_n4, int capacity_n) {
  if (image_size < 8) return HMDPOSE_E_ARG;
  std::vector<float> a;
  hp::compute_anchors_d0(image_size, a);
  const int n = (int)(a.size() / 4);
  if (!anchors_yxyx_n4) return n;
  if (capacity_n < n) return HMDPOSE_E_ARG;
  std::memcpy(anchors_yxyx_n4, a.data(), a.size() * 4);
  return n;
}

int hmdpose_run_d0(hmdpose_t* h, const float* input_nchw, int batch, float threshold, float iou_threshold, int max_out,
                   float* rois, int32_t* class_ids, float* scores, int32_t* kept_anchor_idx, int32_t* counts) {
  return guarded(h, [&](hp::Engine& e) {
    e.run_d0_host(input_nchw, batch, threshold, iou_threshold, max_out, rois, class_ids, scores, kept_anchor_idx, counts);
  });
}

int hmdpose_d0_postprocess(hmdpose_t* h, const float* regression, const float* classification, int batch,
                           float threshold, float iou_threshold, int max_out, float* rois, int32_t* class_ids,
                           float* scores, int32_t* kept_anchor_idx, int32_t* counts) {
  return guarded(h, [&](hp::Engine& e) {
    e.d0_postprocess_host(regression, classification, batch, threshold, iou_threshold, max_out, rois, class_ids, scores,
                          kept_anchor_idx, counts);
  });
}

int hmdpose_run_raw(hmdpose_t* h, const float* input_nchw, int batch, float* regression, float* classification,
                    float* rotation, float* translation_raw, float* hand) {
  return guarded(h, [&](hp::Engine& e) {
    float* outs[5] = {regression, classification, rotation, translation_raw, hand};
    e.run_raw_host(input_nchw, batch, outs);
  });
}

int hmdpose_run_detect(hmdpose_t* h, const float* input_nchw, const float* cam6, int batch, float* boxes, float* scores,
                       int32_t* labels, float* rotation, float* translation, float* hand, int32_t* kept_anchor_idx) {
  return guarded(h, [&](hp::Engine& e) {
    e.run_detect_host(input_nchw, cam6, batch, boxes, scores, labels, rotation, translation, hand, kept_anchor_idx);
  });
}

int hmdpose_run_best(hmdpose_t* h, const float* input_nchw, const float* cam6, float* out11) {
  return guarded(h, [&](hp::Engine& e) { e.run_best_host(input_nchw, cam6, out11); });
}

int hmdpose_postprocess(hmdpose_t* h, const float* regression, const float* classification, const float* rotation,
                        const float* translation_raw, const float* hand, const float* cam6, int batch, float* boxes,
                        float* scores, int32_t* labels, float* rotation_out, float* translation_out, float* hand_out,
                        int32_t* kept_anchor_idx) {
  return guarded(h, [&](hp::Engine& e) {
    e.postprocess_host(regression, classification, rotation, translation_raw, hand, cam6, nullptr, nullptr, batch, boxes,
                       scores, labels, rotation_out, translation_out, hand_out, kept_anchor_idx);
  });
}

int hmdpose_filter_boxes(hmdpose_t* h, const float* boxes_in, const float* classification, const float* rotation,
                         const float* translation, const float* hand, int batch, float* boxes, float* scores,
                         int32_t* labels, float* rotation_out, float* translation_out, float* hand_out,
                         int32_t* kept_anchor_idx) {
  return guarded(h, [&](hp::Engine& e) {
    if (!boxes_in || !translation) throw hp::Error(HMDPOSE_E_ARG, "null boxes / translation");
    e.postprocess_host(nullptr, classification, rotation, nullptr, hand, nullptr, boxes_in, translation, batch, boxes,
                       scores, labels, rotation_out, translation_out, hand_out, kept_anchor_idx);
  });
}

int hmdpose_best_from_raw(hmdpose_t* h, const float* regression, const float* classification, const float* rotation,
                          const float* translation_raw, const float* cam6, float* out11) {
  return guarded(h, [&](hp::Engine& e) {
    e.best_from_raw_host(regression, classification, rotation, translation_raw, cam6, out11);
  });
}

int hmdpose_run_raw_device(hmdpose_t* h, const float* d_input, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                           int64_t stride_w, int batch, float* d_regression, float* d_classification,
                           float* d_rotation, float* d_translation_raw, float* d_hand, void* stream) {
  return guarded(h, [&](hp::Engine& e) {
    float* raw[5] = {d_regression, d_classification, d_rotation, d_translation_raw, d_hand};
    e.run_device(d_input, stride_b, stride_c, stride_h, stride_w, nullptr, batch, true, raw, false, nullptr, nullptr,
                 nullptr, nullptr, nullptr, nullptr, nullptr, false, nullptr, (cudaStream_t)stream);
  });
}

int hmdpose_run_detect_device(hmdpose_t* h, const float* d_input, int64_t stride_b, int64_t stride_c, int64_t stride_h,
                              int64_t stride_w, const float* d_cam6, int batch, float* d_boxes, float* d_scores,
                              int32_t* d_labels, float* d_rotation, float* d_translation, float* d_hand,
                              int32_t* d_kept_anchor_idx, void* stream) {
  return guarded(h, [&](hp::Engine& e) {
    e.run_device(d_input, stride_b, stride_c, stride_h, stride_w, d_cam6, batch, false, nullptr, true, d_boxes, d_scores,
                 d_labels, d_rotation, d_translation, d_hand, d_kept_anchor_idx, false, nullptr, (cudaStream_t)stream);
  });
}

int64_t hmdpose_debug_read(hmdpose_t* h, const char* name, float* out, int64_t capacity) {
  long long n = 0;
  const int rc = guarded(h, [&](hp::Engine& e) { n = e.debug_read(name ? name : "", out, capacity); });
  return rc == HMDPOSE_OK ? (int64_t)n : (int64_t)rc;
}

int hmdpose_profile_steps(hmdpose_t* h, int batch, int mode, int reps, char* names, char* kernels, float* ms,
                          double* bytes, double* flops, int capacity) {
  int n = 0;
  const int rc = guarded(h, [&](hp::Engine& e) { n = e.profile_steps(batch, mode, reps, names, kernels, ms, bytes, flops, capacity); });
  return rc == HMDPOSE_OK ? n : rc;
}

int hmdpose_last_launch_count(const hmdpose_t* h) { return (h && h->eng) ? h->eng->last_launches : HMDPOSE_E_ARG; }

float hmdpose_last_gpu_ms(const hmdpose_t* h) {
  if (!h || !h->eng) return -1.f;
  std::lock_guard<std::mutex> lock(h->eng->mu);
  return h->eng->last_gpu_ms();   // also valid after the asynchronous *_device entry points (waits for their events)
}

// Standalone pointwise-GEMM check (see hmdpose.h).
int hmdpose_test_gemm(int device, int impl, int precision, int M, int N, int K, const float* A, const float* W,
                      const float* bias, const float* a_scale, int rows_per_img, const float* residual, int act,
                      float* D, float* gpu_ms) {
  using namespace hp;
  std::vector<void*> owned;
  int rc = HMDPOSE_OK;
  try {
    if (!A || !W || !bias || !D || M < 1 || N < 1 || K < 1) throw Error(HMDPOSE_E_ARG, "bad gemm arguments");
    const bool fast = precision == HMDPOSE_PRECISION_FAST;
    if ((impl == 1 || impl == 2) && !fast) throw Error(HMDPOSE_E_ARG, "the kind::f16 tcgen05 GEMM runs in fast mode only");
    if (impl == 3 && fast) throw Error(HMDPOSE_E_ARG, "the 3xTF32 tcgen05 GEMM runs in parity mode only");
    HP_CUDA(cudaSetDevice(device));
    auto up = [&](const float* src, size_t n, bool as_half) -> void* {
      void* d = nullptr;
      if (as_half) {
        std::vector<__half> hbuf(n);
        for (size_t i = 0; i < n; ++i) hbuf[i] = __float2half_rn(src[i]);
        HP_CUDA(cudaMalloc(&d, n * 2));
        HP_CUDA(cudaMemcpy(d, hbuf.data(), n * 2, cudaMemcpyHostToDevice));
      } else {
        HP_CUDA(cudaMalloc(&d, n * 4));
        HP_CUDA(cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice));
      }
      owned.push_back(d);
      return d;
    };
    GemmProb p;
    std::memset(&p, 0, sizeof(p));
    p.A = up(A, (size_t)M * K, fast);
    p.W = up(W, (size_t)N * K, fast);
    p.bias = (const float*)up(bias, (size_t)N, false);
    const int rpi = rows_per_img > 0 ? rows_per_img : M;
    if (a_scale) p.a_scale = (const float*)up(a_scale, (size_t)cdiv(M, rpi) * K, false);
    if (residual) p.residual = up(residual, (size_t)M * N, fast);
    void* dout = nullptr;
    HP_CUDA(cudaMalloc(&dout, (size_t)M * N * (fast ? 2 : 4)));
    owned.push_back(dout);
    p.out = dout; p.M = M; p.N = N; p.K = K; p.lda = K; p.ldo = N; p.act = act; p.rows_per_img = rpi;
    p.p_src = 1; p.p_dst = 1;
    auto launch = make_gemm_launcher({p}, fast, impl == 0, owned, nullptr, impl == 2);
    cudaEvent_t e0, e1;
    HP_CUDA(cudaEventCreate(&e0));
    HP_CUDA(cudaEventCreate(&e1));
    launch(0);  // warm-up
    const int reps = 20;
    HP_CUDA(cudaDeviceSynchronize());
    HP_CUDA(cudaEventRecord(e0, 0));
    for (int r = 0; r < reps; ++r) launch(0);   // back-to-back (PDL overlaps the prologues, as inside a plan)
    HP_CUDA(cudaEventRecord(e1, 0));
    HP_CUDA(cudaDeviceSynchronize());
    HP_CUDA(cudaGetLastError());
    float ms = 0.f;
    HP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    if (gpu_ms) *gpu_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (fast) {
      std::vector<__half> hbuf((size_t)M * N);
      HP_CUDA(cudaMemcpy(hbuf.data(), dout, hbuf.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < hbuf.size(); ++i) D[i] = __half2float(hbuf[i]);
    } else {
      HP_CUDA(cudaMemcpy(D, dout, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    }
  } catch (const hp::Error& e) {
    g_create_error = e.what();
    rc = e.code;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    rc = HMDPOSE_E_STATE;
  }
  for (void* q : owned) cudaFree(q);
  return rc;
}

}  // extern "C"
