// Engine: weights, anchors, buffers, per-batch launch plans (+CUDA graphs) for the phi0 hot path.
#include "engine.h"

#include <math.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels_simt.cuh"

namespace hp {

// EfficientNet-B0 trunk as instantiated by the reference (efficientnet/utils.py:235-241 expanded by
// repeats; skip rule efficientnet/model.py:100-104)
const BlockSpec kB0Blocks[16] = {
    {3, 1, 1, 32, 16, false},  {3, 2, 6, 16, 24, false},  {3, 1, 6, 24, 24, true},   {5, 2, 6, 24, 40, false},
    {5, 1, 6, 40, 40, true},   {3, 2, 6, 40, 80, false},  {3, 1, 6, 80, 80, true},   {3, 1, 6, 80, 80, true},
    {5, 1, 6, 80, 112, false}, {5, 1, 6, 112, 112, true}, {5, 1, 6, 112, 112, true}, {5, 2, 6, 112, 192, false},
    {5, 1, 6, 192, 192, true}, {5, 1, 6, 192, 192, true}, {5, 1, 6, 192, 192, true}, {3, 1, 6, 192, 320, false}};

static const char* kHeadNames[5] = {"box", "cls", "rot", "trans", "hand"};
static const char* kNodeNames[8] = {"conv6_up", "conv5_up", "conv4_up", "conv3_up",
                                    "conv4_down", "conv5_down", "conv6_down", "conv7_down"};
static const char* kFwNames[8] = {"p6_w1", "p5_w1", "p4_w1", "p3_w1", "p4_w2", "p5_w2", "p6_w2", "p7_w2"};

bool pdl_enabled() {
  static const bool on = std::getenv("HMDPOSE_NO_PDL") == nullptr;
  return on;
}

// ---- device watchdog record ---------------------------------------------------------------------
static TrapInfo* g_trap_host = nullptr;
static std::mutex g_trap_mu;
void trap_info_setup() {
  std::lock_guard<std::mutex> lock(g_trap_mu);
  if (!g_trap_host) {
    void* p = nullptr;
    HP_CUDA(cudaHostAlloc(&p, sizeof(TrapInfo), cudaHostAllocMapped | cudaHostAllocPortable));
    std::memset(p, 0, sizeof(TrapInfo));
    g_trap_host = (TrapInfo*)p;
  }
  void* dptr = nullptr;
  HP_CUDA(cudaHostGetDevicePointer(&dptr, g_trap_host, 0));
  HP_CUDA(trap_info_install_tu((TrapInfo*)dptr));
  trap_info_install_gemm_tu(dptr);
}
std::string trap_info_describe() {
  const volatile TrapInfo* t = g_trap_host;
  if (!t || t->code == 0) return "";
  char buf[256];
  std::snprintf(buf, sizeof(buf),
                " [device watchdog: mbarrier wait timed out after %.1f ms -- tag 0x%x block %u thread %u (warp %u) barrier "
                "smem 0x%x parity %u]",
                (double)t->waited_ns * 1e-6, t->tag, t->block, t->thread, t->thread >> 5, t->bar_addr, t->parity);
  return buf;
}

// TF "SAME" split (efficientnet/utils_extra.py:36-42)
void same_pad(int n, int k, int s, int* lo, int* hi) {
  const int extra = (cdiv(n, s) - 1) * s - n + k;
  *lo = extra / 2;
  *hi = extra - *lo;
}

// ---------------------------------------------------------------------------------------------
// weight blob
// ---------------------------------------------------------------------------------------------
struct BlobHeader { char magic[8]; uint32_t version, n_tensors, num_classes, reserved; };
struct BlobEntry { char name[96]; uint32_t ndim; uint32_t dims[4]; uint32_t pad; uint64_t offset, count; };
static_assert(sizeof(BlobHeader) == 24 && sizeof(BlobEntry) == 136, "blob layout");

void WeightBlob::parse(const void* blob, size_t bytes) {
  if (!blob || bytes < sizeof(BlobHeader)) throw Error(HMDPOSE_E_WEIGHTS, "weight blob too small");
  storage.assign((const uint8_t*)blob, (const uint8_t*)blob + bytes);
  const BlobHeader* h = reinterpret_cast<const BlobHeader*>(storage.data());
  if (std::memcmp(h->magic, "HMDPOSEW", 8) != 0 || h->version != 1)
    throw Error(HMDPOSE_E_WEIGHTS, "bad weight blob magic/version");
  num_classes = (int)h->num_classes;
  const size_t table = sizeof(BlobHeader) + (size_t)h->n_tensors * sizeof(BlobEntry);
  const size_t data0 = (table + 63) / 64 * 64;
  if (bytes < data0) throw Error(HMDPOSE_E_WEIGHTS, "weight blob truncated (table)");
  const BlobEntry* e = reinterpret_cast<const BlobEntry*>(storage.data() + sizeof(BlobHeader));
  for (uint32_t i = 0; i < h->n_tensors; ++i) {
    HostTensor t;
    size_t cnt = 1;
    for (uint32_t d = 0; d < e[i].ndim && d < 4; ++d) { t.dims.push_back((int)e[i].dims[d]); cnt *= e[i].dims[d]; }
    if (cnt != e[i].count || data0 + e[i].offset + cnt * 4 > bytes)
      throw Error(HMDPOSE_E_WEIGHTS, std::string("weight blob truncated at ") + e[i].name);
    t.data = reinterpret_cast<const float*>(storage.data() + data0 + e[i].offset);
    t.count = cnt;
    tensors[std::string(e[i].name, strnlen(e[i].name, sizeof(e[i].name)))] = t;
  }
}
const HostTensor& WeightBlob::get(const std::string& name) const {
  auto it = tensors.find(name);
  if (it == tensors.end()) throw Error(HMDPOSE_E_WEIGHTS, "weight blob lacks tensor " + name);
  return it->second;
}

// ---------------------------------------------------------------------------------------------
// anchors: generate_anchors / shift / translation_shift / anchors_for_shape
// (generators/utils/anchors.py:273-419).  Row order level -> y -> x -> a, a = scale*3 + ratio.
// ---------------------------------------------------------------------------------------------
int anchors_count(int S) {
  int n = 0;
  for (int l = 3; l <= 7; ++l) { const int s = (S + (1 << l) - 1) >> l; n += 9 * s * s; }
  return n;
}
void compute_anchors(int S, std::vector<float>& boxes, std::vector<float>& tanchors) {
  static const int sizes[5] = {32, 64, 128, 256, 512};
  static const int strides[5] = {8, 16, 32, 64, 128};
  // np.array([1, 0.5, 2], float32) and np.array([2**0, 2**(1/3), 2**(2/3)], float32) (anchors.py:63-64)
  const float ratios[3] = {1.0f, 0.5f, 2.0f};
  const float scales[3] = {1.0f, (float)pow(2.0, 1.0 / 3.0), (float)pow(2.0, 2.0 / 3.0)};
  const int n = anchors_count(S);
  boxes.resize((size_t)n * 4);
  tanchors.resize((size_t)n * 3);
  size_t row = 0;
  for (int li = 0; li < 5; ++li) {
    const int side = (S + (1 << (li + 3)) - 1) >> (li + 3);
    double base[9][4];
    for (int si = 0; si < 3; ++si)
      for (int ri = 0; ri < 3; ++ri) {
        // base_size * scales is evaluated in float32 (python int * float32 array), then float64
        const double sd = (double)((float)sizes[li] * scales[si]);
        const double area = sd * sd;
        const double w = sqrt(area / (double)ratios[ri]);
        const double h = w * (double)ratios[ri];
        double* b = base[si * 3 + ri];
        b[0] = 0.0 - w * 0.5; b[1] = 0.0 - h * 0.5; b[2] = w - w * 0.5; b[3] = h - h * 0.5;
      }
    for (int y = 0; y < side; ++y)
      for (int x = 0; x < side; ++x) {
        const double cx = (x + 0.5) * strides[li], cy = (y + 0.5) * strides[li];
        for (int a = 0; a < 9; ++a, ++row) {
          boxes[row * 4 + 0] = (float)(base[a][0] + cx);
          boxes[row * 4 + 1] = (float)(base[a][1] + cy);
          boxes[row * 4 + 2] = (float)(base[a][2] + cx);
          boxes[row * 4 + 3] = (float)(base[a][3] + cy);
          tanchors[row * 3 + 0] = (float)cx;
          tanchors[row * 3 + 1] = (float)cy;
          tanchors[row * 3 + 2] = (float)strides[li];
        }
      }
  }
}

void compute_anchors_d0(int S, std::vector<float>& yxyx) {
  static const int strides[5] = {8, 16, 32, 64, 128};
  const double scales[3] = {1.0, pow(2.0, 1.0 / 3.0), pow(2.0, 2.0 / 3.0)};
  const double ratios[3][2] = {{1.0, 1.0}, {1.4, 0.7}, {0.7, 1.4}};
  yxyx.clear();
  for (int li = 0; li < 5; ++li) {
    const int st = strides[li];
    int side = 0;  // len(np.arange(stride / 2, S, stride))
    while (st / 2.0 + (double)side * st < (double)S) ++side;
    for (int y = 0; y < side; ++y)
      for (int x = 0; x < side; ++x) {
        const double cy = st / 2.0 + (double)y * st, cx = st / 2.0 + (double)x * st;
        for (int si = 0; si < 3; ++si)
          for (int ri = 0; ri < 3; ++ri) {
            const double base = 4.0 * st * scales[si];
            const double ax2 = base * ratios[ri][0] / 2.0, ay2 = base * ratios[ri][1] / 2.0;
            yxyx.push_back((float)(cy - ay2)); yxyx.push_back((float)(cx - ax2));
            yxyx.push_back((float)(cy + ay2)); yxyx.push_back((float)(cx + ax2));
          }
      }
  }
}

// v2 depthwise tiling: block = cvb channel vectors x sw strips (DW2_OWT output pixels each) x sh rows
static void dw2_tiling(int C, int V, int Ho, int Wo, DwGroup& g) {
  const int CV = C / V;
  g.cv_chunks = cdiv(CV, 32);
  g.cvb = cdiv(CV, g.cv_chunks);
  const int strips = cdiv(Wo, DW2_OWT);
  g.sw = std::max(1, std::min(strips, std::min(8, 256 / g.cvb)));
  g.sh = std::max(1, std::min(Ho, 256 / (g.cvb * g.sw)));
  g.tiles_x = cdiv(strips, g.sw);
  g.tiles_y = cdiv(Ho, g.sh);
}

// v3 depthwise tiling (dw3_kernel): cb = largest divisor of the channel-vector count <= 8, then the (th, tw) that
// uses the most threads within 46 KB of shared memory (input tile + taps + squeeze scratch); returns smem bytes
static int dw3_tiling(int C, int V, int K, int S, int Ho, int Wo, DwGroup& g) {
  const int CV = C / V;
  int cb = 1;
  for (int d = 8; d >= 1; --d)
    if (CV % d == 0) { cb = d; break; }
  g.cb = cb;
  g.cv_chunks = CV / cb;
  // threads per block: 3x3 stencils do better with blocks of <= 128 threads (more independent TMA tiles in flight per SM:
  // blk0 / blk2 14.1 / 20.7 -> 11.3 / 15.6 us with 8 steps in flight), the 5x5 ones with 256 (HMDPOSE_DW_MAXTHREADS overrides)
  static const int thr_env = std::getenv("HMDPOSE_DW_MAXTHREADS") ? std::max(32, std::min(256, std::atoi(std::getenv("HMDPOSE_DW_MAXTHREADS")))) : 0;
  const int thr_cap = thr_env ? thr_env : (K == 3 ? 128 : 256);
  const int ns_max = thr_cap / cb;
  const int wo4 = ((Wo + 3) / 4) * 4;
  int best_threads = 0, best_smem = 0;
  g.th = 1; g.tw = 4;
  for (int tw : {32, 16, 8, 4}) {
    if (tw > std::max(4, wo4)) continue;
    for (int th : {16, 8, 4, 2, 1}) {
      if (th > Ho || th * tw / 4 > ns_max) continue;
      const int threads = cb * th * tw / 4;
      const int ih = (th - 1) * S + K, iw = (tw - 1) * S + K;
      const int smem = ih * iw * cb * 16 + K * K * cb * V * 4 + threads * V * 4;
      if (smem > 46 * 1024) continue;
      if (threads > best_threads || (threads == best_threads && smem < best_smem)) {
        best_threads = threads; best_smem = smem; g.th = th; g.tw = tw;
      }
    }
  }
  if (best_threads == 0) {  // 1 x 4 strip always fits
    best_threads = cb;
    best_smem = K * ((3) * S + K) * cb * 16 + K * K * cb * V * 4 + cb * V * 4;
  }
  g.tiles_per_img = cdiv(Ho, g.th) * cdiv(Wo, g.tw);
  return best_smem;
}

// ---------------------------------------------------------------------------------------------
Plan::~Plan() {
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  for (void* p : owned) cudaFree(p);
}

void* Engine::dalloc(size_t bytes) {
  void* p = nullptr;
  HP_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
  allocs_.push_back(p);
  return p;
}
float* Engine::upload_f32(const float* src, size_t n) {
  float* d = (float*)dalloc(n * 4);
  HP_CUDA(cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice));
  return d;
}
template <> void* Engine::upload_as<float>(const float* src, size_t n) { return upload_f32(src, n); }
template <> void* Engine::upload_as<__half>(const float* src, size_t n) {
  std::vector<__half> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(src[i]);
  void* d = dalloc(n * 2);
  HP_CUDA(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  return d;
}
void Engine::reg_debug(const std::string& name, const Tens& t, bool is_t) { debug_[name] = {t, is_t}; }

// Device weights.  GEMM weights [N][K] are stored in the activation type T (fp16 in fast mode);
// depthwise taps are transposed to [k*k][C] so that a channel vector is one 16/32-byte load; the stem
// is transposed to [(ky,kx,ci)][co].  Everything else (biases, SE matrices) stays fp32.
template <typename T>
void Engine::upload_weights() {
  auto f32 = [&](const std::string& n) { const HostTensor& t = blob_.get(n); wdev_[n] = upload_f32(t.data, t.count); };
  auto gemm_w = [&](const std::string& n) { const HostTensor& t = blob_.get(n); wdev_[n] = upload_as<T>(t.data, t.count); };
  auto dw_w = [&](const std::string& n) {
    const HostTensor& t = blob_.get(n);  // [C][k][k]
    const int C = t.dims[0], kk = t.dims[1] * t.dims[2];
    std::vector<float> tr((size_t)C * kk);
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < kk; ++j) tr[(size_t)j * C + c] = t.data[(size_t)c * kk + j];
    wdev_[n] = upload_f32(tr.data(), tr.size());
  };
  {
    const HostTensor& t = blob_.get("stem.w");  // [32][3][3][3] = co, ci, ky, kx
    std::vector<float> tr(27 * 32);
    for (int co = 0; co < 32; ++co)
      for (int ci = 0; ci < 3; ++ci)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            tr[((ky * 3 + kx) * 3 + ci) * 32 + co] = t.data[((co * 3 + ci) * 3 + ky) * 3 + kx];
    wdev_["stem.w"] = upload_f32(tr.data(), tr.size());
    f32("stem.b");
    // host copy: the stem kernel takes its weights as a kernel parameter (constant-bank operands)
    const HostTensor& tb = blob_.get("stem.b");
    std::memcpy(stem_wb_.w, tr.data(), sizeof(stem_wb_.w));
    std::memcpy(stem_wb_.b, tb.data, sizeof(stem_wb_.b));
  }
  for (int i = 0; i < 16; ++i) {
    const std::string b = "blk" + std::to_string(i);
    if (kB0Blocks[i].e != 1) { gemm_w(b + ".exp.w"); f32(b + ".exp.b"); }
    dw_w(b + ".dw.w"); f32(b + ".dw.b");
    f32(b + ".se_r.w"); f32(b + ".se_r.b"); f32(b + ".se_e.w"); f32(b + ".se_e.b");
    {
      const HostTensor& t = blob_.get(b + ".se_e.w");  // [C][Cse] -> transposed [Cse][C] for coalesced reads
      const int C = t.dims[0], Cse = t.dims[1];
      std::vector<float> tr((size_t)C * Cse);
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < Cse; ++j) tr[(size_t)j * C + c] = t.data[(size_t)c * Cse + j];
      wdev_[b + ".se_e.wT"] = upload_f32(tr.data(), tr.size());
    }
    gemm_w(b + ".proj.w"); f32(b + ".proj.b");
  }
  for (int c = 0; c < 3; ++c) {
    const std::string p = "bifpn" + std::to_string(c);
    for (int n = 0; n < 8; ++n) {
      const std::string q = p + "." + kNodeNames[n];
      dw_w(q + ".dw.w"); gemm_w(q + ".pw.w"); f32(q + ".pw.b");
    }
  }
  for (const char* q : {"p3_dc", "p4_dc", "p5_dc", "p5_to_p6", "p4_dc2", "p5_dc2"}) {
    gemm_w(std::string("bifpn0.") + q + ".w"); f32(std::string("bifpn0.") + q + ".b");
  }
  num_heads_ = blob_.has("head.rot.l0.dw.w") ? 5 : 2;  // EfficientDet checkpoints carry regressor + classifier only
  for (int h = 0; h < num_heads_; ++h) {
    const std::string p = std::string("head.") + kHeadNames[h];
    for (int i = 0; i < 3; ++i) {
      dw_w(p + ".l" + std::to_string(i) + ".dw.w");
      for (int l = 0; l < 5; ++l) {
        const std::string q = p + ".l" + std::to_string(i) + ".lvl" + std::to_string(l);
        gemm_w(q + ".pw.w"); f32(q + ".pw.b");
      }
    }
    const int nh = (h == 3) ? 2 : 1;
    for (int j = 0; j < nh; ++j) {
      const std::string q = p + ".hdr" + std::to_string(j);
      dw_w(q + ".dw.w"); gemm_w(q + ".pw.w"); f32(q + ".pw.b");
    }
  }
  // --iter 1 refinement sub-nets: the (64 + P)-channel separable conv is stored with its channel count padded to
  // a multiple of 32 (zero taps / zero weight columns), so the vector kernels and the tcgen05 GEMM apply unchanged
  iter1_ = num_heads_ == 5 && blob_.has("head.rot.it.dw.w");
  for (int t = 0; t < 3 && iter1_; ++t) {
    const std::string p = std::string("head.") + kHeadNames[2 + t] + ".it";
    const HostTensor& dw = blob_.get(p + ".dw.w");   // [Cin][3][3]
    const HostTensor& pw = blob_.get(p + ".pw.w");   // [64][Cin]
    const int Cin = dw.dims[0], Cpad = ((Cin + 31) / 32) * 32;
    if (pw.dims[0] != 64 || pw.dims[1] != Cin) throw Error(HMDPOSE_E_WEIGHTS, "unexpected refinement sub-net shape: " + p);
    std::vector<float> dwt((size_t)9 * Cpad, 0.f), pwp((size_t)64 * Cpad, 0.f);
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < 9; ++j) dwt[(size_t)j * Cpad + c] = dw.data[(size_t)c * 9 + j];
    for (int n = 0; n < 64; ++n)
      for (int c = 0; c < Cin; ++c) pwp[(size_t)n * Cpad + c] = pw.data[(size_t)n * Cin + c];
    wdev_[p + ".dw.w"] = upload_f32(dwt.data(), dwt.size());
    wdev_[p + ".pw.w"] = upload_as<T>(pwp.data(), pwp.size());
    f32(p + ".pw.b");
    const int nh = (t == 1) ? 2 : 1;
    for (int j = 0; j < nh; ++j) {
      const std::string q = p + ".hdr" + std::to_string(j);
      dw_w(q + ".dw.w"); gemm_w(q + ".pw.w"); f32(q + ".pw.b");
    }
  }
}

// Folded tap matrices of a separable block for the implicit-GEMM kernel (sepconv3_tc.cuh):
// W9[tap][n][c] = w_pw[n][c] * w_dw[c][tap], fp32 product rounded once to fp16.
const void* Engine::w9_for(const std::string& dw_name, const std::string& pw_name) {
  const std::string key = pw_name + "#w9";
  auto it = wdev_.find(key);
  if (it != wdev_.end()) return it->second;
  const HostTensor& dw = blob_.get(dw_name);   // [64][3][3]
  const HostTensor& pw = blob_.get(pw_name);   // [N][64]
  const int C = dw.dims[0], N = pw.dims[0];
  if (C != 64 || pw.dims[1] != 64 || dw.dims[1] != 3 || dw.dims[2] != 3)
    throw Error(HMDPOSE_E_WEIGHTS, "unexpected separable-conv weight shape: " + pw_name);
  std::vector<float> w9((size_t)9 * N * 64);
  for (int tap = 0; tap < 9; ++tap)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < 64; ++c)
        w9[((size_t)tap * N + n) * 64 + c] = pw.data[(size_t)n * 64 + c] * dw.data[(size_t)c * 9 + tap];
  void* d = upload_as<__half>(w9.data(), w9.size());
  wdev_[key] = d;
  return d;
}

template <typename T>
void Engine::alloc_buffers() {
  const int S = cfg.image_size, b = mb_;
  auto mk = [&](int H, int W, int C, void* shared = nullptr) {
    Tens t; t.H = H; t.W = W; t.C = C;
    t.p = shared ? shared : dalloc(t.elems(b) * sizeof(T));
    return t;
  };
  stem_out_ = mk(S / 2, S / 2, 32);
  reg_debug("stem", stem_out_);
  // the 6x-expanded tensors of all blocks share two scratch buffers (re-used addresses stay in L2)
  size_t max_exp = 0, max_dw = 0;
  {
    int H = S / 2;
    for (int i = 0; i < 16; ++i) {
      const BlockSpec& bs = kB0Blocks[i];
      const int Ho = cdiv(H, bs.s);
      max_exp = std::max(max_exp, (size_t)b * H * H * bs.cin * bs.e);
      max_dw = std::max(max_dw, (size_t)b * Ho * Ho * bs.cin * bs.e);
      H = Ho;
    }
  }
  if (!keep_all_) { scratch_exp_ = dalloc(max_exp * sizeof(T)); scratch_dw_ = dalloc(max_dw * sizeof(T)); }
  int H = S / 2;
  for (int i = 0; i < 16; ++i) {
    const BlockSpec& bs = kB0Blocks[i];
    const int Ho = cdiv(H, bs.s), cexp = bs.cin * bs.e;
    BlockBufs& bb = blk_[i];
    if (bs.e != 1) bb.exp = mk(H, H, cexp, keep_all_ ? nullptr : scratch_exp_);
    bb.dw = mk(Ho, Ho, cexp, keep_all_ ? nullptr : scratch_dw_);
    bb.out = mk(Ho, Ho, bs.cout);
    {
      DwGroup tg;
      dw2_tiling(cexp, VecN<T>::N, Ho, Ho, tg);
      bb.tiles = std::max(cdiv(Ho * Ho, DW_TP), tg.tiles_x * tg.tiles_y);  // capacity: v1 tiling or v2 with rep = 1
      DwGroup t3;
      dw3_tiling(cexp, VecN<T>::N, bs.k, bs.s, Ho, Ho, t3);
      bb.tiles = std::max(bb.tiles, t3.tiles_per_img);
      bb.tiles = std::max(bb.tiles, ed_tiles_per_img(bs.k, bs.s, Ho, Ho));   // fused expand + depthwise tiling
    }
    bb.se_partial = (float*)dalloc((size_t)b * bb.tiles * cexp * 4);
    bb.gate = (float*)dalloc((size_t)b * cexp * 4);
    // gate folded into per-image project weights: maps whose images are whole 128-row m tiles and whose weight panel is
    // small enough for se3_kernel to rescale on the side (blocks 0-4: <= 9.6 k weights; for the 75 k weights of blocks
    // 6-10 the rescaling cost se3 more than the un-gated GEMM saved: 33 -> 110 us per step with 8 steps in flight)
    static const int wgate_cap = std::getenv("HMDPOSE_WGATE_CAP") ? std::atoi(std::getenv("HMDPOSE_WGATE_CAP")) : 16384;
    if (fast_ && sizeof(T) == 2 && (Ho * Ho) % 128 == 0 && bs.cout * cexp <= wgate_cap)
      bb.wgated = dalloc((size_t)b * bs.cout * cexp * sizeof(T));
    reg_debug("blk" + std::to_string(i), bb.out);
    if (keep_all_) {
      if (bs.e != 1) reg_debug("blk" + std::to_string(i) + ".exp", bb.exp);
      reg_debug("blk" + std::to_string(i) + ".dw", bb.dw);
      Tens g; g.p = bb.gate; g.H = 1; g.W = 1; g.C = cexp;
      reg_debug("blk" + std::to_string(i) + ".gate", g, false);
    }
    H = Ho;
  }
  for (int l = 0; l < 5; ++l) {
    lvl_side_[l] = (S + (1 << (l + 3)) - 1) >> (l + 3);
    lvl_hw_[l] = lvl_side_[l] * lvl_side_[l];
    lvl_off_[l] = l == 0 ? 0 : lvl_off_[l - 1] + 9 * lvl_hw_[l - 1];
  }
  for (int c = 0; c < 3; ++c) {
    CellBufs& cb = cell_[c];
    for (int l = 0; l < 5; ++l) {
      const int s = lvl_side_[l];
      if (c == 0) cb.in[l] = mk(s, s, 64);
      cb.out[l] = mk(s, s, 64);
      if (l >= 1 && l <= 3) {
        cb.up[l] = mk(s, s, 64);
        reg_debug("cell" + std::to_string(c) + ".up" + std::to_string(l + 3), cb.up[l]);
      }
      cb.fused[l] = mk(s, s, 64);
      cb.dwb[l] = mk(s, s, 64);
      reg_debug("cell" + std::to_string(c) + ".p" + std::to_string(l + 3), cb.out[l]);
    }
    if (c == 0) {
      cb.in2[0] = mk(lvl_side_[1], lvl_side_[1], 64);
      cb.in2[1] = mk(lvl_side_[2], lvl_side_[2], 64);
      cb.p6_pre = mk(lvl_side_[2], lvl_side_[2], 64);
      for (int l = 0; l < 5; ++l) reg_debug("cell0.in" + std::to_string(l + 3), cb.in[l]);
      reg_debug("cell0.in4b", cb.in2[0]);
      reg_debug("cell0.in5b", cb.in2[1]);
    }
  }
  for (int h = 0; h < 5; ++h)
    for (int l = 0; l < 5; ++l) {
      const int s = lvl_side_[l];
      trunk_[h][l][0] = mk(s, s, 64);
      trunk_[h][l][1] = mk(s, s, 64);
      if (keep_all_) trunk_[h][l][2] = mk(s, s, 64);
      hdw_[h][l] = mk(s, s, 64);
      reg_debug(std::string("trunk.") + kHeadNames[h] + ".p" + std::to_string(l + 3), trunk_[h][l][trunk_final()]);
      if (keep_all_)
        for (int i = 0; i < 3; ++i)
          reg_debug("trunk" + std::to_string(i) + "." + kHeadNames[h] + ".p" + std::to_string(l + 3), trunk_[h][l][i]);
    }
  for (int j = 0; j < 6; ++j)
    for (int l = 0; l < 5; ++l) hdrdw_[j][l] = mk(lvl_side_[l], lvl_side_[l], 64);
  if (iter1_) {
    static const int kItCpad[3] = {96, 96, 640};   // 64 + 27, 64 + 27, 64 + 567 padded to a multiple of 32
    for (int t = 0; t < 3; ++t)
      for (int l = 0; l < 5; ++l) {
        const int sd = lvl_side_[l];
        it_in_[t][l] = mk(sd, sd, kItCpad[t]);
        it_dw_[t][l] = mk(sd, sd, kItCpad[t]);
        it_y_[t][l] = mk(sd, sd, 64);
      }
    for (int j = 0; j < 4; ++j)
      for (int l = 0; l < 5; ++l) it_hdw_[j][l] = mk(lvl_side_[l], lvl_side_[l], 64);
  }

  const int C = cfg.num_classes, D = cfg.max_detections;
  o_reg_ = (float*)dalloc((size_t)b * N * 4 * 4);
  o_cls_ = (float*)dalloc((size_t)b * N * C * 4);
  o_rot_ = (float*)dalloc((size_t)b * N * 3 * 4);
  o_traw_ = (float*)dalloc((size_t)b * N * 3 * 4);
  o_hand_ = (float*)dalloc((size_t)b * N * HMDPOSE_NUM_HAND * 4);
  p_boxes_ = (float*)dalloc((size_t)b * N * 4 * 4);
  p_trans_ = (float*)dalloc((size_t)b * N * 3 * 4);
  pb_.cap = 1;
  while (pb_.cap < N) pb_.cap <<= 1;
  pb_.keys = (unsigned long long*)dalloc((size_t)b * C * pb_.cap * 8);
  pb_.kept_idx = (int*)dalloc((size_t)b * C * D * 4);
  pb_.kept_score = (float*)dalloc((size_t)b * C * D * 4);
  pb_.kept_count = (int*)dalloc((size_t)b * C * 4);
  det_boxes_ = (float*)dalloc((size_t)b * D * 4 * 4);
  det_scores_ = (float*)dalloc((size_t)b * D * 4);
  det_labels_ = (int32_t*)dalloc((size_t)b * D * 4);
  det_rot_ = (float*)dalloc((size_t)b * D * 3 * 4);
  det_trans_ = (float*)dalloc((size_t)b * D * 3 * 4);
  det_hand_ = (float*)dalloc((size_t)b * D * HMDPOSE_NUM_HAND * 4);
  det_idx_ = (int32_t*)dalloc((size_t)b * D * 4);
  d_best_ = (float*)dalloc((size_t)b * HMDPOSE_BEST_LEN * 4);
  d_cam_local_ = (float*)dalloc((size_t)b * 6 * 4);
  se_counters_ = (int*)dalloc((size_t)16 * b * 4);
  HP_CUDA(cudaMemset(se_counters_, 0, (size_t)16 * b * 4));
  // split-K partial tiles of the fused MBConv kernel (mbconv_tc.cuh): [b][cl <= 6][<= 256 px][<= 128 ch] fp32, one
  // L2-resident scratch shared by all blocks of a step (launches of a stream are ordered)
  if (fast_) {
    mb_part_bytes_ = std::max((size_t)b * 6 * 256 * 128 * 4, (size_t)cdiv(b * 256, 128) * 6 * 128 * 320 * 4);   // mbconv / projk
    mb_part_ = (float*)dalloc(mb_part_bytes_);
  }
}

// ---------------------------------------------------------------------------------------------
// launch plan for `b` frames: everything after the stem up to the five head tensors
// ---------------------------------------------------------------------------------------------
template <typename T>
std::unique_ptr<Plan> Engine::build_plan(int b, int mode) {
  std::unique_ptr<Plan> plan(new Plan());
  plan->b = b;
  plan->mode = mode;
  const double sT = sizeof(T);
  int last_dw_tiles = 0;
  bool se_inplace = false;
  bool d0_fused = false;
  std::vector<Step>& steps = plan->steps;
  std::vector<void*>& owned = plan->owned;
  auto W = [&](const std::string& n) -> void* {
    auto it = wdev_.find(n);
    if (it == wdev_.end()) throw Error(HMDPOSE_E_WEIGHTS, "device weight missing: " + n);
    return it->second;
  };
  auto add_dw = [&](const std::string& name, std::vector<DwGroup> gs, bool fused = false) {
    constexpr int V = VecN<T>::N;
    int blocks = 0;
    const int K = gs[0].k, S = gs[0].stride;
    if (!v1_ && !fused && gs.size() == 1 && std::getenv("HMDPOSE_DW2") == nullptr) {
      DwGroup& g = gs[0];
      int smem = dw3_tiling(g.C, V, K, S, g.Ho, g.Wo, g);
      const int threads = g.cb * g.th * g.tw / 4;
      if (g.se_counter) smem = std::max(smem, (threads * 4 + g.C + 64) * 4);   // se_tail scratch
      blocks = b * g.tiles_per_img * g.cv_chunks;
      g.block_start = 0; g.nblocks = blocks;
      last_dw_tiles = g.tiles_per_img;
      {
        CUtensorMap tm;
        encode_act_4d(&tm, g.in, sizeof(T) == 2, g.C, g.W, g.H, b, g.cb * V, (g.tw - 1) * S + K, (g.th - 1) * S + K);
        CUtensorMap* dtm = nullptr;
        HP_CUDA(cudaMalloc(&dtm, sizeof(CUtensorMap)));
        HP_CUDA(cudaMemcpy(dtm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        owned.push_back(dtm);
        g.tmap = dtm;
      }
      const DwGroup gv = g;   // passed by value: kernel parameter space
      void (*kern)(const DwGroup) = nullptr;
#define HP_DW3(KK, SS)                                                   \
  if (K == KK && S == SS) {                                              \
    if (g.cb == 4) kern = dw3_kernel<T, KK, SS, 4>;                      \
    else if (g.cb == 6) kern = dw3_kernel<T, KK, SS, 6>;                 \
    else if (g.cb == 7) kern = dw3_kernel<T, KK, SS, 7>;                 \
    else if (g.cb == 8) kern = dw3_kernel<T, KK, SS, 8>;                 \
  }
      HP_DW3(3, 1) HP_DW3(3, 2) HP_DW3(5, 1) HP_DW3(5, 2)
#undef HP_DW3
      if (!kern) throw Error(HMDPOSE_E_STATE, "unsupported depthwise stencil / channel blocking");
      Step s;
      s.name = name;
      s.kernel = "dw3_kernel";
      s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(kern, dim3(blocks), dim3(threads), smem, st, gv)); };
      const double oe = (double)b * g.Ho * g.Wo * g.C;
      s.bytes = (double)b * g.H * g.W * g.C * sT + oe * sT + (double)K * K * g.C * 4 +
                (g.se_partial ? (double)b * g.tiles_per_img * g.C * 4 : 0.0);
      s.flops = 2.0 * K * K * oe;
      steps.push_back(s);
      return;
    }
    for (DwGroup& g : gs) {
      if (g.k != K || g.stride != S) throw Error(HMDPOSE_E_STATE, "mixed stencils in one depthwise launch");
      const int CV = g.C / V;
      g.cv_chunks = cdiv(CV, 32);
      g.cvb = cdiv(CV, g.cv_chunks);
      if (v1_) {
        g.tiles_per_img = cdiv(g.Ho * g.Wo, DW_TP);
        g.nblocks = b * g.tiles_per_img * g.cv_chunks;
      } else {
        dw2_tiling(g.C, V, g.Ho, g.Wo, g);
        g.rep = 1;
        g.tiles_per_img = g.tiles_x * cdiv(g.tiles_y, g.rep);
        g.nblocks = b * g.tiles_per_img * g.cv_chunks;
      }
      g.block_start = blocks;
      blocks += g.nblocks;
    }
    last_dw_tiles = gs[0].tiles_per_img;
    DwGroup* d = nullptr;
    HP_CUDA(cudaMalloc(&d, sizeof(DwGroup) * gs.size()));
    HP_CUDA(cudaMemcpy(d, gs.data(), sizeof(DwGroup) * gs.size(), cudaMemcpyHostToDevice));
    owned.push_back(d);
    const int n = (int)gs.size();
    Step s;
    s.name = name;
    if (v1_) {
      s.kernel = "dw_kernel";
      s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(dw_kernel<T>, dim3(blocks), dim3(DW_THREADS), 0, st, d, n)); };
    } else {
      int cvb_max = 1;
      for (const DwGroup& g : gs) cvb_max = std::max(cvb_max, g.cvb);
      const int smem = (K * K * cvb_max * V + 256 * V) * 4;
      s.kernel = fused ? "dw2_kernel<fused>" : "dw2_kernel";
      void (*kern)(const DwGroup*, int) = nullptr;
      if (fused && K == 3 && S == 1) kern = dw2_kernel<T, 3, 1, true>;
      else if (K == 3 && S == 1) kern = dw2_kernel<T, 3, 1, false>;
      else if (K == 3 && S == 2) kern = dw2_kernel<T, 3, 2, false>;
      else if (K == 5 && S == 1) kern = dw2_kernel<T, 5, 1, false>;
      else if (K == 5 && S == 2) kern = dw2_kernel<T, 5, 2, false>;
      else throw Error(HMDPOSE_E_STATE, "unsupported depthwise stencil");
      s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(kern, dim3(blocks), dim3(256), smem, st, d, n)); };
    }
    for (const DwGroup& g : gs) {
      const double oe = (double)b * g.Ho * g.Wo * g.C;
      double in_e = (double)b * g.H * g.W * g.C;
      if (fused) {
        auto rs_elems = [&](int m) { return m == RS_UP2 ? 0.25 : (m == RS_POOL ? 4.0 : (m == RS_SAME ? 1.0 : 0.0)); };
        in_e *= 1.0 + rs_elems(g.mode_b) + rs_elems(g.mode_c);
      }
      s.bytes += in_e * sT + oe * sT + (double)g.k * g.k * g.C * 4 +
                 (g.se_partial ? (double)b * g.tiles_per_img * g.C * 4 : 0.0);
      s.flops += 2.0 * g.k * g.k * oe;
    }
    steps.push_back(s);
  };
  auto add_gemm = [&](const std::string& name, std::vector<GemmProb> ps) {
    Step s;
    s.name = name;
    for (const GemmProb& p : ps) {
      s.bytes += (double)p.M * p.K * sT + (double)p.N * p.K * sT + p.N * 4.0 +
                 (double)p.M * p.N * (p.out_mode ? 4.0 : sT) + (p.residual ? (double)p.M * p.N * sT : 0.0) +
                 (p.a_scale ? (double)cdiv(p.M, p.rows_per_img) * p.K * 4 : 0.0);
      s.flops += 2.0 * p.M * p.N * p.K;
    }
    s.launch = make_gemm_launcher(std::move(ps), fast_, force_simt_, owned, &s.kernel, v1_);
    steps.push_back(s);
  };
  const bool use_sep = fast_ && !v1_ && !force_simt_ && std::getenv("HMDPOSE_NO_SEPCONV") == nullptr;
  auto add_sep = [&](const std::string& name, std::vector<SepSpec> specs) {
    Step s;
    s.name = name;
    s.kernel = "sepconv_kernel";
    for (const SepSpec& q : specs) {
      auto rs_elems = [&](int m) { return m == RS_UP2 ? 0.25 : (m == RS_POOL ? 4.0 : (m == RS_SAME ? 1.0 : 0.0)); };
      const double px = (double)b * q.H * q.W;
      s.bytes += px * 64 * sT * (1.0 + (q.fused ? rs_elems(q.mode_b) + rs_elems(q.mode_c) : 0.0)) +
                 px * q.p.N * (q.p.out_mode ? 4.0 : sT) + q.p.N * 64 * sT + q.p.N * 4.0 + 9 * 64 * 4.0;
      s.flops += 2.0 * px * 64 * (9.0 + q.p.N);
    }
    s.launch = make_sepconv_launcher(std::move(specs), owned, &s.kernel);
    steps.push_back(s);
  };
  auto sep_spec = [&](const Tens& in, const std::string& dw, const std::string& w, const std::string& bias, int N_, int act,
                      void* out) {
    SepSpec q;
    std::memset(&q.p, 0, sizeof(q.p));
    q.p.W = W(w); q.p.bias = (const float*)W(bias); q.p.N = N_; q.p.ldo = N_; q.p.act = act; q.p.out = out;
    q.p.p_src = 1; q.p.p_dst = 1;
    q.in = in.p; q.dw_w = (const float*)W(dw); q.H = in.H; q.W = in.W; q.Bn = b;
    if (std::is_same<T, __half>::value) q.w9 = w9_for(dw, w);
    return q;
  };
  auto gemm_prob = [&](const Tens& in, const std::string& w, const std::string& bias, int N_, int act, void* out) {
    GemmProb p;
    std::memset(&p, 0, sizeof(p));
    p.A = in.p; p.W = W(w); p.bias = (const float*)W(bias);
    p.M = b * in.H * in.W; p.N = N_; p.K = in.C; p.lda = in.C; p.ldo = N_;
    p.act = act; p.rows_per_img = in.H * in.W; p.out = out; p.p_src = 1; p.p_dst = 1;
    return p;
  };
  auto dw_group = [&](const Tens& in, const Tens& out, const std::string& w, const float* bias, float* se, int k,
                      int s, int act) {
    DwGroup g;
    std::memset(&g, 0, sizeof(g));
    int lo, hi;
    same_pad(in.H, k, s, &lo, &hi);
    g.in = in.p; g.out = out.p; g.w = (const float*)W(w); g.bias = bias; g.se_partial = se;
    g.H = in.H; g.W = in.W; g.Ho = out.H; g.Wo = out.W; g.C = in.C; g.k = k; g.stride = s; g.pad = lo; g.act = act;
    return g;
  };

  // ---- backbone: 16 MBConv blocks (efficientnet/model.py:69-104) ----
  // HMDPOSE_MBFUSE=1 (opt-in): blocks 6-15 at 256x256 as one cluster kernel each (mbconv_tc.cuh).  Measured at batch 16
  // (profiles/r2_steps_b16_insitu_mbfuse.txt): 56 instead of 86 launches and 328 instead of 410 us for the ten blocks on
  // one stream, but 217 instead of 130 us per step with 8 steps in flight (25.1 k vs 28.3 k frames/s): a 200 KB-smem /
  // 512-TMEM-column cluster CTA holds its SM exclusively while it sits in latency-bound phases -- so the four-launch path
  // stays the default until the kernel is below ~15 us per block.
  const bool use_mbfuse = fast_ && !v1_ && !force_simt_ && mbfuse_;
  const unsigned mbfuse_mask = std::getenv("HMDPOSE_MBFUSE_MASK") ? (unsigned)std::strtoul(std::getenv("HMDPOSE_MBFUSE_MASK"), nullptr, 0) : 0xFFFFu;
  Tens x = stem_out_;
  for (int i = 0; i < 16; ++i) {
    const BlockSpec& bs = kB0Blocks[i];
    const std::string n = "blk" + std::to_string(i);
    BlockBufs& bb = blk_[i];
    bool use_wgate = false;
    // small maps (<= 256 pixels per image): the whole block as ONE cluster kernel, the 6x tensor never leaves the SM
    if (std::is_same<T, __half>::value && use_mbfuse && bs.e != 1 && ((mbfuse_mask >> i) & 1)) {
      MbSpec ms;
      std::memset(&ms, 0, sizeof(ms));
      ms.x = (const __half*)x.p; ms.out = (__half*)bb.out.p;
      ms.w_exp = (const __half*)W(n + ".exp.w"); ms.b_exp = (const float*)W(n + ".exp.b");
      ms.w_dw = (const float*)W(n + ".dw.w"); ms.b_dw = (const float*)W(n + ".dw.b");
      ms.se_wr = (const float*)W(n + ".se_r.w"); ms.se_br = (const float*)W(n + ".se_r.b");
      ms.se_weT = (const float*)W(n + ".se_e.wT"); ms.se_be = (const float*)W(n + ".se_e.b");
      ms.w_proj = (const __half*)W(n + ".proj.w"); ms.b_proj = (const float*)W(n + ".proj.b");
      if (keep_all_) { ms.dbg_exp = (__half*)bb.exp.p; ms.dbg_dw = (__half*)bb.dw.p; ms.gate_out = bb.gate; }
      int lo, hi;
      same_pad(x.H, bs.k, bs.s, &lo, &hi);
      ms.H = x.H; ms.W = x.W; ms.Ho = bb.dw.H; ms.Wo = bb.dw.W; ms.cin = bs.cin; ms.cexp = bs.cin * bs.e; ms.cout = bs.cout;
      ms.cse = std::max(1, bs.cin / 4); ms.k = bs.k; ms.stride = bs.s; ms.pad = lo; ms.skip = bs.skip ? 1 : 0;
      ms.inv_hw = 1.0f / (float)(bb.dw.H * bb.dw.W);
      if (const char* ce = std::getenv("HMDPOSE_MB_CL")) ms.cl = std::max(1, std::min(8, std::atoi(ce)));   // cluster-size cap (experiments)
      auto launch = make_mbconv_launcher(ms, b, owned, mb_part_, mb_part_bytes_);
      if (launch) {
        Step s;
        s.name = n + ".mbconv";
        s.kernel = "mbconv_fused_kernel";
        s.launch = launch;
        const double px = (double)b * x.H * x.W, pxo = (double)b * bb.dw.H * bb.dw.W;
        s.bytes = px * bs.cin * sT * (bs.skip ? 2.0 : 1.0) + pxo * bs.cout * sT +
                  ((double)ms.cexp * (bs.cin + bs.cout) * sT + (double)ms.cexp * (bs.k * bs.k + 2.0 * ms.cse + 3) * 4);
        s.flops = 2.0 * px * bs.cin * ms.cexp + 2.0 * pxo * ms.cexp * (bs.k * bs.k + bs.cout) + 4.0 * b * ms.cexp * ms.cse;
        steps.push_back(s);
        x = bb.out;
        continue;
      }
    }
    // HMDPOSE_EXPDW=1 (opt-in): expand fused into the depthwise kernel (expdw_tc.cuh), the 6x tensor never leaves the SM.
    // Measured at batch 16 (blocks 1-5): 161 us in situ against 160 us for expand GEMM + dw3_kernel, 149 against 117 us per
    // step with 8 steps in flight -- the halo recompute and the CUDA-core epilogue + stencil of one fat CTA per SM cost
    // more issue slots than the two HBM/L2 round trips of the 6x tensor they save, so the pair of launches stays the default.
    bool did_expdw = false;
    if (std::is_same<T, __half>::value && fast_ && !v1_ && !force_simt_ && bs.e != 1 && expdw_) {
      EdSpec es;
      std::memset(&es, 0, sizeof(es));
      int lo, hi;
      same_pad(x.H, bs.k, bs.s, &lo, &hi);
      es.x = (const __half*)x.p; es.out = (__half*)bb.dw.p;
      es.w_exp = (const __half*)W(n + ".exp.w"); es.b_exp = (const float*)W(n + ".exp.b");
      es.w_dw = (const float*)W(n + ".dw.w"); es.b_dw = (const float*)W(n + ".dw.b");
      es.se_partial = bb.se_partial;
      es.dbg_exp = keep_all_ ? (__half*)bb.exp.p : nullptr;
      es.B = b; es.H = x.H; es.W = x.W; es.Ho = bb.dw.H; es.Wo = bb.dw.W; es.cin = bs.cin; es.cexp = bs.cin * bs.e;
      es.k = bs.k; es.stride = bs.s; es.pad = lo;
      int tiles = 0;
      auto launch = make_expdw_launcher(es, owned, &tiles);
      if (launch) {
        Step s;
        s.name = n + ".expdw";
        s.kernel = "expdw_kernel";
        s.launch = launch;
        const double px = (double)b * x.H * x.W, pxo = (double)b * bb.dw.H * bb.dw.W;
        s.bytes = px * bs.cin * sT + pxo * es.cexp * sT + (double)es.cexp * bs.cin * sT + (double)es.cexp * (bs.k * bs.k + 2) * 4 +
                  (double)b * tiles * es.cexp * 4;
        s.flops = 2.0 * px * bs.cin * es.cexp + 2.0 * pxo * es.cexp * bs.k * bs.k;
        steps.push_back(s);
        last_dw_tiles = tiles;
        did_expdw = true;
      }
    }
    Tens e = x;
    if (bs.e != 1 && !did_expdw) {
      add_gemm(n + ".expand", {gemm_prob(x, n + ".exp.w", n + ".exp.b", bs.cin * bs.e, ACT_SWISH, bb.exp.p)});
      e = bb.exp;
    }
    // small maps: either se3 scales the depthwise output in place (HMDPOSE_SE_INPLACE), or -- default -- the project
    // GEMM applies the gate to its A tiles with all 320 non-producer threads (single-tile CTAs, gemm_tc.cuh)
    se_inplace = !v1_ && std::getenv("HMDPOSE_SE2") == nullptr && bb.dw.H * bb.dw.W <= 256 &&
                 (std::getenv("HMDPOSE_SE_INPLACE") != nullptr || force_simt_);   // both tcgen05 GEMMs gate their A tiles
    // HMDPOSE_SE_FOLD=1: squeeze-excite folded into the tail of the depthwise kernel (last block per image, blocks
    // 0..5 where the FC layers are tiny).  Measured: the serial tail costs what the saved se3 launch cost (1.166 vs
    // 1.158 ms per step), so the separate launch stays the default.
    const bool se_fold = !did_expdw && !v1_ && !se_inplace && std::getenv("HMDPOSE_SE2") == nullptr &&
                         std::getenv("HMDPOSE_DW2") == nullptr && std::getenv("HMDPOSE_SE_FOLD") != nullptr &&
                         bb.dw.C * std::max(1, bs.cin / 4) <= 2400;
    if (!did_expdw) {
      DwGroup dg = dw_group(e, bb.dw, n + ".dw.w", (const float*)W(n + ".dw.b"), bb.se_partial, bs.k, bs.s, ACT_SWISH);
      if (se_fold) {
        dg.se_counter = se_counters_ + (size_t)i * mb_;
        dg.se_wr = (const float*)W(n + ".se_r.w"); dg.se_br = (const float*)W(n + ".se_r.b");
        dg.se_weT = (const float*)W(n + ".se_e.wT"); dg.se_be = (const float*)W(n + ".se_e.b");
        dg.se_gate = bb.gate; dg.se_cse = std::max(1, bs.cin / 4);
        dg.se_inv_hw = 1.0f / (float)(bb.dw.H * bb.dw.W);
      }
      add_dw(n + ".dw", {dg});
    }
    if (!se_fold) {
      const int C = bb.dw.C, Cse = std::max(1, bs.cin / 4), tiles = last_dw_tiles;
      const float inv = 1.0f / (float)(bb.dw.H * bb.dw.W);
      const float *partial = bb.se_partial, *wr = (const float*)W(n + ".se_r.w"), *br = (const float*)W(n + ".se_r.b"),
                  *we = (const float*)W(n + ".se_e.w"), *be = (const float*)W(n + ".se_e.b");
      float* gate = bb.gate;
      const float* weT = (const float*)W(n + ".se_e.wT");
      Step s;
      s.name = n + ".se";
      if (v1_) {
        s.kernel = "se_kernel";
        s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(se_kernel<T>, dim3(b), dim3(256), 0, st, partial, tiles, C, Cse, inv, wr, br, we, be, gate)); };
      } else if (std::getenv("HMDPOSE_SE2") != nullptr) {
        s.kernel = "se2_kernel";
        s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(se2_kernel<T>, dim3(b), dim3(SE2_THREADS), 0, st, partial, tiles, C, Cse, inv, wr, br, weT, be, gate)); };
      } else {
        s.kernel = "se3_kernel";   // cluster of 8 CTAs per image (cluster dims are a kernel attribute)
        T* xs = se_inplace ? (T*)bb.dw.p : nullptr;
        const int hw = bb.dw.H * bb.dw.W;
        // large maps: the kernel also writes W'[img] = W_proj . diag(gate[img]); the project GEMM below then runs un-gated
        // (not when the split-K kernel of the latency plans takes the project convolution: it gates its own A k-blocks)
        bool want_projk = false;
        if (projk_ && b <= projk_max_batch_) {
          PkSpec ps;
          std::memset(&ps, 0, sizeof(ps));
          ps.M = b * hw; ps.N = bs.cout; ps.K = C; ps.rows_per_img = hw;
          want_projk = projk_fits(ps, mb_part_bytes_);
        }
        use_wgate = wgate_ && fast_ && !se_inplace && !force_simt_ && !want_projk && bb.wgated != nullptr &&
                    std::is_same<T, __half>::value;
        const T* wproj = use_wgate ? (const T*)W(n + ".proj.w") : nullptr;
        T* wg = use_wgate ? (T*)bb.wgated : nullptr;
        const int wN = bs.cout;
        s.launch = [=](cudaStream_t st) { HP_CUDA(launch_k(se3_kernel<T>, dim3(b * SE3_CL), dim3(SE3_THREADS), 0, st, partial, tiles, C, Cse, inv, wr, br, weT, be, gate, xs, hw, wproj, wg, wN)); };
        if (se_inplace) s.bytes += 2.0 * b * hw * C * sT;
      }
      s.bytes += (double)b * tiles * C * 4 + 2.0 * C * Cse * 4 + (double)b * C * 4;
      if (use_wgate) s.bytes += (1.0 + b) * bs.cout * C * sT;
      s.flops = 4.0 * b * C * Cse;
      steps.push_back(s);
    }
    GemmProb pj = gemm_prob(bb.dw, n + ".proj.w", n + ".proj.b", bs.cout, ACT_NONE, bb.out.p);
    pj.a_scale = se_inplace ? nullptr : bb.gate;
    if (use_wgate) { pj.a_scale = nullptr; pj.W = bb.wgated; pj.w_img_rows = bs.cout; }
    pj.residual = bs.skip ? x.p : nullptr;
    // deep-K project convolutions of the small maps: split-K over a cluster (projk_tc.cuh) instead of one CTA per row panel
    bool did_projk = false;
    // (latency plans only, batch <= projk_max_batch_: at batch 16 it shortens the single-stream step by 45 us, 1.04 -> 1.00 ms,
    // but its 48-128 fat CTAs cost 99 instead of 82 us per step with 8 steps in flight, 26.8 k against 27.9 k frames/s)
    if (std::is_same<T, __half>::value && fast_ && !v1_ && !force_simt_ && !se_inplace && projk_ && b <= projk_max_batch_) {
      PkSpec ps;
      std::memset(&ps, 0, sizeof(ps));
      ps.a = (const __half*)bb.dw.p; ps.gate = bb.gate; ps.bias = (const float*)W(n + ".proj.b");
      ps.residual = bs.skip ? (const __half*)x.p : nullptr; ps.out = (__half*)bb.out.p;
      ps.M = b * bb.dw.H * bb.dw.W; ps.N = bs.cout; ps.K = bb.dw.C; ps.rows_per_img = bb.dw.H * bb.dw.W;
      auto launch = make_projk_launcher(ps, W(n + ".proj.w"), owned, mb_part_, mb_part_bytes_);
      if (launch) {
        Step s;
        s.name = n + ".project";
        s.kernel = "projk_kernel";
        s.launch = launch;
        s.bytes = (double)ps.M * ps.K * sT + (double)ps.M * ps.N * sT * (bs.skip ? 2.0 : 1.0) + (double)ps.N * ps.K * sT +
                  (double)b * ps.K * 4 + ps.N * 4.0;
        s.flops = 2.0 * ps.M * ps.N * ps.K;
        steps.push_back(s);
        did_projk = true;
      }
    }
    if (!did_projk) add_gemm(n + ".project", {pj});
    x = bb.out;
  }
  const Tens P3 = blk_[4].out, P4 = blk_[10].out, P5 = blk_[15].out;  // efficientdet/model.py:452-457

  // ---- BiFPN x3 (efficientdet/model.py:194-266) ----
  auto add_fuse = [&](const std::string& name, const Tens& a, const Tens* bt, int mb, const Tens* ct, int mc,
                      const float* w, const Tens& out) {
    FuseArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.a = a.p; fa.b = bt ? bt->p : nullptr; fa.c = ct ? ct->p : nullptr; fa.out = out.p;
    fa.B = b; fa.H = a.H; fa.W = a.W; fa.C = a.C; fa.mode_b = bt ? mb : RS_NONE; fa.mode_c = ct ? mc : RS_NONE;
    fa.w0 = w[0]; fa.w1 = w[1]; fa.w2 = ct ? w[2] : 0.f;
    const long long total = (long long)b * a.H * a.W * (a.C / VecN<T>::N);
    const int blocks = (int)((total + 255) / 256);
    Step s{name, [=](cudaStream_t st) { HP_CUDA(launch_k(fuse_kernel<T>, dim3(blocks), dim3(256), 0, st, fa)); }, "fuse_kernel"};
    auto rs_elems = [&](int m) { return m == RS_UP2 ? 0.25 : (m == RS_POOL ? 4.0 : (m == RS_SAME ? 1.0 : 0.0)); };
    s.bytes = (double)b * a.H * a.W * a.C * sT * (2.0 + rs_elems(fa.mode_b) + rs_elems(fa.mode_c));
    steps.push_back(s);
  };
  auto add_pool = [&](const std::string& name, const Tens& src, const Tens& out) {
    const long long total = (long long)b * out.H * out.W * (out.C / VecN<T>::N);
    const int blocks = (int)((total + 255) / 256);
    const T* s = (const T*)src.p; T* o = (T*)out.p;
    const int H = out.H, Wd = out.W, C = out.C;
    Step stp{name, [=](cudaStream_t st) { HP_CUDA(launch_k(pool_kernel<T>, dim3(blocks), dim3(256), 0, st, s, o, b, H, Wd, C)); }, "pool_kernel"};
    stp.bytes = (double)b * H * Wd * C * sT * 5.0;
    steps.push_back(stp);
  };
  // consecutive nodes of the small levels (<= 128 pixels per `chain_nb` images) run as ONE chain launch: a CTA owns
  // chain_nb images and walks the nodes back to back (sepconv_tc.cuh)
  const int chain_nb = std::getenv("HMDPOSE_CHAIN_NB") ? std::max(1, std::atoi(std::getenv("HMDPOSE_CHAIN_NB"))) : 1;   // one image per CTA: the chain is latency-bound, more CTAs shorten every phase
  const bool use_chain = use_sep && std::getenv("HMDPOSE_NO_CHAIN") == nullptr;
  std::vector<SepSpec> chain_specs;
  std::string chain_name;
  auto flush_chain = [&]() {
    if (chain_specs.empty()) return;
    if (chain_specs.size() == 1) {
      add_sep(chain_name + ".sepconv", chain_specs);
    } else {
      Step s;
      s.name = chain_name + ".chain";
      s.kernel = "sepconv_kernel";
      for (const SepSpec& q : chain_specs) {
        auto rs_elems = [&](int m) { return m == RS_UP2 ? 0.25 : (m == RS_POOL ? 4.0 : (m == RS_SAME ? 1.0 : 0.0)); };
        const double px = (double)b * q.H * q.W;
        s.bytes += px * 64 * sT * (2.0 + rs_elems(q.mode_b) + rs_elems(q.mode_c)) + 64 * 64 * sT + 64 * 4.0 + 9 * 64 * 4.0;
        s.flops += 2.0 * px * 64 * (9.0 + 64);
      }
      s.launch = make_sepconv_chain_launcher(chain_specs, chain_nb, owned);
      steps.push_back(s);
    }
    chain_specs.clear();
    chain_name.clear();
  };
  Tens feat[5];
  for (int c = 0; c < 3; ++c) {
    CellBufs& cb = cell_[c];
    const std::string cn = "bifpn" + std::to_string(c);
    Tens in[5];
    if (c == 0) {
      add_gemm("bifpn0.proj", {gemm_prob(P3, "bifpn0.p3_dc.w", "bifpn0.p3_dc.b", 64, ACT_NONE, cb.in[0].p),
                               gemm_prob(P4, "bifpn0.p4_dc.w", "bifpn0.p4_dc.b", 64, ACT_NONE, cb.in[1].p),
                               gemm_prob(P5, "bifpn0.p5_dc.w", "bifpn0.p5_dc.b", 64, ACT_NONE, cb.in[2].p),
                               gemm_prob(P5, "bifpn0.p5_to_p6.w", "bifpn0.p5_to_p6.b", 64, ACT_NONE, cb.p6_pre.p),
                               gemm_prob(P4, "bifpn0.p4_dc2.w", "bifpn0.p4_dc2.b", 64, ACT_NONE, cb.in2[0].p),
                               gemm_prob(P5, "bifpn0.p5_dc2.w", "bifpn0.p5_dc2.b", 64, ACT_NONE, cb.in2[1].p)});
      add_pool("bifpn0.p6_in", cb.p6_pre, cb.in[3]);
      add_pool("bifpn0.p7_in", cb.in[3], cb.in[4]);
      for (int l = 0; l < 5; ++l) in[l] = cb.in[l];
    } else {
      for (int l = 0; l < 5; ++l) in[l] = feat[l];
    }
    auto node = [&](int ni, int l, const Tens& a, const Tens* bt, int mb, const Tens* ct, int mc, const Tens& out) {
      const std::string q = cn + "." + kNodeNames[ni];
      const HostTensor& fw = blob_.get(cn + ".fw." + kFwNames[ni]);
      if (use_sep) {
        SepSpec sq = sep_spec(a, q + ".dw.w", q + ".pw.w", q + ".pw.b", 64, ACT_NONE, out.p);
        sq.fused = 1; sq.fb = bt ? bt->p : nullptr; sq.fc = ct ? ct->p : nullptr;
        sq.mode_b = bt ? mb : RS_NONE; sq.mode_c = ct ? mc : RS_NONE;
        sq.w0 = fw.data[0]; sq.w1 = fw.data[1]; sq.w2 = ct ? fw.data[2] : 0.f;
        if (use_chain && a.H * a.W * chain_nb <= 128) {
          chain_name += (chain_name.empty() ? "" : "+") + q;
          chain_specs.push_back(sq);
          return;
        }
        flush_chain();
        add_sep(q + ".sepconv", {sq});
        return;
      }
      if (v1_) {
        add_fuse(q + ".fuse", a, bt, mb, ct, mc, fw.data, cb.fused[l]);
        add_dw(q + ".dw", {dw_group(cb.fused[l], cb.dwb[l], q + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE)});
      } else {
        DwGroup g = dw_group(a, cb.dwb[l], q + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE);
        g.fb = bt ? bt->p : nullptr; g.fc = ct ? ct->p : nullptr;
        g.mode_b = bt ? mb : RS_NONE; g.mode_c = ct ? mc : RS_NONE;
        g.w0 = fw.data[0]; g.w1 = fw.data[1]; g.w2 = ct ? fw.data[2] : 0.f;
        add_dw(q + ".fuse_dw", {g}, true);
      }
      add_gemm(q + ".pw", {gemm_prob(cb.dwb[l], q + ".pw.w", q + ".pw.b", 64, ACT_NONE, out.p)});
    };
    // top-down: P6_up, P5_up, P4_up, P3_out
    node(0, 3, in[3], &in[4], RS_UP2, nullptr, 0, cb.up[3]);
    node(1, 2, in[2], &cb.up[3], RS_UP2, nullptr, 0, cb.up[2]);
    node(2, 1, in[1], &cb.up[2], RS_UP2, nullptr, 0, cb.up[1]);
    node(3, 0, in[0], &cb.up[1], RS_UP2, nullptr, 0, cb.out[0]);
    // bottom-up (cell 0 re-projects P4/P5 with the *_down_channel_2 weights, model.py:235-237)
    const Tens in4 = c == 0 ? cb.in2[0] : in[1];
    const Tens in5 = c == 0 ? cb.in2[1] : in[2];
    node(4, 1, in4, &cb.up[1], RS_SAME, &cb.out[0], RS_POOL, cb.out[1]);
    node(5, 2, in5, &cb.up[2], RS_SAME, &cb.out[1], RS_POOL, cb.out[2]);
    node(6, 3, in[3], &cb.up[3], RS_SAME, &cb.out[2], RS_POOL, cb.out[3]);
    node(7, 4, in[4], &cb.out[3], RS_POOL, nullptr, 0, cb.out[4]);
    for (int l = 0; l < 5; ++l) feat[l] = cb.out[l];
  }
  flush_chain();

  // ---- heads (efficientdet/model.py:361-417, hmdegopose/model.py:55-228): 5 heads x 5 levels per launch ----
  if (mode != PLAN_D0 && num_heads_ < 5)
    throw Error(HMDPOSE_E_STATE, "detector-only weights (no rotation/translation/hand sub-nets): use hmdpose_run_d0");
  const int nheads = mode == PLAN_D0 ? 2 : 5;
  for (int i = 0; i < 3; ++i) {
    std::vector<DwGroup> dg;
    std::vector<GemmProb> gp;
    std::vector<SepSpec> sps;
    for (int h = 0; h < nheads; ++h)
      for (int l = 0; l < 5; ++l) {
        const std::string p = std::string("head.") + kHeadNames[h] + ".l" + std::to_string(i);
        const Tens& src = i == 0 ? feat[l] : trunk_[h][l][trunk_slot(i - 1)];
        if (use_sep) {
          const std::string ql = p + ".lvl" + std::to_string(l);
          SepSpec sq = sep_spec(src, p + ".dw.w", ql + ".pw.w", ql + ".pw.b", 64, ACT_SWISH, trunk_[h][l][trunk_slot(i)].p);
          if (blob_.has(p + ".pw.raw") && blob_.has(ql + ".pw.scale")) {
            // one set of tap matrices for the five levels; the per-level BN scale moves to the epilogue
            sq.w9 = w9_for(p + ".dw.w", p + ".pw.raw");
            auto it = wdev_.find(ql + ".pw.scale");
            if (it == wdev_.end()) {
              const HostTensor& t = blob_.get(ql + ".pw.scale");
              it = wdev_.emplace(ql + ".pw.scale", upload_f32(t.data, t.count)).first;
            }
            sq.scale = (const float*)it->second;
          }
          sps.push_back(sq);
          continue;
        }
        dg.push_back(dw_group(src, hdw_[h][l], p + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE));
        const std::string q = p + ".lvl" + std::to_string(l);
        gp.push_back(gemm_prob(hdw_[h][l], q + ".pw.w", q + ".pw.b", 64, ACT_SWISH, trunk_[h][l][trunk_slot(i)].p));
      }
    if (use_sep) { add_sep("heads.l" + std::to_string(i) + ".sepconv", sps); continue; }
    add_dw("heads.l" + std::to_string(i) + ".dw", dg);
    add_gemm("heads.l" + std::to_string(i) + ".pw", gp);
  }
  {
    const int C = cfg.num_classes;
    struct Hdr { int head, j, cout, p_src, p_dst, p_off, act; float* out; };
    const Hdr hdrs[6] = {{0, 0, 36, 4, 4, 0, ACT_NONE, o_reg_},        {1, 0, 9 * C, C, C, 0, ACT_SIGMOID, o_cls_},
                         {2, 0, 27, 3, 3, 0, ACT_NONE, o_rot_},        {3, 0, 18, 2, 3, 0, ACT_NONE, o_traw_},
                         {3, 1, 9, 1, 3, 2, ACT_NONE, o_traw_},        {4, 0, 567, 63, 63, 0, ACT_NONE, o_hand_}};
    std::vector<DwGroup> dg;
    std::vector<GemmProb> gp;
    std::vector<SepSpec> sps;
    // EfficientDet-d0 plans: max / arg-max over the classes and the score threshold inside the classifier header's
    // epilogue -- the (B, N, C) score tensor (17.7 MB per 512x512 frame at 90 classes) is never written
    d0_fused = mode == PLAN_D0 && use_sep && std::getenv("HMDPOSE_D0_DENSE") == nullptr;
    const D0Args d0a = mode == PLAN_D0 ? d0_args() : D0Args();
    // detection / best-pose plans evaluate the hand header only at the kept anchors (hand_gather_kernel)
    const bool full_hand = full_hand_for(mode);
    // ... and, single-class detection plans, the rotation / translation headers too (pose_gather_kernel)
    for (int k = 0; k < (mode == PLAN_D0 || gather_pose_for(mode) ? 2 : (full_hand ? 6 : 5)); ++k)
      for (int l = 0; l < 5; ++l) {
        const Hdr& hd = hdrs[k];
        const std::string p = std::string("head.") + kHeadNames[hd.head] + ".hdr" + std::to_string(hd.j);
        const Tens& src = trunk_[hd.head][l][trunk_final()];  // after 3 layers the trunk output sits in buffer 0
        if (use_sep) {
          SepSpec sq = sep_spec(src, p + ".dw.w", p + ".pw.w", p + ".pw.b", hd.cout, hd.act,
                                hd.out + (size_t)lvl_off_[l] * hd.p_dst);
          sq.p.out_mode = 1; sq.p.p_src = hd.p_src; sq.p.p_dst = hd.p_dst; sq.p.p_off = hd.p_off;
          sq.p.pix_stride = 9 * hd.p_dst; sq.p.img_stride = (long long)N * hd.p_dst;
          if (d0_fused && k == 1) {
            sq.p.out_mode = 2;
            sq.p.d0_keys = d0a.keys; sq.p.d0_cand_cls = d0a.cand_cls; sq.p.d0_cand_count = d0a.cand_count;
            sq.p.d0_cap = d0a.cap; sq.p.d0_ntot = N; sq.p.d0_anchor0 = lvl_off_[l]; sq.p.d0_thr = d0a.threshold;
          }
          sps.push_back(sq);
          continue;
        }
        dg.push_back(dw_group(src, hdrdw_[k][l], p + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE));
        GemmProb g = gemm_prob(hdrdw_[k][l], p + ".pw.w", p + ".pw.b", hd.cout, hd.act,
                               hd.out + (size_t)lvl_off_[l] * hd.p_dst);
        g.out_mode = 1; g.p_src = hd.p_src; g.p_dst = hd.p_dst; g.p_off = hd.p_off;
        g.pix_stride = 9 * hd.p_dst; g.img_stride = (long long)N * hd.p_dst;
        gp.push_back(g);
      }
    if (d0_fused) {   // candidate counters: zero before the classifier header appends to them
      int* cc = d0a.cand_count;
      Step z{"post.d0_zero", [=](cudaStream_t st) { HP_CUDA(cudaMemsetAsync(cc, 0, sizeof(int) * b, st)); }, "memset"};
      steps.push_back(z);
    }
    if (use_sep) add_sep("heads.hdr.sepconv", sps);
    else {
      add_dw("heads.hdr.dw", dg);
      add_gemm("heads.hdr.pw", gp);
    }
  }
  // ---- --iter 1: one refinement step of rotation / translation / hand (hmdegopose/model.py:76-82,147-152,214-220,
  // 232-346): estimate += head( swish(BN(sepconv(cat(feat, estimate)))) ).  The hand refinement only matters when hand
  // coordinates are returned (raw and detection plans).
  if (iter1_ && mode != PLAN_D0) {
    const bool full_hand = mode == PLAN_RAW || (mode & PLAN_DET);
    const int nt = full_hand ? 3 : 2;
    struct ItHead { float* est; int pw_; };   // estimate tensor, row width
    const ItHead ih[3] = {{o_rot_, 3}, {o_traw_, 3}, {o_hand_, HMDPOSE_NUM_HAND}};
    // (1) cat(feat, estimate) -> NHWC
    {
      std::vector<ConcatProb> cps;
      int blocks = 0;
      for (int t = 0; t < nt; ++t)
        for (int l = 0; l < 5; ++l) {
          ConcatProb cp;
          std::memset(&cp, 0, sizeof(cp));
          cp.feat = trunk_[2 + t][l][trunk_final()].p;
          cp.head = ih[t].est + (size_t)lvl_off_[l] * ih[t].pw_;
          cp.out = it_in_[t][l].p;
          cp.HW = lvl_hw_[l]; cp.npix = b * lvl_hw_[l]; cp.Cpad = it_in_[t][l].C; cp.P = 9 * ih[t].pw_;
          cp.trans = t == 1; cp.img_stride = (long long)N * ih[t].pw_;
          cp.blk_start = blocks;
          blocks += cdiv(cp.npix * (cp.Cpad / VecN<T>::N), 256);
          cps.push_back(cp);
        }
      ConcatProb* d = nullptr;
      HP_CUDA(cudaMalloc(&d, sizeof(ConcatProb) * cps.size()));
      HP_CUDA(cudaMemcpy(d, cps.data(), sizeof(ConcatProb) * cps.size(), cudaMemcpyHostToDevice));
      owned.push_back(d);
      const int ncp = (int)cps.size();
      Step s{"heads.it.concat", [=](cudaStream_t st) { HP_CUDA(launch_k(concat_kernel<T>, dim3(blocks), dim3(256), 0, st, d, ncp)); }, "concat_kernel"};
      for (const ConcatProb& cp : cps) s.bytes += (double)cp.npix * (64 * sT + cp.P * 4.0 + cp.Cpad * sT);
      steps.push_back(s);
    }
    // (2) depthwise 3x3 over the (64 + P) channels, (3) pointwise -> 64 + norm_layer[0][0] + swish
    {
      std::vector<DwGroup> dg;
      std::vector<GemmProb> gp;
      for (int t = 0; t < nt; ++t)
        for (int l = 0; l < 5; ++l) {
          const std::string p = std::string("head.") + kHeadNames[2 + t] + ".it";
          dg.push_back(dw_group(it_in_[t][l], it_dw_[t][l], p + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE));
          gp.push_back(gemm_prob(it_dw_[t][l], p + ".pw.w", p + ".pw.b", 64, ACT_SWISH, it_y_[t][l].p));
        }
      add_dw("heads.it.dw", dg);
      add_gemm("heads.it.pw", gp);
    }
    // (4) refinement heads, accumulated into the estimates with the scatter of the initial headers
    {
      struct Hdr { int t, j, cout, p_src, p_dst, p_off; float* out; };
      const Hdr hdrs[4] = {{0, 0, 27, 3, 3, 0, o_rot_}, {1, 0, 18, 2, 3, 0, o_traw_}, {1, 1, 9, 1, 3, 2, o_traw_},
                           {2, 0, 567, 63, 63, 0, o_hand_}};
      std::vector<DwGroup> dg;
      std::vector<GemmProb> gp;
      std::vector<SepSpec> sps;
      for (int k = 0; k < (full_hand ? 4 : 3); ++k)
        for (int l = 0; l < 5; ++l) {
          const Hdr& hd = hdrs[k];
          const std::string p = std::string("head.") + kHeadNames[2 + hd.t] + ".it.hdr" + std::to_string(hd.j);
          const Tens& src = it_y_[hd.t][l];
          float* out = hd.out + (size_t)lvl_off_[l] * hd.p_dst;
          if (use_sep) {
            SepSpec sq = sep_spec(src, p + ".dw.w", p + ".pw.w", p + ".pw.b", hd.cout, ACT_NONE, out);
            sq.p.out_mode = 1; sq.p.p_src = hd.p_src; sq.p.p_dst = hd.p_dst; sq.p.p_off = hd.p_off;
            sq.p.pix_stride = 9 * hd.p_dst; sq.p.img_stride = (long long)N * hd.p_dst;
            sq.p.accumulate = 1;
            sps.push_back(sq);
            continue;
          }
          dg.push_back(dw_group(src, it_hdw_[k][l], p + ".dw.w", nullptr, nullptr, 3, 1, ACT_NONE));
          GemmProb g = gemm_prob(it_hdw_[k][l], p + ".pw.w", p + ".pw.b", hd.cout, ACT_NONE, out);
          g.out_mode = 1; g.p_src = hd.p_src; g.p_dst = hd.p_dst; g.p_off = hd.p_off;
          g.pix_stride = 9 * hd.p_dst; g.img_stride = (long long)N * hd.p_dst;
          g.accumulate = 1;
          gp.push_back(g);
        }
      if (use_sep) add_sep("heads.it.hdr.sepconv", sps);
      else {
        add_dw("heads.it.hdr.dw", dg);
        add_gemm("heads.it.hdr.pw", gp);
      }
    }
  }
  if (mode == PLAN_D0) {
    const D0Args da = d0_args();
    const bool fused = d0_fused;
    Step s{"post.d0", [=](cudaStream_t st) { if (fused) launch_d0_nms(da, b, st); else launch_d0(da, b, st); }, "d0_nms_kernel"};
    s.bytes = fused ? (double)b * N * 8 : (double)b * N * cfg.num_classes * 4;
    steps.push_back(s);
    return plan;
  }
  add_post_steps(steps, b, mode, true, true, full_hand_for(mode), gather_pose_for(mode));
  if (gather_pose_for(mode)) {
    PoseGatherArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    for (int l = 0; l < 5; ++l) {
      for (int t = 0; t < 3; ++t) pa.trunk[t][l] = trunk_[2 + t][l][trunk_final()].p;
      pa.side[l] = lvl_side_[l]; pa.lvl_off[l] = lvl_off_[l];
    }
    pa.lvl_off[5] = N;
    static const char* kHdr[4] = {"head.rot.hdr0", "head.trans.hdr0", "head.trans.hdr1", "head.hand.hdr0"};
    for (int h = 0; h < 4; ++h) {
      pa.dw_w[h] = (const float*)W(std::string(kHdr[h]) + ".dw.w");
      pa.pw_w[h] = W(std::string(kHdr[h]) + ".pw.w");
      pa.bias[h] = (const float*)W(std::string(kHdr[h]) + ".pw.b");
    }
    pa.tanchors = d_tanchors_; pa.cam = d_cam_local_;
    pa.det_idx = det_idx_; pa.det_rot = det_rot_; pa.det_trans = det_trans_; pa.det_hand = det_hand_;
    pa.B = b; pa.D = cfg.max_detections;
    const int blocks = cdiv(b * cfg.max_detections, 8);
    Step s{"post.pose_gather", [=](cudaStream_t st) { HP_CUDA(launch_k(pose_gather_kernel<T>, dim3(blocks), dim3(256), 0, st, pa)); }, "pose_gather_kernel"};
    s.bytes = (double)b * cfg.max_detections * ((63 + 6) * 4 + 3 * 9 * 64 * sT) + (567 + 54) * 64 * sT;
    s.flops = 2.0 * b * cfg.max_detections * 64 * (4 * 9 + 63 + 6);
    steps.push_back(s);
  } else if ((mode & PLAN_DET) && !full_hand_for(mode)) {
    HandGatherArgs ha;
    std::memset(&ha, 0, sizeof(ha));
    for (int l = 0; l < 5; ++l) { ha.trunk[l] = trunk_[4][l][trunk_final()].p; ha.side[l] = lvl_side_[l]; ha.lvl_off[l] = lvl_off_[l]; }
    ha.lvl_off[5] = N;
    ha.dw_w = (const float*)W("head.hand.hdr0.dw.w");
    ha.pw_w = W("head.hand.hdr0.pw.w");
    ha.bias = (const float*)W("head.hand.hdr0.pw.b");
    ha.det_idx = det_idx_; ha.det_hand = det_hand_; ha.B = b; ha.D = cfg.max_detections;
    const int blocks = cdiv(b * cfg.max_detections, 8);
    Step s{"post.hand_gather", [=](cudaStream_t st) { HP_CUDA(launch_k(hand_gather_kernel<T>, dim3(blocks), dim3(256), 0, st, ha)); }, "hand_gather_kernel"};
    s.bytes = (double)b * cfg.max_detections * (63 * 4 + 9 * 64 * sT) + 567 * 64 * sT;
    s.flops = 2.0 * b * cfg.max_detections * 64 * (9 + 63);
    steps.push_back(s);
  }
  return plan;
}

// Post-processing steps on the micro-batch-local head tensors (o_*_), camera rows in d_cam_local_.
void Engine::add_post_steps(std::vector<Step>& steps, int b, int mode, bool decode_boxes, bool decode_trans,
                            bool hand_from_raw, bool pose_from_gather) {
  const int S = cfg.image_size, C = cfg.num_classes, D = cfg.max_detections, Nn = N;
  const float thr = cfg.score_threshold, iou = cfg.iou_threshold;
  if ((mode & PLAN_DET) && C == 1 && !post_v1_) {
    // single class: one fused kernel (threshold + candidate decode + NMS + translation recovery + gather)
    FilterArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.boxes = decode_boxes ? nullptr : p_boxes_;
    fa.anchors = d_anchors_; fa.reg = o_reg_; fa.wmax = (float)(S - 1); fa.hmax = (float)(S - 1);
    fa.scores = o_cls_; fa.rotation = o_rot_;
    fa.translation = decode_trans ? nullptr : p_trans_;
    fa.tanchors = d_tanchors_; fa.traw = o_traw_; fa.cam = d_cam_local_;
    fa.hand = hand_from_raw ? o_hand_ : nullptr;
    fa.N = Nn; fa.H = HMDPOSE_NUM_HAND; fa.cap = pb_.cap; fa.max_det = D; fa.score_thr = thr; fa.iou_thr = iou;
    fa.keys = pb_.keys;
    fa.o_boxes = det_boxes_; fa.o_scores = det_scores_; fa.o_labels = det_labels_;
    fa.o_rot = pose_from_gather ? nullptr : det_rot_;       // pose_gather_kernel writes the pose rows of the kept anchors
    fa.o_trans = pose_from_gather ? nullptr : det_trans_;
    fa.o_hand = hand_from_raw ? det_hand_ : nullptr; fa.o_idx = det_idx_;
    Step s{"post.filter_fused", [=](cudaStream_t st) { launch_filter_fused(fa, b, st); }, "filter_fused_kernel"};
    s.bytes = (double)b * Nn * 4 + (double)b * D * 12 * 4 * 2;
    steps.push_back(s);
  } else if (mode & PLAN_DET) {
    if (decode_boxes) {
      Step s{"post.decode_boxes", [=](cudaStream_t st) { launch_decode_boxes(d_anchors_, o_reg_, b, Nn, S, S, p_boxes_, st); },
             "decode_boxes_kernel"};
      s.bytes = (double)b * Nn * 32 + Nn * 16.0;
      steps.push_back(s);
    }
    if (decode_trans) {
      Step s{"post.decode_translation",
             [=](cudaStream_t st) { launch_decode_translation(d_tanchors_, o_traw_, d_cam_local_, b, Nn, p_trans_, st); },
             "decode_translation_kernel"};
      s.bytes = (double)b * Nn * 24 + Nn * 12.0;
      steps.push_back(s);
    }
    {
      const PostBuffers pb = pb_;
      Step s{"post.filter_nms", [=](cudaStream_t st) {
               launch_filter_nms(pb, p_boxes_, o_cls_, b, Nn, C, thr, iou, D, st);
             }, "filter_nms_kernel"};
      s.bytes = (double)b * Nn * C * 4;
      steps.push_back(s);
      Step g{"post.topk_gather", [=](cudaStream_t st) {
               launch_topk_gather(pb, p_boxes_, o_rot_, p_trans_, o_hand_, b, Nn, C, HMDPOSE_NUM_HAND, D, det_boxes_,
                                  det_scores_, det_labels_, det_rot_, det_trans_, hand_from_raw ? det_hand_ : nullptr,
                                  det_idx_, st);
             }, "topk_gather_kernel"};
      g.bytes = (double)b * D * (4 + 1 + 1 + 3 + 3 + HMDPOSE_NUM_HAND + 1) * 4 * 2;
      steps.push_back(g);
    }
  }
  if (mode & PLAN_BEST) {
    Step s{"post.best", [=](cudaStream_t st) {
             launch_best(d_anchors_, d_tanchors_, o_reg_, o_cls_, o_rot_, o_traw_, d_cam_local_, b, Nn, C, thr, S, S,
                         d_best_, st);
           }, "best_kernel"};
    s.bytes = (double)b * Nn * C * 4;
    steps.push_back(s);
  }
}

template <typename T>
Plan* Engine::get_plan(int b, int mode) {
  const int key = b * 8 + mode;
  auto it = plans_.find(key);
  if (it != plans_.end()) return it->second.get();
  std::unique_ptr<Plan> p = build_plan<T>(b, mode);
  Plan* raw = p.get();
  raw->launches = (int)raw->steps.size();
  if (cfg.use_graph && std::getenv("HMDPOSE_NO_GRAPH") == nullptr) {
    cudaStream_t cs;
    HP_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    HP_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    for (Step& s : raw->steps) s.launch(cs);
    cudaError_t e = cudaStreamEndCapture(cs, &raw->graph);
    cudaStreamDestroy(cs);
    if (e != cudaSuccess) throw Error(HMDPOSE_E_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
    HP_CUDA(cudaGraphInstantiate(&raw->exec, raw->graph, 0));
  }
  plans_[key] = std::move(p);
  return raw;
}

void Engine::wait_stream() {
  if (ev_block_) {
    HP_CUDA(cudaEventRecord(ev_block_, stream));
    HP_CUDA(cudaEventSynchronize(ev_block_));
  } else {
    HP_CUDA(cudaStreamSynchronize(stream));
  }
}

void Engine::enter(cudaStream_t st) {
  if (has_last_ && st != last_stream_) HP_CUDA(cudaStreamWaitEvent(st, ev_done_, 0));
}
void Engine::leave(cudaStream_t st) {
  HP_CUDA(cudaEventRecord(ev_done_, st));
  last_stream_ = st;
  has_last_ = true;
}
float Engine::last_gpu_ms() {
  if (timing_valid_) {
    cudaSetDevice(cfg.device);
    if (cudaEventSynchronize(ev1_) == cudaSuccess) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ev0_, ev1_) == cudaSuccess) last_ms = t;
    }
    cudaGetLastError();
  }
  return last_ms;
}

void Engine::run_plan(Plan* p, cudaStream_t st) {
  if (p->exec) {
    HP_CUDA(cudaGraphLaunch(p->exec, st));
  } else {
    for (Step& s : p->steps) {
      s.launch(st);
      if (keep_all_) {  // debug mode: attribute launch/execution failures to the step
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) throw Error(HMDPOSE_E_CUDA, "step " + s.name + ": " + cudaGetErrorString(e));
      }
    }
  }
  HP_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
Engine::Engine(const hmdpose_config_t& c, const void* blob, size_t bytes) : cfg(c) {
  if (cfg.image_size < 128 || cfg.image_size % 128 != 0)
    throw Error(HMDPOSE_E_ARG, "image_size must be a positive multiple of 128");
  if (cfg.max_batch < 1) throw Error(HMDPOSE_E_ARG, "max_batch must be >= 1");
  if (cfg.max_detections < 1 || cfg.max_detections > MAX_DET_CAP)
    throw Error(HMDPOSE_E_ARG, "max_detections must be in [1, 256]");
  if (cfg.precision != HMDPOSE_PRECISION_PARITY && cfg.precision != HMDPOSE_PRECISION_FAST)
    throw Error(HMDPOSE_E_ARG, "unknown precision mode");
  // the fused BiFPN / head kernels of the fast mode index their tiles with shifts: power-of-two feature maps only.
  // Parity mode runs any multiple of 128 (e.g. the reference's phi1 size 640).
  // ... and at least 256: at 128 the coarsest pyramid level is 1 x 1, 128 images with their halo rows overflow the
  // staging tile of sepconv_kernel (found by compute-sanitizer in round 2; parity mode runs 128 through the generic kernels)
  if (cfg.precision == HMDPOSE_PRECISION_FAST && ((cfg.image_size & (cfg.image_size - 1)) != 0 || cfg.image_size < 256))
    throw Error(HMDPOSE_E_ARG, "fast mode needs a power-of-two image_size >= 256 (256, 512, 1024); use the parity mode for " +
                                   std::to_string(cfg.image_size));
  blob_.parse(blob, bytes);
  if (cfg.num_classes <= 0) cfg.num_classes = blob_.num_classes;
  if (cfg.num_classes != blob_.num_classes)
    throw Error(HMDPOSE_E_WEIGHTS, "num_classes does not match the weight blob");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error(HMDPOSE_E_CUDA, "no CUDA device: libhmdpose has no CPU fallback");
  if (cfg.device < 0 || cfg.device >= ndev) throw Error(HMDPOSE_E_ARG, "bad device ordinal");
  HP_CUDA(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  HP_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
  if (prop.major != 10) throw Error(HMDPOSE_E_CUDA, "libhmdpose is built for sm_100a (Blackwell B200) only");
  try {
  trap_info_setup();
  fast_ = cfg.precision == HMDPOSE_PRECISION_FAST;
  keep_all_ = std::getenv("HMDPOSE_KEEP_ALL") != nullptr;
  mbfuse_ = std::getenv("HMDPOSE_MBFUSE") != nullptr;
  expdw_ = std::getenv("HMDPOSE_EXPDW") != nullptr;
  projk_ = std::getenv("HMDPOSE_NO_PROJK") == nullptr;
  wgate_ = std::getenv("HMDPOSE_NO_WGATE") == nullptr;
  if (const char* e = std::getenv("HMDPOSE_PROJK_MAX_BATCH")) projk_max_batch_ = std::atoi(e);
  v1_ = std::getenv("HMDPOSE_V1") != nullptr;
  gather_hand_off_ = std::getenv("HMDPOSE_FULL_HAND") != nullptr;
  dense_pose_ = std::getenv("HMDPOSE_DENSE_POSE") != nullptr;
  post_v1_ = std::getenv("HMDPOSE_POST_V1") != nullptr;
  force_simt_ = std::getenv("HMDPOSE_FORCE_SIMT") != nullptr;
  mb_ = cfg.micro_batch > 0 ? cfg.micro_batch : 16;
  mb_ = std::min(mb_, cfg.max_batch);
  N = anchors_count(cfg.image_size);
  compute_anchors(cfg.image_size, h_anchors, h_tanchors);
  HP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  HP_CUDA(cudaEventCreate(&ev0_));
  HP_CUDA(cudaEventCreate(&ev1_));
  HP_CUDA(cudaEventCreateWithFlags(&ev_done_, cudaEventDisableTiming));
  // HMDPOSE_BLOCKING_SYNC=1: host-API callers sleep on a blocking event instead of spinning in cudaStreamSynchronize
  // (for hosts with more caller threads than cores; costs a wake-up per call)
  if (std::getenv("HMDPOSE_BLOCKING_SYNC") != nullptr)
    HP_CUDA(cudaEventCreateWithFlags(&ev_block_, cudaEventBlockingSync | cudaEventDisableTiming));
  d_anchors_ = upload_f32(h_anchors.data(), h_anchors.size());
  d_tanchors_ = upload_f32(h_tanchors.data(), h_tanchors.size());
  if (fast_) { upload_weights<__half>(); alloc_buffers<__half>(); }
  else { upload_weights<float>(); alloc_buffers<float>(); }
  HP_CUDA(cudaDeviceSynchronize());
  } catch (...) {   // a partially constructed Engine never runs ~Engine: release what was acquired, then report
    release();
    throw;
  }
}

Engine::~Engine() { release(); }

void Engine::release() {
  cudaSetDevice(cfg.device);
  cudaDeviceSynchronize();
  plans_.clear();
  post_plans_.clear();
  for (void* p : allocs_) cudaFree(p);
  if (d_u8_) cudaFree(d_u8_);
  if (h_pinned_) cudaFreeHost(h_pinned_);
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  if (ev_block_) cudaEventDestroy(ev_block_);
  if (ev_done_) cudaEventDestroy(ev_done_);
  if (stream) cudaStreamDestroy(stream);
  allocs_.clear();
  d_u8_ = nullptr; h_pinned_ = nullptr; ev0_ = ev1_ = ev_block_ = ev_done_ = nullptr; stream = nullptr;
}

template <typename T>
Step Engine::stem_step(const float* d_in, long long sb, long long sc, long long sh, long long sw, int b) {
  const int S = cfg.image_size;
  const long long total = (long long)b * (S / 2) * (S / 2);
  const int blocks = (int)((total + 127) / 128);
  T* out = (T*)stem_out_.p;
  const StemW wb = stem_wb_;
  Step s{"stem", [=](cudaStream_t st) { HP_CUDA(launch_k(stem_kernel<T>, dim3(blocks), dim3(128), 0, st, d_in, sb, sc, sh, sw, b, S, wb, out)); },
         "stem_kernel"};
  s.bytes = (double)b * 3 * S * S * 4 + (double)b * (S / 2) * (S / 2) * 32 * sizeof(T);
  s.flops = 2.0 * 27 * 32 * b * (S / 2) * (S / 2);
  return s;
}

void Engine::run_device(const float* d_in, long long sb, long long sc, long long sh, long long sw, const float* d_cam,
                        int batch, bool want_raw, float* raw[5], bool want_det, float* d_boxes, float* d_scores,
                        int32_t* d_labels, float* d_rot, float* d_trans, float* d_hand, int32_t* d_idx,
                        bool want_best, float* d_best, cudaStream_t st) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!d_in) throw Error(HMDPOSE_E_ARG, "null input");
  if ((want_det || want_best) && !d_cam) throw Error(HMDPOSE_E_ARG, "null camera parameters");
  HP_CUDA(cudaSetDevice(cfg.device));
  if (!st) st = stream;
  const int C = cfg.num_classes, D = cfg.max_detections;
  const int mode = (want_det ? PLAN_DET : 0) | (want_best ? PLAN_BEST : 0);
  last_launches = 0;
  enter(st);
  HP_CUDA(cudaEventRecord(ev0_, st));
  for (int f0 = 0; f0 < batch; f0 += mb_) {
    const int b = std::min(mb_, batch - f0);
    Plan* plan = fast_ ? get_plan<__half>(b, mode) : get_plan<float>(b, mode);
    if (mode) HP_CUDA(cudaMemcpyAsync(d_cam_local_, d_cam + 6 * f0, (size_t)b * 24, cudaMemcpyDeviceToDevice, st));
    Step stem = fast_ ? stem_step<__half>(d_in + f0 * sb, sb, sc, sh, sw, b) : stem_step<float>(d_in + f0 * sb, sb, sc, sh, sw, b);
    stem.launch(st);
    run_plan(plan, st);
    last_launches += 1 + plan->launches;
    auto copy = [&](void* dst, const void* src, size_t bytes) {
      if (dst) HP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    };
    if (want_raw) {
      copy(raw[0] ? raw[0] + (size_t)f0 * N * 4 : nullptr, o_reg_, (size_t)b * N * 4 * 4);
      copy(raw[1] ? raw[1] + (size_t)f0 * N * C : nullptr, o_cls_, (size_t)b * N * C * 4);
      copy(raw[2] ? raw[2] + (size_t)f0 * N * 3 : nullptr, o_rot_, (size_t)b * N * 3 * 4);
      copy(raw[3] ? raw[3] + (size_t)f0 * N * 3 : nullptr, o_traw_, (size_t)b * N * 3 * 4);
      copy(raw[4] ? raw[4] + (size_t)f0 * N * HMDPOSE_NUM_HAND : nullptr, o_hand_, (size_t)b * N * HMDPOSE_NUM_HAND * 4);
    }
    if (want_det) {
      copy(d_boxes ? d_boxes + (size_t)f0 * D * 4 : nullptr, det_boxes_, (size_t)b * D * 4 * 4);
      copy(d_scores ? d_scores + (size_t)f0 * D : nullptr, det_scores_, (size_t)b * D * 4);
      copy(d_labels ? d_labels + (size_t)f0 * D : nullptr, det_labels_, (size_t)b * D * 4);
      copy(d_rot ? d_rot + (size_t)f0 * D * 3 : nullptr, det_rot_, (size_t)b * D * 3 * 4);
      copy(d_trans ? d_trans + (size_t)f0 * D * 3 : nullptr, det_trans_, (size_t)b * D * 3 * 4);
      copy(d_hand ? d_hand + (size_t)f0 * D * HMDPOSE_NUM_HAND : nullptr, det_hand_, (size_t)b * D * HMDPOSE_NUM_HAND * 4);
      copy(d_idx ? d_idx + (size_t)f0 * D : nullptr, det_idx_, (size_t)b * D * 4);
    }
    if (want_best && d_best && d_best != d_best_)
      copy(d_best + (size_t)HMDPOSE_BEST_LEN * f0, d_best_, (size_t)b * HMDPOSE_BEST_LEN * 4);
    last_b_ = b;
  }
  HP_CUDA(cudaEventRecord(ev1_, st));
  timing_valid_ = true;
  leave(st);
}

// Per-step device time: every launch bracketed by CUDA events on the handle's stream (un-graphed),
// inputs = whatever the last host-API call staged.  Used by bench.py for the roofline of the dominant kernel.
int Engine::profile_steps(int batch, int mode, int reps, char* names, char* kernels, float* ms, double* bytes,
                          double* flops, int capacity) {
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  const int S = cfg.image_size;
  const int b = std::min(mb_, std::max(batch, 1));
  // in-situ cost: T(steps[0..k]) - T(steps[0..k-1]), each prefix as a CUDA graph.  With 0x200 as well the prefix graph
  // runs on 8 streams at once (diagnostic: the marginal cost of a launch with 8 steps in flight; the streams share
  // this handle's buffers, so only the timing is meaningful)
  const bool prefix_mode = (mode & 0x300) != 0;
  const int nstreams = (mode & 0x200) ? 8 : 1;
  mode &= 0xff;
  Plan* plan = fast_ ? get_plan<__half>(b, mode) : get_plan<float>(b, mode);
  std::vector<Step> all;
  all.push_back(fast_ ? stem_step<__half>(d_in_stage_, 3LL * S * S, (long long)S * S, S, 1, b)
                      : stem_step<float>(d_in_stage_, 3LL * S * S, (long long)S * S, S, 1, b));
  for (const Step& s : plan->steps) all.push_back(s);
  const int n = (int)all.size();
  if (!ms) return n;
  if (capacity < n) throw Error(HMDPOSE_E_ARG, "profile_steps capacity too small");
  enter(stream);
  timing_valid_ = false;
  HP_CUDA(cudaMemcpyAsync(d_cam_local_, d_cam_stage_, (size_t)b * 24, cudaMemcpyDeviceToDevice, stream));
  std::vector<cudaEvent_t> ev((size_t)n + 1);
  for (auto& e : ev) HP_CUDA(cudaEventCreate(&e));
  std::vector<double> acc((size_t)n, 0.0);
  reps = std::max(reps, 1);
  if (prefix_mode) {
    wait_stream();
    double prev = 0.0;
    for (int k = 1; k <= n; ++k) {
      // concurrent replays share this handle's buffers: harmless for the network kernels (same values rewritten), but
      // the post-processing kernels index through data they are rewriting -- leave them out of the 8-stream mode
      if (nstreams > 1 && all[k - 1].name.rfind("post.", 0) == 0) { acc[k - 1] = 0.0; continue; }
      cudaStream_t cs;
      cudaGraph_t g = nullptr;
      cudaGraphExec_t ge = nullptr;
      HP_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      HP_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < k; ++i) all[i].launch(cs);
      HP_CUDA(cudaStreamEndCapture(cs, &g));
      cudaStreamDestroy(cs);
      HP_CUDA(cudaGraphInstantiate(&ge, g, 0));
      double cur = 0.0;
      if (nstreams == 1) {
        for (int r = 0; r < 3; ++r) HP_CUDA(cudaGraphLaunch(ge, stream));
        HP_CUDA(cudaEventRecord(ev[0], stream));
        for (int r = 0; r < reps; ++r) HP_CUDA(cudaGraphLaunch(ge, stream));
        HP_CUDA(cudaEventRecord(ev[1], stream));
        wait_stream();
        float t = 0.f;
        HP_CUDA(cudaEventElapsedTime(&t, ev[0], ev[1]));
        cur = (double)t / reps;
      } else {
        std::vector<cudaStream_t> ss((size_t)nstreams);
        std::vector<cudaGraphExec_t> ges((size_t)nstreams);
        std::vector<cudaEvent_t> done((size_t)nstreams);
        for (int q = 0; q < nstreams; ++q) {
          HP_CUDA(cudaStreamCreateWithFlags(&ss[q], cudaStreamNonBlocking));
          HP_CUDA(cudaGraphInstantiate(&ges[q], g, 0));
          HP_CUDA(cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming));
        }
        auto round = [&](int nr) {
          HP_CUDA(cudaEventRecord(ev[0], stream));
          for (int q = 0; q < nstreams; ++q) HP_CUDA(cudaStreamWaitEvent(ss[q], ev[0], 0));
          for (int r = 0; r < nr; ++r)
            for (int q = 0; q < nstreams; ++q) HP_CUDA(cudaGraphLaunch(ges[q], ss[q]));
          for (int q = 0; q < nstreams; ++q) {
            HP_CUDA(cudaEventRecord(done[q], ss[q]));
            HP_CUDA(cudaStreamWaitEvent(stream, done[q], 0));
          }
          HP_CUDA(cudaEventRecord(ev[1], stream));
          wait_stream();
        };
        round(2);
        round(reps);
        float t = 0.f;
        HP_CUDA(cudaEventElapsedTime(&t, ev[0], ev[1]));
        cur = (double)t / (reps * nstreams);   // time per step with `nstreams` steps in flight
        for (int q = 0; q < nstreams; ++q) { cudaGraphExecDestroy(ges[q]); cudaStreamDestroy(ss[q]); cudaEventDestroy(done[q]); }
      }
      acc[k - 1] = (cur - prev) * reps;
      prev = cur;
      cudaGraphExecDestroy(ge);
      cudaGraphDestroy(g);
    }
  }
  for (int r = 0; r < reps + 1 && !prefix_mode; ++r) {  // first repetition is a warm-up
    HP_CUDA(cudaEventRecord(ev[0], stream));
    for (int i = 0; i < n; ++i) {
      all[i].launch(stream);
      HP_CUDA(cudaEventRecord(ev[i + 1], stream));
    }
    wait_stream();
    HP_CUDA(cudaGetLastError());
    if (r == 0) continue;
    for (int i = 0; i < n; ++i) {
      float t = 0.f;
      HP_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
      acc[i] += t;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  leave(stream);
  for (int i = 0; i < n; ++i) {
    ms[i] = (float)(acc[i] / reps);
    if (bytes) bytes[i] = all[i].bytes;
    if (flops) flops[i] = all[i].flops;
    if (names) { std::strncpy(names + 64 * i, all[i].name.c_str(), 63); names[64 * i + 63] = 0; }
    if (kernels) { std::strncpy(kernels + 64 * i, all[i].kernel, 63); kernels[64 * i + 63] = 0; }
  }
  return n;
}

// ---- host-buffer API: pinned staging + H2D / D2H inside the call ---------------------------------
// pinned layout of the D0 results: per frame { rois[512][4], cls[512], scores[512], idx[512], count }
static constexpr size_t kD0FrameBytes = (size_t)D0_MAX_OUT * (16 + 4 + 4 + 4) + 16;

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

void Engine::ensure_host_staging(int batch, bool need_raw) {
  (void)batch;
  const int S = cfg.image_size, C = cfg.num_classes, D = cfg.max_detections, B = cfg.max_batch;
  if (!d_in_stage_) {
    d_in_stage_ = (float*)dalloc((size_t)B * 3 * S * S * 4);
    d_cam_stage_ = (float*)dalloc((size_t)B * 6 * 4);
    df_boxes_ = (float*)dalloc((size_t)B * D * 4 * 4);
    df_scores_ = (float*)dalloc((size_t)B * D * 4);
    df_labels_ = (int32_t*)dalloc((size_t)B * D * 4);
    df_rot_ = (float*)dalloc((size_t)B * D * 3 * 4);
    df_trans_ = (float*)dalloc((size_t)B * D * 3 * 4);
    df_hand_ = (float*)dalloc((size_t)B * D * HMDPOSE_NUM_HAND * 4);
    df_idx_ = (int32_t*)dalloc((size_t)B * D * 4);
  }
  const size_t raw_sizes[5] = {(size_t)N * 4, (size_t)N * C, (size_t)N * 3, (size_t)N * 3, (size_t)N * HMDPOSE_NUM_HAND};
  if (need_raw && !d_full_[0])  // full-batch copies of the five head tensors: only the Session.Run twin needs them
    for (int i = 0; i < 5; ++i) d_full_[i] = (float*)dalloc(raw_sizes[i] * B * 4);
  // pinned: input + cam + the largest result set this handle has been asked for
  const size_t in_b = (size_t)B * 3 * S * S * 4, cam_b = (size_t)B * 6 * 4;
  size_t out_b = std::max((size_t)B * D * (4 + 1 + 1 + 3 + 3 + HMDPOSE_NUM_HAND + 1) * 4, (size_t)B * kD0FrameBytes);
  if (need_raw) {
    size_t raw_b = 0;
    for (int i = 0; i < 5; ++i) raw_b += raw_sizes[i] * B * 4;
    out_b = std::max(out_b, raw_b);
  }
  const size_t need = in_b + cam_b + out_b + 4096;
  if (need > h_pinned_bytes_) {
    if (h_pinned_) { wait_stream(); cudaFreeHost(h_pinned_); h_pinned_ = nullptr; }
    HP_CUDA(cudaMallocHost((void**)&h_pinned_, need));
    h_pinned_bytes_ = need;
  }
}

void Engine::run_raw_host(const float* in, int batch, float* outs[5]) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!in) throw Error(HMDPOSE_E_ARG, "null input");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch, true);
  const int S = cfg.image_size, C = cfg.num_classes;
  const size_t in_b = (size_t)batch * 3 * S * S * 4;
  std::memcpy(h_pinned_, in, in_b);
  HP_CUDA(cudaMemcpyAsync(d_in_stage_, h_pinned_, in_b, cudaMemcpyHostToDevice, stream));
  run_device(d_in_stage_, 3LL * S * S, (long long)S * S, S, 1, nullptr, batch, true, d_full_, false, nullptr, nullptr,
             nullptr, nullptr, nullptr, nullptr, nullptr, false, nullptr, stream);
  const size_t sizes[5] = {(size_t)N * 4, (size_t)N * C, (size_t)N * 3, (size_t)N * 3, (size_t)N * HMDPOSE_NUM_HAND};
  uint8_t* hp = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4 + (size_t)cfg.max_batch * 24;
  uint8_t* cur = hp;
  for (int i = 0; i < 5; ++i) {
    if (outs[i]) HP_CUDA(cudaMemcpyAsync(cur, d_full_[i], sizes[i] * batch * 4, cudaMemcpyDeviceToHost, stream));
    cur += sizes[i] * batch * 4;
  }
  wait_stream();
  cur = hp;
  for (int i = 0; i < 5; ++i) {
    if (outs[i]) std::memcpy(outs[i], cur, sizes[i] * batch * 4);
    cur += sizes[i] * batch * 4;
  }
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::run_detect_host(const float* in, const float* cam, int batch, float* boxes, float* scores,
                             int32_t* labels, float* rot, float* trans, float* hand, int32_t* idx) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!in || !cam) throw Error(HMDPOSE_E_ARG, "null input / camera parameters");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  const int S = cfg.image_size, D = cfg.max_detections;
  const size_t in_b = (size_t)batch * 3 * S * S * 4;
  uint8_t* h_cam = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4;
  uint8_t* h_out = h_cam + (size_t)cfg.max_batch * 24;
  const void* src = in;
  if (!is_pinned(in)) { std::memcpy(h_pinned_, in, in_b); src = h_pinned_; }  // page-locked callers skip the staging copy
  std::memcpy(h_cam, cam, (size_t)batch * 24);
  HP_CUDA(cudaMemcpyAsync(d_in_stage_, src, in_b, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(d_cam_stage_, h_cam, (size_t)batch * 24, cudaMemcpyHostToDevice, stream));
  run_device(d_in_stage_, 3LL * S * S, (long long)S * S, S, 1, d_cam_stage_, batch, false, nullptr, true, df_boxes_,
             df_scores_, df_labels_, df_rot_, df_trans_, df_hand_, df_idx_, false, nullptr, stream);
  struct Out { void* user; const void* dev; size_t bytes; };
  const Out outs[7] = {{boxes, df_boxes_, (size_t)batch * D * 16}, {scores, df_scores_, (size_t)batch * D * 4},
                       {labels, df_labels_, (size_t)batch * D * 4}, {rot, df_rot_, (size_t)batch * D * 12},
                       {trans, df_trans_, (size_t)batch * D * 12},
                       {hand, df_hand_, (size_t)batch * D * HMDPOSE_NUM_HAND * 4}, {idx, df_idx_, (size_t)batch * D * 4}};
  uint8_t* cur = h_out;
  for (const Out& o : outs) {
    if (o.user) HP_CUDA(cudaMemcpyAsync(cur, o.dev, o.bytes, cudaMemcpyDeviceToHost, stream));
    cur += o.bytes;
  }
  wait_stream();
  cur = h_out;
  for (const Out& o : outs) {
    if (o.user) std::memcpy(o.user, cur, o.bytes);
    cur += o.bytes;
  }
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::run_best_host(const float* in, const float* cam, float* out11) {
  if (!in || !cam || !out11) throw Error(HMDPOSE_E_ARG, "null argument");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(1);
  const int S = cfg.image_size;
  const size_t in_b = (size_t)3 * S * S * 4;
  uint8_t* h_cam = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4;
  uint8_t* h_out = h_cam + (size_t)cfg.max_batch * 24;
  std::memcpy(h_pinned_, in, in_b);
  std::memcpy(h_cam, cam, 24);
  HP_CUDA(cudaMemcpyAsync(d_in_stage_, h_pinned_, in_b, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(d_cam_stage_, h_cam, 24, cudaMemcpyHostToDevice, stream));
  run_device(d_in_stage_, 3LL * S * S, (long long)S * S, S, 1, d_cam_stage_, 1, false, nullptr, false, nullptr, nullptr,
             nullptr, nullptr, nullptr, nullptr, nullptr, true, d_best_, stream);
  HP_CUDA(cudaMemcpyAsync(h_out, d_best_, HMDPOSE_BEST_LEN * 4, cudaMemcpyDeviceToHost, stream));
  wait_stream();
  std::memcpy(out11, h_out, HMDPOSE_BEST_LEN * 4);
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::postprocess_host(const float* reg, const float* cls, const float* rot, const float* traw,
                              const float* hand, const float* cam, const float* boxes_in, const float* trans_in,
                              int batch, float* boxes, float* scores, int32_t* labels, float* rot_o, float* trans_o,
                              float* hand_o, int32_t* idx) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!cls || !rot || !hand) throw Error(HMDPOSE_E_ARG, "null head tensor");
  if (!boxes_in && (!reg || !traw || !cam)) throw Error(HMDPOSE_E_ARG, "null head tensor");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  const int C = cfg.num_classes, D = cfg.max_detections;
  auto up = [&](float* dst, const float* src, size_t n) {
    if (src) HP_CUDA(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyHostToDevice, stream));
  };
  last_launches = 0;
  enter(stream);
  HP_CUDA(cudaEventRecord(ev0_, stream));
  for (int f0 = 0; f0 < batch; f0 += mb_) {
    const int b = std::min(mb_, batch - f0);
    up(o_reg_, reg ? reg + (size_t)f0 * N * 4 : nullptr, (size_t)b * N * 4);
    up(o_cls_, cls + (size_t)f0 * N * C, (size_t)b * N * C);
    up(o_rot_, rot + (size_t)f0 * N * 3, (size_t)b * N * 3);
    up(o_traw_, traw ? traw + (size_t)f0 * N * 3 : nullptr, (size_t)b * N * 3);
    up(o_hand_, hand + (size_t)f0 * N * HMDPOSE_NUM_HAND, (size_t)b * N * HMDPOSE_NUM_HAND);
    up(d_cam_local_, cam ? cam + (size_t)f0 * 6 : nullptr, (size_t)b * 6);
    up(p_boxes_, boxes_in ? boxes_in + (size_t)f0 * N * 4 : nullptr, (size_t)b * N * 4);
    up(p_trans_, trans_in ? trans_in + (size_t)f0 * N * 3 : nullptr, (size_t)b * N * 3);
    {
      std::vector<Step> ps;
      add_post_steps(ps, b, PLAN_DET, boxes_in == nullptr, trans_in == nullptr, true);
      for (Step& q : ps) q.launch(stream);
      last_launches += (int)ps.size();
      HP_CUDA(cudaGetLastError());
    }
    auto down = [&](void* dst, const void* src, size_t bytes) {
      if (dst) HP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    };
    down(boxes ? boxes + (size_t)f0 * D * 4 : nullptr, det_boxes_, (size_t)b * D * 16);
    down(scores ? scores + (size_t)f0 * D : nullptr, det_scores_, (size_t)b * D * 4);
    down(labels ? labels + (size_t)f0 * D : nullptr, det_labels_, (size_t)b * D * 4);
    down(rot_o ? rot_o + (size_t)f0 * D * 3 : nullptr, det_rot_, (size_t)b * D * 12);
    down(trans_o ? trans_o + (size_t)f0 * D * 3 : nullptr, det_trans_, (size_t)b * D * 12);
    down(hand_o ? hand_o + (size_t)f0 * D * HMDPOSE_NUM_HAND : nullptr, det_hand_, (size_t)b * D * HMDPOSE_NUM_HAND * 4);
    down(idx ? idx + (size_t)f0 * D : nullptr, det_idx_, (size_t)b * D * 4);
    wait_stream();  // pageable host buffers: finish before the next chunk reuses staging
  }
  HP_CUDA(cudaEventRecord(ev1_, stream));
  timing_valid_ = true;
  leave(stream);
  wait_stream();
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::best_from_raw_host(const float* reg, const float* cls, const float* rot, const float* traw,
                                const float* cam, float* out11) {
  if (!reg || !cls || !rot || !traw || !cam || !out11) throw Error(HMDPOSE_E_ARG, "null argument");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(1);
  const int C = cfg.num_classes;
  enter(stream);
  HP_CUDA(cudaMemcpyAsync(o_reg_, reg, (size_t)N * 16, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(o_cls_, cls, (size_t)N * C * 4, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(o_rot_, rot, (size_t)N * 12, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(o_traw_, traw, (size_t)N * 12, cudaMemcpyHostToDevice, stream));
  HP_CUDA(cudaMemcpyAsync(d_cam_local_, cam, 24, cudaMemcpyHostToDevice, stream));
  last_launches = 0;
  {
    std::vector<Step> ps;
    add_post_steps(ps, 1, PLAN_BEST, false, false, true);
    for (Step& q : ps) q.launch(stream);
    last_launches += (int)ps.size();
    HP_CUDA(cudaGetLastError());
  }
  HP_CUDA(cudaMemcpyAsync(out11, d_best_, HMDPOSE_BEST_LEN * 4, cudaMemcpyDeviceToHost, stream));
  timing_valid_ = false;
  leave(stream);
  wait_stream();
}

// ---- EfficientDet-d0 detection variant ------------------------------------------------------------
void Engine::ensure_d0(float thr, float iou) {
  if (!d_anchors_d0_) {
    std::vector<float> a;
    compute_anchors_d0(cfg.image_size, a);
    if ((int)(a.size() / 4) != N) throw Error(HMDPOSE_E_STATE, "D0 anchor count mismatch");
    d_anchors_d0_ = upload_f32(a.data(), a.size());
    const int b = mb_;
    d0_cand_cls_ = (int*)dalloc((size_t)b * N * 4);
    d0_count_ = (int*)dalloc((size_t)b * 4);
    d0_orois_ = (float*)dalloc((size_t)b * D0_MAX_OUT * 16);
    d0_ocls_ = (int*)dalloc((size_t)b * D0_MAX_OUT * 4);
    d0_oscores_ = (float*)dalloc((size_t)b * D0_MAX_OUT * 4);
    d0_oidx_ = (int*)dalloc((size_t)b * D0_MAX_OUT * 4);
    d0_ocount_ = (int*)dalloc((size_t)b * 4);
    d0_sel_ = (float*)dalloc((size_t)b * D0_MAX_OUT * 16);
  }
  if (thr != d0_thr_ || iou != d0_iou_) {  // thresholds are baked into the captured launch plans
    wait_stream();
    for (auto it = plans_.begin(); it != plans_.end();)
      it = (it->first % 8 == PLAN_D0) ? plans_.erase(it) : std::next(it);
    d0_thr_ = thr; d0_iou_ = iou;
  }
}

D0Args Engine::d0_args() const {
  D0Args a;
  std::memset(&a, 0, sizeof(a));
  a.anchors_yxyx = d_anchors_d0_; a.reg = o_reg_; a.cls = o_cls_;
  a.N = N; a.C = cfg.num_classes; a.cap = pb_.cap; a.max_out = D0_MAX_OUT;
  a.wmax = (float)(cfg.image_size - 1); a.hmax = (float)(cfg.image_size - 1);
  a.threshold = d0_thr_; a.iou_thr = d0_iou_;
  a.keys = pb_.keys; a.cand_cls = d0_cand_cls_; a.cand_count = d0_count_; a.box_scratch = p_boxes_;
  a.o_rois = d0_orois_; a.o_cls = d0_ocls_; a.o_scores = d0_oscores_; a.o_idx = d0_oidx_; a.o_count = d0_ocount_;
  a.sel_scratch = d0_sel_;
  return a;
}

void Engine::d0_download(int f0, int b, uint8_t* h_out) {
  for (int i = 0; i < b; ++i) {
    uint8_t* dst = h_out + (size_t)(f0 + i) * kD0FrameBytes;
    HP_CUDA(cudaMemcpyAsync(dst, d0_orois_ + (size_t)i * D0_MAX_OUT * 4, D0_MAX_OUT * 16, cudaMemcpyDeviceToHost, stream));
    HP_CUDA(cudaMemcpyAsync(dst + D0_MAX_OUT * 16, d0_ocls_ + (size_t)i * D0_MAX_OUT, D0_MAX_OUT * 4, cudaMemcpyDeviceToHost, stream));
    HP_CUDA(cudaMemcpyAsync(dst + D0_MAX_OUT * 20, d0_oscores_ + (size_t)i * D0_MAX_OUT, D0_MAX_OUT * 4, cudaMemcpyDeviceToHost, stream));
    HP_CUDA(cudaMemcpyAsync(dst + D0_MAX_OUT * 24, d0_oidx_ + (size_t)i * D0_MAX_OUT, D0_MAX_OUT * 4, cudaMemcpyDeviceToHost, stream));
    HP_CUDA(cudaMemcpyAsync(dst + D0_MAX_OUT * 28, d0_ocount_ + i, 4, cudaMemcpyDeviceToHost, stream));
  }
}

void Engine::d0_scatter(const uint8_t* h_out, int batch, int max_out, float* rois, int32_t* class_ids, float* scores,
                        int32_t* idx, int32_t* counts) {
  for (int f = 0; f < batch; ++f) {
    const uint8_t* src = h_out + (size_t)f * kD0FrameBytes;
    if (rois) std::memcpy(rois + (size_t)f * max_out * 4, src, (size_t)max_out * 16);
    if (class_ids) std::memcpy(class_ids + (size_t)f * max_out, src + D0_MAX_OUT * 16, (size_t)max_out * 4);
    if (scores) std::memcpy(scores + (size_t)f * max_out, src + D0_MAX_OUT * 20, (size_t)max_out * 4);
    if (idx) std::memcpy(idx + (size_t)f * max_out, src + D0_MAX_OUT * 24, (size_t)max_out * 4);
    if (counts) {
      int32_t c;
      std::memcpy(&c, src + D0_MAX_OUT * 28, 4);
      // c < 0: the device capacity was reached with candidates left; more rows than the caller's max_out: same flag
      const int kept = c < 0 ? -c : c;
      counts[f] = (c < 0 || kept > max_out) ? -std::min(kept, max_out) : kept;
    }
  }
}

void Engine::run_d0_host(const float* in, int batch, float thr, float iou, int max_out, float* rois, int32_t* class_ids,
                         float* scores, int32_t* idx, int32_t* counts) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!in) throw Error(HMDPOSE_E_ARG, "null input");
  if (max_out < 1 || max_out > D0_MAX_OUT) throw Error(HMDPOSE_E_ARG, "max_out must be in [1, 4096]");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  ensure_d0(thr, iou);
  const int S = cfg.image_size;
  const size_t in_b = (size_t)batch * 3 * S * S * 4;
  uint8_t* h_out = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4 + (size_t)cfg.max_batch * 24;
  const void* src = in;
  if (!is_pinned(in)) { std::memcpy(h_pinned_, in, in_b); src = h_pinned_; }
  HP_CUDA(cudaMemcpyAsync(d_in_stage_, src, in_b, cudaMemcpyHostToDevice, stream));
  last_launches = 0;
  enter(stream);
  HP_CUDA(cudaEventRecord(ev0_, stream));
  const long long sb = 3LL * S * S;
  for (int f0 = 0; f0 < batch; f0 += mb_) {
    const int b = std::min(mb_, batch - f0);
    Plan* plan = fast_ ? get_plan<__half>(b, PLAN_D0) : get_plan<float>(b, PLAN_D0);
    Step stem = fast_ ? stem_step<__half>(d_in_stage_ + f0 * sb, sb, (long long)S * S, S, 1, b)
                      : stem_step<float>(d_in_stage_ + f0 * sb, sb, (long long)S * S, S, 1, b);
    stem.launch(stream);
    run_plan(plan, stream);
    last_launches += 2 + plan->launches;  // the D0 post step is two kernels
    d0_download(f0, b, h_out);
    last_b_ = b;
  }
  HP_CUDA(cudaEventRecord(ev1_, stream));
  timing_valid_ = true;
  leave(stream);
  wait_stream();
  d0_scatter(h_out, batch, max_out, rois, class_ids, scores, idx, counts);
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::d0_postprocess_host(const float* reg, const float* cls, int batch, float thr, float iou, int max_out,
                                 float* rois, int32_t* class_ids, float* scores, int32_t* idx, int32_t* counts) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!reg || !cls) throw Error(HMDPOSE_E_ARG, "null head tensor");
  if (max_out < 1 || max_out > D0_MAX_OUT) throw Error(HMDPOSE_E_ARG, "max_out must be in [1, 4096]");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  ensure_d0(thr, iou);
  const int S = cfg.image_size, C = cfg.num_classes;
  uint8_t* h_out = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4 + (size_t)cfg.max_batch * 24;
  last_launches = 0;
  enter(stream);
  HP_CUDA(cudaEventRecord(ev0_, stream));
  for (int f0 = 0; f0 < batch; f0 += mb_) {
    const int b = std::min(mb_, batch - f0);
    HP_CUDA(cudaMemcpyAsync(o_reg_, reg + (size_t)f0 * N * 4, (size_t)b * N * 16, cudaMemcpyHostToDevice, stream));
    HP_CUDA(cudaMemcpyAsync(o_cls_, cls + (size_t)f0 * N * C, (size_t)b * N * C * 4, cudaMemcpyHostToDevice, stream));
    launch_d0(d0_args(), b, stream);
    HP_CUDA(cudaGetLastError());
    last_launches += 2;
    d0_download(f0, b, h_out);
    wait_stream();  // pageable host inputs: finish before the next chunk reuses o_*
  }
  HP_CUDA(cudaEventRecord(ev1_, stream));
  timing_valid_ = true;
  leave(stream);
  wait_stream();
  d0_scatter(h_out, batch, max_out, rois, class_ids, scores, idx, counts);
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

// ---- uint8 frames: pre-processing on the device --------------------------------------------------
float Engine::stage_u8(const uint8_t* imgs, int batch, int h, int w) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!imgs || h < 1 || w < 1) throw Error(HMDPOSE_E_ARG, "null / empty frame");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  const int S = cfg.image_size;
  const size_t bytes = (size_t)batch * h * w * 3;
  if (bytes > d_u8_bytes_) {   // grows with the largest frame seen (not captured in any graph)
    wait_stream();
    if (d_u8_) cudaFree(d_u8_);
    HP_CUDA(cudaMalloc((void**)&d_u8_, bytes));
    d_u8_bytes_ = bytes;
  }
  HP_CUDA(cudaMemcpyAsync(d_u8_, imgs, bytes, cudaMemcpyHostToDevice, stream));
  PreArgs a;
  a.img = d_u8_; a.out = d_in_stage_; a.B = batch; a.h = h; a.w = w; a.S = S;
  double scale;
  if (h > w) { scale = (double)S / h; a.rh = S; a.rw = (int)(w * scale); }        // colibri_common.py:633-640
  else { scale = (double)S / w; a.rh = (int)(h * scale); a.rw = S; }
  launch_preprocess(a, stream);
  HP_CUDA(cudaGetLastError());
  return (float)scale;
}

// I420 frames through the C# receiver's path (Program.cs:137-200): H2D of the 1.5-byte-per-pixel frame + one kernel
float Engine::stage_i420(const uint8_t* frames, int batch, int h, int w, int crop, int mid) {
  if (batch < 1 || batch > cfg.max_batch) throw Error(HMDPOSE_E_ARG, "batch out of range [1, max_batch]");
  if (!frames || h < 2 || w < 2 || (h & 1) || (w & 1)) throw Error(HMDPOSE_E_ARG, "I420 frames need even, positive height / width");
  if (crop < 1 || crop > h || crop > w || mid < 1) throw Error(HMDPOSE_E_ARG, "bad crop / rescale size");
  HP_CUDA(cudaSetDevice(cfg.device));
  ensure_host_staging(batch);
  const int S = cfg.image_size;
  const size_t bytes = (size_t)batch * h * w * 3 / 2;
  if (bytes > d_u8_bytes_) {
    wait_stream();
    if (d_u8_) cudaFree(d_u8_);
    d_u8_ = nullptr; d_u8_bytes_ = 0;
    HP_CUDA(cudaMalloc((void**)&d_u8_, bytes));
    d_u8_bytes_ = bytes;
  }
  HP_CUDA(cudaMemcpyAsync(d_u8_, frames, bytes, cudaMemcpyHostToDevice, stream));
  I420Args a;
  a.img = d_u8_; a.out = d_in_stage_; a.B = batch; a.h = h; a.w = w; a.crop = crop; a.mid = mid; a.S = S;
  // ResizeAndNormalizeMat (Program.cs:397-418) on the mid x mid image: float32 scale, truncating int cast
  const float scale = (float)S / (float)mid;
  a.rh = S; a.rw = (int)((float)mid * scale);
  a.rh = a.rw;   // square input: both sides scale alike
  launch_preprocess_i420(a, stream);
  HP_CUDA(cudaGetLastError());
  return scale;
}

void Engine::preprocess_i420_host(const uint8_t* frames, int batch, int h, int w, int crop, int mid, float* out_nhwc,
                                  float* scale) {
  if (!out_nhwc) throw Error(HMDPOSE_E_ARG, "null output");
  const float sc = stage_i420(frames, batch, h, w, crop, mid);
  const int S = cfg.image_size;
  HP_CUDA(cudaMemcpyAsync(out_nhwc, d_in_stage_, (size_t)batch * S * S * 3 * 4, cudaMemcpyDeviceToHost, stream));
  wait_stream();
  if (scale) *scale = sc;
}

void Engine::run_best_i420_host(const uint8_t* frame, int h, int w, int crop, int mid, const float* cam, float* out11,
                                float* scale) {
  if (!cam || !out11) throw Error(HMDPOSE_E_ARG, "null argument");
  const float sc = stage_i420(frame, 1, h, w, crop, mid);
  if (scale) *scale = sc;
  const int S = cfg.image_size;
  uint8_t* h_cam = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4;
  uint8_t* h_out = h_cam + (size_t)cfg.max_batch * 24;
  std::memcpy(h_cam, cam, 24);
  HP_CUDA(cudaMemcpyAsync(d_cam_stage_, h_cam, 24, cudaMemcpyHostToDevice, stream));
  run_device(d_in_stage_, 3LL * S * S, 1, 3LL * S, 3, d_cam_stage_, 1, false, nullptr, false, nullptr, nullptr, nullptr,
             nullptr, nullptr, nullptr, nullptr, true, d_best_, stream);
  HP_CUDA(cudaMemcpyAsync(h_out, d_best_, HMDPOSE_BEST_LEN * 4, cudaMemcpyDeviceToHost, stream));
  wait_stream();
  std::memcpy(out11, h_out, HMDPOSE_BEST_LEN * 4);
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::preprocess_host(const uint8_t* imgs, int batch, int h, int w, float* out_nhwc, float* scale) {
  if (!out_nhwc) throw Error(HMDPOSE_E_ARG, "null output");
  const float sc = stage_u8(imgs, batch, h, w);
  const int S = cfg.image_size;
  HP_CUDA(cudaMemcpyAsync(out_nhwc, d_in_stage_, (size_t)batch * S * S * 3 * 4, cudaMemcpyDeviceToHost, stream));
  wait_stream();
  if (scale) *scale = sc;
}

void Engine::run_detect_u8_host(const uint8_t* imgs, int batch, int h, int w, const float* cam, float* boxes,
                                float* scores, int32_t* labels, float* rot, float* trans, float* hand, int32_t* idx,
                                float* scale) {
  if (!cam) throw Error(HMDPOSE_E_ARG, "null camera parameters");
  const float sc = stage_u8(imgs, batch, h, w);
  if (scale) *scale = sc;
  const int S = cfg.image_size, D = cfg.max_detections;
  uint8_t* h_cam = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4;
  uint8_t* h_out = h_cam + (size_t)cfg.max_batch * 24;
  std::memcpy(h_cam, cam, (size_t)batch * 24);
  HP_CUDA(cudaMemcpyAsync(d_cam_stage_, h_cam, (size_t)batch * 24, cudaMemcpyHostToDevice, stream));
  // the staged tensor is NHWC: element strides of the NCHW view the network consumes (eval/common.py:397)
  run_device(d_in_stage_, 3LL * S * S, 1, 3LL * S, 3, d_cam_stage_, batch, false, nullptr, true, df_boxes_, df_scores_,
             df_labels_, df_rot_, df_trans_, df_hand_, df_idx_, false, nullptr, stream);
  struct Out { void* user; const void* dev; size_t bytes; };
  const Out outs[7] = {{boxes, df_boxes_, (size_t)batch * D * 16}, {scores, df_scores_, (size_t)batch * D * 4},
                       {labels, df_labels_, (size_t)batch * D * 4}, {rot, df_rot_, (size_t)batch * D * 12},
                       {trans, df_trans_, (size_t)batch * D * 12},
                       {hand, df_hand_, (size_t)batch * D * HMDPOSE_NUM_HAND * 4}, {idx, df_idx_, (size_t)batch * D * 4}};
  uint8_t* cur = h_out;
  for (const Out& o : outs) {
    if (o.user) HP_CUDA(cudaMemcpyAsync(cur, o.dev, o.bytes, cudaMemcpyDeviceToHost, stream));
    cur += o.bytes;
  }
  wait_stream();
  cur = h_out;
  for (const Out& o : outs) {
    if (o.user) std::memcpy(o.user, cur, o.bytes);
    cur += o.bytes;
  }
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

void Engine::run_best_u8_host(const uint8_t* img, int h, int w, const float* cam, float* out11, float* scale) {
  if (!cam || !out11) throw Error(HMDPOSE_E_ARG, "null argument");
  const float sc = stage_u8(img, 1, h, w);
  if (scale) *scale = sc;
  const int S = cfg.image_size;
  uint8_t* h_cam = h_pinned_ + (size_t)cfg.max_batch * 3 * S * S * 4;
  uint8_t* h_out = h_cam + (size_t)cfg.max_batch * 24;
  std::memcpy(h_cam, cam, 24);
  HP_CUDA(cudaMemcpyAsync(d_cam_stage_, h_cam, 24, cudaMemcpyHostToDevice, stream));
  run_device(d_in_stage_, 3LL * S * S, 1, 3LL * S, 3, d_cam_stage_, 1, false, nullptr, false, nullptr, nullptr, nullptr,
             nullptr, nullptr, nullptr, nullptr, true, d_best_, stream);
  HP_CUDA(cudaMemcpyAsync(h_out, d_best_, HMDPOSE_BEST_LEN * 4, cudaMemcpyDeviceToHost, stream));
  wait_stream();
  std::memcpy(out11, h_out, HMDPOSE_BEST_LEN * 4);
  HP_CUDA(cudaEventElapsedTime(&last_ms, ev0_, ev1_));
}

long long Engine::debug_read(const std::string& name, float* out, long long cap) {
  if (name == "__s3_timeline") {
    HP_CUDA(cudaSetDevice(cfg.device));
    wait_stream();
    if (!out) return 32;
    return sep3_debug_timeline(out, (int)cap);
  }
  if (name == "__mb_timeline") {
    HP_CUDA(cudaSetDevice(cfg.device));
    wait_stream();
    if (!out) return 32;
    return mb_debug_timeline(out, (int)cap);
  }
  auto it = debug_.find(name);
  if (it == debug_.end()) throw Error(HMDPOSE_E_ARG, "unknown debug tensor " + name);
  const Tens& t = it->second.first;
  const int b = std::max(last_b_, 1);
  const long long n = (long long)t.elems(b);
  if (!out) return n;
  if (cap < n) throw Error(HMDPOSE_E_ARG, "debug_read capacity too small");
  HP_CUDA(cudaSetDevice(cfg.device));
  enter(stream);   // the last run may have used a caller stream
  wait_stream();
  if (fast_ && it->second.second) {
    std::vector<__half> h((size_t)n);
    HP_CUDA(cudaMemcpy(h.data(), t.p, (size_t)n * 2, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n; ++i) out[i] = __half2float(h[(size_t)i]);
  } else {
    HP_CUDA(cudaMemcpy(out, t.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  return n;
}

}  // namespace hp
