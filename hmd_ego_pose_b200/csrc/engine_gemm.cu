// Pointwise-conv GEMM dispatch: FFMA kernel (parity mode / cross-check) or tcgen05+TMA kernel (fast mode).
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstring>
#include <set>

#include "engine.h"
#include "gemm_tc.cuh"
#include "gemm_tf32.cuh"
#include "sepconv_tc.cuh"
#include "sepconv3_tc.cuh"
#include "mbconv_tc.cuh"
#include "expdw_tc.cuh"
#include "projk_tc.cuh"

namespace hp {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_num_sms = 148;

void trap_info_install_gemm_tu(void* mapped) { HP_CUDA(trap_info_install_tu((TrapInfo*)mapped)); }

void init_gemm_kernels() {
  // function attributes are per device: run once for every device this process uses
  static std::mutex mu;
  static std::set<int> done;
  int cur = 0;
  HP_CUDA(cudaGetDevice(&cur));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(cur)) return;
  [] {
    // libcuda is resolved at run time through the runtime API so that the library still loads
    // (and exports its symbols) on a machine without a driver.
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      throw Error(HMDPOSE_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    HP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc_smem_bytes(TC_MAX_STAGES, 128)));
    HP_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    HP_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    HP_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    HP_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    HP_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    HP_CUDA(cudaFuncSetAttribute(sepconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sep_smem_bytes(128, true)));
    HP_CUDA(cudaFuncSetAttribute(sepconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sep_smem_bytes(128)));
    HP_CUDA(cudaFuncSetAttribute(sepconv3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    HP_CUDA(cudaFuncSetAttribute(sepconv3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    int dev = 0;
    HP_CUDA(cudaGetDevice(&dev));
    HP_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }();
  done.insert(cur);
}

// Split N into tiles of width bn (multiple of 16, <= cap <= 128) wasting as few padded columns as possible.
int gemm_choose_bn(int N, int* n_tiles, int cap) {
  int best_bn = cap, best_t = cdiv(N, cap), best_waste = best_t * cap - N;
  for (int t = cdiv(N, cap); t <= cdiv(N, cap) + 3; ++t) {
    int bn = ((cdiv(N, t) + 15) / 16) * 16;
    if (bn > cap) continue;
    int waste = t * bn - N;
    if (waste < best_waste) { best_waste = waste; best_bn = bn; best_t = t; }
  }
  *n_tiles = best_t;
  return best_bn;
}

static void encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes,
                      uint32_t box_inner, uint32_t box_outer, bool f32 = false) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(HMDPOSE_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") inner=" +
                                    std::to_string(inner) + " outer=" + std::to_string(outer) + " pitch=" +
                                    std::to_string(row_bytes));
}

// 4-D tensor map of an NHWC activation tensor (C, W, H, B innermost first), un-swizzled box (box_c, box_w, box_h, 1).
void encode_act_4d(CUtensorMap* tm, const void* base, bool is_half, int C, int W, int H, int B, int box_c, int box_w,
                   int box_h, bool swizzle128) {
  init_gemm_kernels();
  const uint64_t es = is_half ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(HMDPOSE_E_CUDA, "cuTensorMapEncodeTiled(4d) failed (" + std::to_string((int)r) + ") C=" + std::to_string(C) +
                                    " W=" + std::to_string(W) + " box=" + std::to_string(box_c) + "x" + std::to_string(box_w) +
                                    "x" + std::to_string(box_h));
}

std::function<void(cudaStream_t)> make_gemm_launcher(std::vector<GemmProb> probs, bool fast, bool force_simt,
                                                     std::vector<void*>& owned, const char** kernel_name, bool v1) {
  const int n = (int)probs.size();
  if (kernel_name) *kernel_name = (fast && !force_simt) ? (v1 ? "gemm_tc_kernel" : "gemm_tc2_kernel") : "gemm_simt_kernel";
  if (fast && !force_simt) {
    init_gemm_kernels();
    std::vector<TcProb> tp(n);
    int tiles = 0, bn_max = 16, kb_max = 1;
    for (int i = 0; i < n; ++i) {
      GemmProb& p = probs[i];
      if (p.K % 8 != 0 || p.lda % 8 != 0)
        throw Error(HMDPOSE_E_STATE, "tcgen05 GEMM needs K and lda multiples of 8 (16-byte TMA pitch)");
      p.bn = gemm_choose_bn(p.N, &p.n_tiles);
      p.m_tiles = cdiv(p.M, TC_BM);
      p.tile_start = tiles;
      tiles += p.m_tiles * p.n_tiles;
      bn_max = std::max(bn_max, p.bn);
      kb_max = std::max(kb_max, cdiv(p.K, TC_BK));
      std::memset(&tp[i], 0, sizeof(TcProb));
      encode_2d(&tp[i].tmA, p.A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.lda * 2, TC_BK, TC_BM);
      uint64_t w_rows = (uint64_t)p.N;
      if (p.w_img_rows > 0) {   // one weight panel per image (gate folded into the weights)
        if (v1 || n != 1 || p.rows_per_img % TC_BM != 0 || p.w_img_rows != p.N)
          throw Error(HMDPOSE_E_STATE, "per-image weight panels need a single problem whose images are whole m tiles");
        w_rows = (uint64_t)p.N * (uint64_t)(p.M / p.rows_per_img);
      }
      encode_2d(&tp[i].tmB, p.W, (uint64_t)p.K, w_rows, (uint64_t)p.K * 2, TC_BK, (uint32_t)p.bn);
      tp[i].p = p;
    }
    TcProb* d = nullptr;
    HP_CUDA(cudaMalloc(&d, sizeof(TcProb) * n));
    owned.push_back(d);
    if (!v1) {
      bool gated = false;
      for (const GemmProb& p : probs) gated = gated || p.a_scale != nullptr;
      bool headout = false, plain = false;
      for (const GemmProb& p : probs) { headout = headout || p.out_mode != 0; plain = plain || p.out_mode == 0; }
      if (headout && (plain || gated)) throw Error(HMDPOSE_E_STATE, "head-tensor and NHWC outputs cannot share a GEMM launch");
      // weight panels that fit stay resident in shared memory (one TMA per run of tiles instead of one per tile);
      // the others stream through the ring next to their A tiles
      int ring_bytes = 0, res_bytes = 0;
      const bool reuse = tiles > 2 * g_num_sms && std::getenv("HMDPOSE_NO_BRES") == nullptr;   // some CTA walks more than one tile
      for (int i = 0; i < n; ++i) {
        const GemmProb& p = tp[i].p;
        const int panel = cdiv(p.K, TC_BK) * p.bn * TC_BK * 2;
        tp[i].p.b_res = (reuse && panel <= TC2_RES_MAX) ? 1 : 0;
        if (tp[i].p.b_res) res_bytes = std::max(res_bytes, panel);
        else ring_bytes = std::max(ring_bytes, p.bn * TC_BK * 2);
      }
      HP_CUDA(cudaMemcpy(d, tp.data(), sizeof(TcProb) * n, cudaMemcpyHostToDevice));
      // ring depth: enough k-blocks in flight to hide the TMA latency (K = 1152 -> 18 k-blocks; single-k-block
      // problems keep four tiles in flight), capped so that two CTAs stay resident per SM when possible
      int stages = std::max(reuse ? 4 : 2, std::min(kb_max, TC2_MAX_STAGES));
      const int smem_cap = (tiles <= g_num_sms || ring_bytes > 0) ? 200 * 1024 : 112 * 1024;
      while (stages > 2 && tc2_smem_bytes(stages, ring_bytes, res_bytes, gated, headout) > smem_cap) --stages;
      const int smem2 = tc2_smem_bytes(stages, ring_bytes, res_bytes, gated, headout);
      int per_sm = std::max(1, std::min(2, (227 * 1024) / (smem2 + 1024)));
      if (const char* e = std::getenv("HMDPOSE_GEMM_PER_SM")) per_sm = std::max(1, std::min(per_sm, std::atoi(e)));
      int grid = std::min(tiles, per_sm * g_num_sms);
      if (const char* e = std::getenv("HMDPOSE_MAX_TILES_PER_CTA")) grid = std::min(tiles, std::max(grid, cdiv(tiles, std::max(1, std::atoi(e)))));
      const int threads = gated ? TC2_THREADS_GATED : TC2_THREADS;
      if (headout)
        return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<false, true>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages, ring_bytes, res_bytes)); };
      if (gated)
        return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<true, false>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages, ring_bytes, res_bytes)); };
      return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<false, false>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages, ring_bytes, res_bytes)); };
    }
    HP_CUDA(cudaMemcpy(d, tp.data(), sizeof(TcProb) * n, cudaMemcpyHostToDevice));
    // stages: enough to cover K, capped so that >= 2 CTAs fit per SM
    int stages = std::min(kb_max, TC_MAX_STAGES);
    while (stages > 2 && tc_smem_bytes(stages, bn_max) > 100 * 1024) --stages;
    const int smem = tc_smem_bytes(stages, bn_max);
    return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc_kernel, dim3(tiles), dim3(TC_THREADS), smem, st, d, n, stages, bn_max)); };
  }
  if (!fast && !force_simt) {
    // parity mode: split-precision (3xTF32) tcgen05 GEMM on fp32 activations (gemm_tf32.cuh)
    init_gemm_kernels();
    if (kernel_name) *kernel_name = "gemm_tf32_kernel";
    std::vector<T32Prob> tp(n);
    int tiles = 0, bn_max = 16, kb_max = 1;
    bool headout = false, plain = false;
    // n tiles of at most 64 columns: accumulators of 64 TMEM columns leave room for a chunk ring of 6 (128 columns: 2),
    // and the deep-K / small-M problems of the late blocks get twice the CTAs
    const int bn_cap = std::getenv("HMDPOSE_TF32_BN") ? std::max(16, std::min(128, std::atoi(std::getenv("HMDPOSE_TF32_BN")))) : 128;
    cudaStream_t ss = nullptr;   // the weight split runs on its own stream (other threads may be capturing graphs)
    HP_CUDA(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
    for (int i = 0; i < n; ++i) {
      GemmProb& p = probs[i];
      if (p.K % 4 != 0 || p.lda % 4 != 0)
        throw Error(HMDPOSE_E_STATE, "tf32 GEMM needs K and lda multiples of 4 (16-byte TMA pitch)");
      p.bn = gemm_choose_bn(p.N, &p.n_tiles, bn_cap);
      p.m_tiles = cdiv(p.M, TC_BM);
      p.tile_start = tiles;
      tiles += p.m_tiles * p.n_tiles;
      bn_max = std::max(bn_max, p.bn);
      kb_max = std::max(kb_max, cdiv(p.K, T32_BK));
      headout = headout || p.out_mode != 0;
      plain = plain || p.out_mode == 0;
      // W -> (W_hi, W_lo) once per plan
      const long long nw = (long long)p.N * p.K;
      float *whi = nullptr, *wlo = nullptr;
      HP_CUDA(cudaMalloc(&whi, (size_t)nw * 4));
      owned.push_back(whi);
      HP_CUDA(cudaMalloc(&wlo, (size_t)nw * 4));
      owned.push_back(wlo);
      split_tf32_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, ss>>>((const float*)p.W, whi, wlo, nw);
      HP_CUDA(cudaGetLastError());
      std::memset(&tp[i], 0, sizeof(T32Prob));
      encode_2d(&tp[i].tmA, p.A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.lda * 4, T32_BK, TC_BM, true);
      encode_2d(&tp[i].tmBhi, whi, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.K * 4, T32_BK, (uint32_t)p.bn, true);
      encode_2d(&tp[i].tmBlo, wlo, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.K * 4, T32_BK, (uint32_t)p.bn, true);
      tp[i].p = p;
    }
    HP_CUDA(cudaStreamSynchronize(ss));
    cudaStreamDestroy(ss);
    if (headout && plain) throw Error(HMDPOSE_E_STATE, "head-tensor and NHWC outputs cannot share a GEMM launch");
    int ring_bytes = 0, res_bytes = 0;
    const bool reuse = tiles > g_num_sms && std::getenv("HMDPOSE_NO_BRES") == nullptr;   // some CTA walks more than one tile
    for (int i = 0; i < n; ++i) {
      const GemmProb& p = tp[i].p;
      const int panel = 2 * cdiv(p.K, T32_BK) * p.bn * 128;
      tp[i].p.b_res = (reuse && panel <= T32_RES_MAX) ? 1 : 0;
      if (tp[i].p.b_res) res_bytes = std::max(res_bytes, panel);
      else ring_bytes = std::max(ring_bytes, 2 * p.bn * 128);
    }
    T32Prob* d = nullptr;
    HP_CUDA(cudaMalloc(&d, sizeof(T32Prob) * n));
    owned.push_back(d);
    HP_CUDA(cudaMemcpy(d, tp.data(), sizeof(T32Prob) * n, cudaMemcpyHostToDevice));
    int stages = std::max(2, std::min(kb_max + 1, T32_MAX_STAGES));
    bool gated32 = false;
    for (int i = 0; i < n; ++i) gated32 = gated32 || tp[i].p.a_scale != nullptr;
    while (stages > 2 && t32_smem_bytes(stages, ring_bytes, res_bytes, gated32) > 224 * 1024) --stages;
    const int smem = t32_smem_bytes(stages, ring_bytes, res_bytes, gated32);
    if (smem > 226 * 1024) throw Error(HMDPOSE_E_STATE, "tf32 GEMM shared-memory budget exceeded");
    const int grid = std::min(tiles, g_num_sms);
    // k-blocks (of 32) per chunk accumulator: 1 = every 32 k's are summed in registers (see gemm_tf32.cuh)
    const char* ce = std::getenv("HMDPOSE_TF32_CHUNK");
    const int chunk = ce ? std::max(1, std::atoi(ce)) : 1;
    // expected truncation loss of the chunk accumulators, in units of 2^-24: debias0 + debias1 * (MMAs per chunk)
    float db0 = 0.2f, db1 = 0.31f;
    if (const char* de = std::getenv("HMDPOSE_TF32_DEBIAS")) std::sscanf(de, "%f,%f", &db0, &db1);
    if (headout)
      return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tf32_kernel<true>, dim3(grid), dim3(T32_THREADS), smem, st, d, n, tiles, bn_max, stages, ring_bytes, res_bytes, chunk, db0, db1)); };
    return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tf32_kernel<false>, dim3(grid), dim3(T32_THREADS), smem, st, d, n, tiles, bn_max, stages, ring_bytes, res_bytes, chunk, db0, db1)); };
  }
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    GemmProb& p = probs[i];
    p.bn = SG_BN;
    p.n_tiles = cdiv(p.N, SG_BN);
    p.m_tiles = cdiv(p.M, SG_BM);
    p.tile_start = tiles;
    tiles += p.m_tiles * p.n_tiles;
  }
  GemmProb* d = nullptr;
  HP_CUDA(cudaMalloc(&d, sizeof(GemmProb) * n));
  HP_CUDA(cudaMemcpy(d, probs.data(), sizeof(GemmProb) * n, cudaMemcpyHostToDevice));
  owned.push_back(d);
  if (fast) return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_simt_kernel<__half>, dim3(tiles), dim3(256), 0, st, d, n)); };
  return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_simt_kernel<float>, dim3(tiles), dim3(256), 0, st, d, n)); };
}

// Fused depthwise-separable conv (sepconv_tc.cuh).  SepSpec -> device table with the TMA descriptor of the
// pointwise weights; one CTA per 128 output pixels.
// implicit-GEMM path (sepconv3_tc.cuh): plain inputs with folded tap matrices, N <= 128, W + 2 <= 128
static bool sep3_ok(const SepSpec& q) {
  return q.w9 != nullptr && !q.fused && q.p.N <= (q.p.out_mode ? 64 : 128) && q.W + 2 <= 128 &&
         std::getenv("HMDPOSE_NO_SEP3") == nullptr;
}

static std::function<void(cudaStream_t)> make_sep3_launcher(std::vector<SepSpec> specs, std::vector<void*>& owned) {
  const int n = (int)specs.size();
  std::vector<Sep3Prob> sp(n);
  int tiles = 0, bn_max = 16, wp_max = 4;
  bool headout = false, plain = false;
  for (int i = 0; i < n; ++i) {
    SepSpec& q = specs[i];
    GemmProb& p = q.p;
    Sep3Prob& d = sp[i];
    std::memset(&d, 0, sizeof(Sep3Prob));
    p.K = 64;
    p.M = q.Bn * q.H * q.W;
    p.rows_per_img = q.H * q.W;
    p.bn = ((p.N + 15) / 16) * 16;
    p.n_tiles = 1;
    d.H = q.H; d.W = q.W; d.Bn = q.Bn; d.Wp = q.W + 2;
    const int blk = (((q.H + 2) * d.Wp + 7) / 8) * 8;
    if (2 * blk <= 128) {   // several whole images per tile
      d.ipt = 128 / blk; d.blk = blk; d.R = q.H; d.tpi = 1; d.box_rows = q.H + 2;
      d.n_tiles = cdiv(q.Bn, d.ipt);
    } else {
      d.ipt = 1; d.blk = blk; d.R = std::min(q.H, 128 / d.Wp); d.tpi = cdiv(q.H, d.R); d.box_rows = d.R + 2;
      d.n_tiles = q.Bn * d.tpi;
    }
    d.box_bytes = 128 * d.Wp * d.box_rows;
    d.inv_wp = (65536 + d.Wp - 1) / d.Wp;
    d.inv_blk = (65536 + d.blk - 1) / d.blk;
    d.tile_start = tiles;
    tiles += d.n_tiles;
    p.tile_start = d.tile_start;
    bn_max = std::max(bn_max, p.bn);
    wp_max = std::max(wp_max, d.Wp);
    headout = headout || p.out_mode != 0;
    plain = plain || p.out_mode == 0;
    encode_act_4d(&d.tmIn, q.in, true, 64, q.W, q.H, q.Bn, 64, d.Wp, d.box_rows, true);
    encode_2d(&d.tmW, q.w9, 64, (uint64_t)9 * p.N, 128, 64, (uint32_t)p.bn);
    d.wkey = q.w9;
    d.scale = q.scale;
    d.p = p;
  }
  if (headout && plain) throw Error(HMDPOSE_E_STATE, "head-tensor and NHWC outputs cannot share a sepconv launch");
  Sep3Prob* dev = nullptr;
  HP_CUDA(cudaMalloc(&dev, sizeof(Sep3Prob) * n));
  HP_CUDA(cudaMemcpy(dev, sp.data(), sizeof(Sep3Prob) * n, cudaMemcpyHostToDevice));
  owned.push_back(dev);
  const int stage_bytes = (((130 + 2 * wp_max) * 128 + 1023) / 1024) * 1024;
  int stages = S3_MAX_STAGES;
  while (stages > 2 && sep3_smem_bytes(bn_max, headout, stage_bytes, stages) > 226 * 1024) --stages;
  const int smem = sep3_smem_bytes(bn_max, headout, stage_bytes, stages);
  if (smem > 226 * 1024) throw Error(HMDPOSE_E_STATE, "sepconv3 shared-memory budget exceeded");
  int grid = std::min(tiles, g_num_sms);
  if (const char* e = std::getenv("HMDPOSE_SEP3_GRID")) grid = std::min(tiles, std::atoi(e));
  if (const char* e = std::getenv("HMDPOSE_MAX_TILES_PER_CTA")) grid = std::min(tiles, std::max(grid, cdiv(tiles, std::max(1, std::atoi(e)))));
  if (headout)
    return [=](cudaStream_t st) { HP_CUDA(launch_k(sepconv3_kernel<true>, dim3(grid), dim3(S3_THREADS), smem, st, dev, n, tiles, bn_max, stage_bytes, stages)); };
  return [=](cudaStream_t st) { HP_CUDA(launch_k(sepconv3_kernel<false>, dim3(grid), dim3(S3_THREADS), smem, st, dev, n, tiles, bn_max, stage_bytes, stages)); };
}

// debug: timeline of CTA 0 of the last sepconv3 launch, microseconds relative to kernel entry
int sep3_debug_timeline(float* out, int cap) {
  unsigned long long ts[32];
  if (cudaMemcpyFromSymbol(ts, g_s3_ts, sizeof(ts)) != cudaSuccess) return 0;
  const int n = std::min(cap, 32);
  for (int i = 0; i < n; ++i) {
    const unsigned long long t0 = ts[i < 16 ? 0 : 16];
    out[i] = ts[i] >= t0 ? (float)((double)(ts[i] - t0) * 1e-3) : -1.f;
  }
  return n;
}

std::function<void(cudaStream_t)> make_sepconv_launcher(std::vector<SepSpec> all_specs, std::vector<void*>& owned,
                                                        const char** kernel_name) {
  init_gemm_kernels();
  std::vector<SepSpec> specs, specs3;
  for (const SepSpec& q : all_specs) (sep3_ok(q) ? specs3 : specs).push_back(q);
  if (kernel_name) *kernel_name = specs3.empty() ? "sepconv_kernel" : "sepconv3_kernel";
  if (!specs3.empty()) {
    auto l3 = make_sep3_launcher(specs3, owned);
    if (specs.empty()) return l3;
    auto l2 = make_sepconv_launcher(specs, owned, nullptr);
    return [=](cudaStream_t st) { l3(st); l2(st); };
  }
  const int n = (int)specs.size();
  std::vector<SepProb> sp(n);
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    SepSpec& q = specs[i];
    GemmProb& p = q.p;
    if (q.W > 128 || (q.H * q.W >= 128 ? (128 % q.W) != 0 || (q.H * q.W) % 128 != 0 : 128 % (q.H * q.W) != 0))
      throw Error(HMDPOSE_E_STATE, "sepconv tile geometry needs power-of-two feature maps (power-of-two image_size)");
    p.K = 64;
    p.M = q.Bn * q.H * q.W;
    p.rows_per_img = q.H * q.W;
    p.bn = gemm_choose_bn(p.N, &p.n_tiles);
    p.m_tiles = cdiv(p.M, 128);
    p.tile_start = tiles;
    tiles += p.m_tiles;
    std::memset(&sp[i], 0, sizeof(SepProb));
    encode_2d(&sp[i].tmW, p.W, 64, (uint64_t)p.N, 128, 64, (uint32_t)p.bn);
    sp[i].p = p;
    sp[i].in = q.in; sp[i].fb = q.fb; sp[i].fc = q.fc; sp[i].dw_w = q.dw_w;
    sp[i].H = q.H; sp[i].W = q.W; sp[i].Bn = q.Bn; sp[i].fused = q.fused; sp[i].mode_b = q.mode_b; sp[i].mode_c = q.mode_c;
    sp[i].w0 = q.w0; sp[i].w1 = q.w1; sp[i].w2 = q.w2;
  }
  SepProb* d = nullptr;
  HP_CUDA(cudaMalloc(&d, sizeof(SepProb) * n));
  HP_CUDA(cudaMemcpy(d, sp.data(), sizeof(SepProb) * n, cudaMemcpyHostToDevice));
  owned.push_back(d);
  int bn_max = 16;
  for (const SepProb& q : sp) bn_max = std::max(bn_max, q.p.bn);
  bn_max = bn_max <= 64 ? 64 : 128;   // swizzled tiles stay 1024-byte aligned
  bool d0_state = false;
  for (const SepProb& q : sp) d0_state = d0_state || q.p.out_mode == 2;
  const int smem = sep_smem_bytes(bn_max, d0_state);
  return [=](cudaStream_t st) { HP_CUDA(launch_k(sepconv_kernel<false>, dim3(tiles), dim3(SEP_THREADS), smem, st, d, n, bn_max, 0, 0)); };
}

// Chain launch (sepconv_tc.cuh): CTA g runs all `specs` in order for the images [g*nb, (g+1)*nb)
std::function<void(cudaStream_t)> make_sepconv_chain_launcher(std::vector<SepSpec> specs, int nb, std::vector<void*>& owned) {
  init_gemm_kernels();
  const int n = (int)specs.size();
  std::vector<SepProb> sp(n);
  int bn_max = 16;
  for (int i = 0; i < n; ++i) {
    SepSpec& q = specs[i];
    GemmProb& p = q.p;
    if (q.H * q.W * nb > 128 || (q.H * q.W & (q.H * q.W - 1)) != 0 || (q.W & (q.W - 1)) != 0 || q.Bn != specs[0].Bn)
      throw Error(HMDPOSE_E_STATE, "sepconv chain needs power-of-two maps of at most 128 pixels per CTA");
    p.K = 64;
    p.M = q.Bn * q.H * q.W;
    p.rows_per_img = q.H * q.W;
    p.bn = gemm_choose_bn(p.N, &p.n_tiles);
    p.m_tiles = cdiv(q.Bn, nb);
    p.tile_start = 0;
    std::memset(&sp[i], 0, sizeof(SepProb));
    encode_2d(&sp[i].tmW, p.W, 64, (uint64_t)p.N, 128, 64, (uint32_t)p.bn);
    sp[i].p = p;
    sp[i].in = q.in; sp[i].fb = q.fb; sp[i].fc = q.fc; sp[i].dw_w = q.dw_w;
    sp[i].H = q.H; sp[i].W = q.W; sp[i].Bn = q.Bn; sp[i].fused = q.fused; sp[i].mode_b = q.mode_b; sp[i].mode_c = q.mode_c;
    sp[i].w0 = q.w0; sp[i].w1 = q.w1; sp[i].w2 = q.w2;
    bn_max = std::max(bn_max, p.bn);
  }
  SepProb* d = nullptr;
  HP_CUDA(cudaMalloc(&d, sizeof(SepProb) * n));
  HP_CUDA(cudaMemcpy(d, sp.data(), sizeof(SepProb) * n, cudaMemcpyHostToDevice));
  owned.push_back(d);
  bn_max = bn_max <= 64 ? 64 : 128;
  const int smem = sep_smem_bytes(bn_max);
  const int grid = cdiv(specs[0].Bn, nb);
  return [=](cudaStream_t st) { HP_CUDA(launch_k(sepconv_kernel<true>, dim3(grid), dim3(SEP_THREADS), smem, st, d, n, bn_max, n, nb)); };
}

}  // namespace hp

namespace hp {

int mb_debug_timeline(float* out, int cap) {
  unsigned long long ts[32];
  if (cudaMemcpyFromSymbol(ts, g_mb_ts, sizeof(ts)) != cudaSuccess) return 0;
  const int n = std::min(cap, 32);
  for (int i = 0; i < n; ++i) out[i] = ts[i] >= ts[0] ? (float)((double)(ts[i] - ts[0]) * 1e-3) : -1.f;
  return n;
}

// Fused MBConv block (mbconv_tc.cuh): one cluster of 4 or 6 CTAs per image.  Returns an empty function when the block does
// not fit the kernel (the caller then keeps the four-launch path).  `part` = the handle's split-K scratch.
std::function<void(cudaStream_t)> make_mbconv_launcher(MbSpec sp, int batch, std::vector<void*>& owned, float* part,
                                                       size_t part_bytes) {
  if (!mb_plan(sp)) return nullptr;
  if (mb_part_bytes(sp, batch) > part_bytes) return nullptr;
  void (*kern)(const MbSpec) = nullptr;
  const int spx = sp.Wo == 16 ? 4 : 1;
  if (sp.k == 3 && sp.stride == 1 && spx == 4) kern = mbconv_fused_kernel<3, 1, 4>;
  else if (sp.k == 5 && sp.stride == 1 && spx == 4) kern = mbconv_fused_kernel<5, 1, 4>;
  else if (sp.k == 5 && sp.stride == 2 && spx == 1) kern = mbconv_fused_kernel<5, 2, 1>;
  else if (sp.k == 5 && sp.stride == 1 && spx == 1) kern = mbconv_fused_kernel<5, 1, 1>;
  else if (sp.k == 3 && sp.stride == 1 && spx == 1) kern = mbconv_fused_kernel<3, 1, 1>;
  else if (sp.k == 3 && sp.stride == 2 && spx == 1) kern = mbconv_fused_kernel<3, 2, 1>;
  else return nullptr;
  init_gemm_kernels();
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.smem_bytes));
  // tensor maps: x rows of all images, W_exp, W_proj (K-major, SWIZZLE_128B, out-of-bounds = zero fill)
  CUtensorMap tm[3];
  encode_2d(&tm[0], sp.x, (uint64_t)sp.cin, (uint64_t)batch * sp.P, (uint64_t)sp.cin * 2, 64, (uint32_t)std::min(sp.P, 128));
  encode_2d(&tm[1], sp.w_exp, (uint64_t)sp.cin, (uint64_t)sp.cexp, (uint64_t)sp.cin * 2, 64, 64);
  encode_2d(&tm[2], sp.w_proj, (uint64_t)sp.cexp, (uint64_t)sp.cout, (uint64_t)sp.cexp * 2, 64,
            (uint32_t)(sp.cout > 256 ? sp.cout / 2 : sp.cout));
  CUtensorMap* d_tm = nullptr;
  HP_CUDA(cudaMalloc(&d_tm, sizeof(tm)));
  owned.push_back(d_tm);
  HP_CUDA(cudaMemcpy(d_tm, tm, sizeof(tm), cudaMemcpyHostToDevice));
  sp.tm = d_tm;
  sp.part = part;
  const int smem = sp.smem_bytes;
  const int cl = sp.cl;
  auto launch = [=](cudaStream_t st, int* max_clusters) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(batch * cl); cfg.blockDim = dim3(MB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (max_clusters) return cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg);
    return cudaLaunchKernelEx(&cfg, kern, sp);
  };
  if (std::getenv("HMDPOSE_DEBUG") != nullptr) {
    int ncl = -1;
    cudaError_t e = launch(nullptr, &ncl);
    std::fprintf(stderr, "[hmdpose] mbconv k%d s%d P=%d cin=%d cexp=%d cout=%d: cluster %d x %d slices, smem %d B, tmem %d cols, max active clusters %d (%s)\n",
                 sp.k, sp.stride, sp.P, sp.cin, sp.cexp, sp.cout, cl, sp.nmine, smem, sp.tmem_cols, ncl, cudaGetErrorString(e));
  }
  return [=](cudaStream_t st) { HP_CUDA(launch(st, nullptr)); };
}

// Fused expand + depthwise kernel of the large feature maps (expdw_tc.cuh).  Returns an empty function when the block
// does not fit the kernel (the caller keeps the expand GEMM + dw3_kernel pair); *tiles_per_img = squeeze partials per image.
std::function<void(cudaStream_t)> make_expdw_launcher(EdSpec sp, std::vector<void*>& owned, int* tiles_per_img) {
  if (!ed_plan(sp)) return nullptr;
  void (*kern)(const EdSpec) = nullptr;
  if (sp.k == 3 && sp.stride == 1) kern = expdw_kernel<3, 1, 4>;
  else if (sp.k == 5 && sp.stride == 1) kern = expdw_kernel<5, 1, 3>;
  else if (sp.k == 3 && sp.stride == 2) kern = expdw_kernel<3, 2, 2>;
  else if (sp.k == 5 && sp.stride == 2) kern = expdw_kernel<5, 2, 2>;
  else return nullptr;
  init_gemm_kernels();
  HP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.smem_bytes));
  CUtensorMap tm[2];
  encode_act_4d(&tm[0], sp.x, true, sp.cin, sp.W, sp.H, sp.B, 64, ED_WIN, ED_WIN, true);
  encode_2d(&tm[1], sp.w_exp, (uint64_t)sp.cin, (uint64_t)sp.cexp, (uint64_t)sp.cin * 2, 64, (uint32_t)sp.cexp);
  CUtensorMap* d_tm = nullptr;
  HP_CUDA(cudaMalloc(&d_tm, sizeof(tm)));
  owned.push_back(d_tm);
  HP_CUDA(cudaMemcpy(d_tm, tm, sizeof(tm), cudaMemcpyHostToDevice));
  sp.tm = d_tm;
  if (tiles_per_img) *tiles_per_img = sp.tiles_per_img;
  const int grid = std::min(sp.total_tiles, g_num_sms);
  const int smem = sp.smem_bytes;
  if (std::getenv("HMDPOSE_DEBUG") != nullptr)
    std::fprintf(stderr, "[hmdpose] expdw k%d s%d %dx%d cin=%d cexp=%d: %d tiles (%d per image, %dx%d outputs each), grid %d, smem %d B\n",
                 sp.k, sp.stride, sp.H, sp.W, sp.cin, sp.cexp, sp.total_tiles, sp.tiles_per_img, sp.TO, sp.TO, grid, smem);
  return [=](cudaStream_t st) { HP_CUDA(launch_k(kern, dim3(grid), dim3(ED_THREADS), smem, st, sp)); };
}

bool projk_fits(PkSpec sp, size_t part_bytes) {
  init_gemm_kernels();
  return pk_plan(sp, g_num_sms) && pk_part_bytes(sp) <= part_bytes;
}

// Split-K project GEMM of the small maps (projk_tc.cuh).  Returns an empty function when the problem does not fit the
// kernel (the caller keeps the gated gemm_tc2 launch).  `part` = the handle's split-K scratch.
std::function<void(cudaStream_t)> make_projk_launcher(PkSpec sp, const void* w, std::vector<void*>& owned, float* part,
                                                      size_t part_bytes) {
  init_gemm_kernels();
  if (!pk_plan(sp, g_num_sms)) return nullptr;
  if (pk_part_bytes(sp) > part_bytes) return nullptr;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    HP_CUDA(cudaFuncSetAttribute(projk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  CUtensorMap tm[2];
  encode_2d(&tm[0], sp.a, (uint64_t)sp.K, (uint64_t)sp.M, (uint64_t)sp.K * 2, 64, 128);
  encode_2d(&tm[1], w, (uint64_t)sp.K, (uint64_t)sp.N, (uint64_t)sp.K * 2, 64, (uint32_t)(sp.N > 256 ? sp.N / 2 : sp.N));
  CUtensorMap* d_tm = nullptr;
  HP_CUDA(cudaMalloc(&d_tm, sizeof(tm)));
  owned.push_back(d_tm);
  HP_CUDA(cudaMemcpy(d_tm, tm, sizeof(tm), cudaMemcpyHostToDevice));
  sp.tm = d_tm;
  sp.part = part;
  const int smem = sp.smem_bytes, S = sp.S, grid = sp.m_tiles * sp.S;
  if (std::getenv("HMDPOSE_DEBUG") != nullptr)
    std::fprintf(stderr, "[hmdpose] projk M=%d N=%d K=%d: %d m tiles x cluster %d (%d k-blocks each), smem %d B, tmem %d cols\n",
                 sp.M, sp.N, sp.K, sp.m_tiles, S, sp.nmine, smem, sp.tmem_cols);
  return [=](cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PK_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    HP_CUDA(cudaLaunchKernelEx(&cfg, projk_kernel, sp));
  };
}

}  // namespace hp
