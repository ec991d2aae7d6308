// Pointwise-conv GEMM dispatch: FFMA kernel (parity mode / cross-check) or tcgen05+TMA kernel (fast mode).
#include <cudaTypedefs.h>

#include <cstring>

#include "engine.h"
#include "gemm_tc.cuh"
#include "sepconv_tc.cuh"

namespace hp {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_num_sms = 148;

void init_gemm_kernels() {
  static std::once_flag once;
  std::call_once(once, [] {
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
    HP_CUDA(cudaFuncSetAttribute(sepconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sep_smem_bytes(128)));
    int dev = 0;
    HP_CUDA(cudaGetDevice(&dev));
    HP_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  });
}

// Split N into tiles of width bn (multiple of 16, <= 128) wasting as few padded columns as possible.
int gemm_choose_bn(int N, int* n_tiles) {
  int best_bn = 128, best_t = cdiv(N, 128), best_waste = best_t * 128 - N;
  for (int t = cdiv(N, 128); t <= cdiv(N, 128) + 3; ++t) {
    int bn = ((cdiv(N, t) + 15) / 16) * 16;
    if (bn > 128) continue;
    int waste = t * bn - N;
    if (waste < best_waste) { best_waste = waste; best_bn = bn; best_t = t; }
  }
  *n_tiles = best_t;
  return best_bn;
}

static void encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(HMDPOSE_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") inner=" +
                                    std::to_string(inner) + " outer=" + std::to_string(outer) + " pitch=" +
                                    std::to_string(row_bytes));
}

// 4-D tensor map of an NHWC activation tensor (C, W, H, B innermost first), un-swizzled box (box_c, box_w, box_h, 1).
void encode_act_4d(CUtensorMap* tm, const void* base, bool is_half, int C, int W, int H, int B, int box_c, int box_w,
                   int box_h) {
  init_gemm_kernels();
  const uint64_t es = is_half ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
      encode_2d(&tp[i].tmB, p.W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.K * 2, TC_BK, (uint32_t)p.bn);
      tp[i].p = p;
    }
    TcProb* d = nullptr;
    HP_CUDA(cudaMalloc(&d, sizeof(TcProb) * n));
    HP_CUDA(cudaMemcpy(d, tp.data(), sizeof(TcProb) * n, cudaMemcpyHostToDevice));
    owned.push_back(d);
    if (!v1) {
      bool gated = false;
      for (const GemmProb& p : probs) gated = gated || p.a_scale != nullptr;
      // ring depth: enough k-blocks in flight to hide the TMA latency of deep-K problems (K = 1152 -> 18 k-blocks);
      // shallow-K launches keep 2 stages so that two CTAs stay resident per SM
      int stages = std::max(2, std::min(kb_max, TC2_MAX_STAGES));
      while (stages > 2 && tc2_smem_bytes(bn_max, stages, gated) > 200 * 1024) --stages;
      const int smem2 = tc2_smem_bytes(bn_max, stages, gated);
      const int per_sm = std::max(1, std::min(2, (227 * 1024) / (smem2 + 1024)));
      const int grid = std::min(tiles, per_sm * g_num_sms);
      const int threads = gated ? TC2_THREADS_GATED : TC2_THREADS;
      bool headout = false, plain = false;
      for (const GemmProb& p : probs) { headout = headout || p.out_mode != 0; plain = plain || p.out_mode == 0; }
      if (headout && (plain || gated)) throw Error(HMDPOSE_E_STATE, "head-tensor and NHWC outputs cannot share a GEMM launch");
      if (headout)
        return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<false, true>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages)); };
      if (gated)
        return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<true, false>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages)); };
      return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc2_kernel<false, false>, dim3(grid), dim3(threads), smem2, st, d, n, tiles, bn_max, stages)); };
    }
    // stages: enough to cover K, capped so that >= 2 CTAs fit per SM
    int stages = std::min(kb_max, TC_MAX_STAGES);
    while (stages > 2 && tc_smem_bytes(stages, bn_max) > 100 * 1024) --stages;
    const int smem = tc_smem_bytes(stages, bn_max);
    return [=](cudaStream_t st) { HP_CUDA(launch_k(gemm_tc_kernel, dim3(tiles), dim3(TC_THREADS), smem, st, d, n, stages, bn_max)); };
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
std::function<void(cudaStream_t)> make_sepconv_launcher(std::vector<SepSpec> specs, std::vector<void*>& owned) {
  init_gemm_kernels();
  const int n = (int)specs.size();
  std::vector<SepProb> sp(n);
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    SepSpec& q = specs[i];
    GemmProb& p = q.p;
    if (q.W > 128 || (q.H * q.W >= 128 ? (128 % q.W) != 0 || (q.H * q.W) % 128 != 0 : 128 % (q.H * q.W) != 0))
      throw Error(HMDPOSE_E_STATE, "sepconv tile geometry needs power-of-two maps (S multiple of 128)");
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
  const int smem = sep_smem_bytes(bn_max);
  return [=](cudaStream_t st) { HP_CUDA(launch_k(sepconv_kernel, dim3(tiles), dim3(SEP_THREADS), smem, st, d, n, bn_max)); };
}

}  // namespace hp
