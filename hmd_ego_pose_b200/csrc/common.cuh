// Shared device/host definitions for libhmdpose (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pdl.cuh"

namespace hp {

enum Act { ACT_NONE = 0, ACT_SWISH = 1, ACT_SIGMOID = 2 };

// ---------------------------------------------------------------------------------------------
// 16-byte vectors of the activation storage type T (float: 4 lanes, __half: 8 lanes); all channel
// counts of the network are multiples of 8, so NHWC rows are always 16-byte aligned.
// ---------------------------------------------------------------------------------------------
template <typename T> struct VecN;
template <> struct VecN<float> { static constexpr int N = 4; };
template <> struct VecN<__half> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void ldv(const T* p, float* v);
template <> __device__ __forceinline__ void ldv<float>(const float* p, float* v) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<__half>(const __half* p, float* v) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void stv(T* p, const float* v);
template <> __device__ __forceinline__ void stv<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void stv<__half>(__half* p, const float* v) {
  uint4 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// x*sigmoid(x) (efficientnet/utils.py:57-59).  T = float is the parity mode: IEEE division and the
// accurate expf; T = __half is the fast mode: SFU exp + approximate reciprocal.
template <typename T> __device__ __forceinline__ float sigmoid_t(float x);
template <> __device__ __forceinline__ float sigmoid_t<float>(float x) { return 1.0f / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ float sigmoid_t<__half>(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// swish: parity mode evaluates x*sigmoid(x) as written; fast mode uses x*sigmoid(x) = h + h*tanh(h), h = x/2
// with tanh.approx.f32 (one MUFU op, rel. error 2^-11 -- below the fp16 rounding of the stored result).
template <typename T> __device__ __forceinline__ float swish_t(float x);
template <> __device__ __forceinline__ float swish_t<float>(float x) { return x * sigmoid_t<float>(x); }
template <> __device__ __forceinline__ float swish_t<__half>(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
template <typename T> __device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SWISH) return swish_t<T>(x);
  if (act == ACT_SIGMOID) return sigmoid_t<T>(x);
  return x;
}

// ---------------------------------------------------------------------------------------------
// Pointwise-conv-as-GEMM problem: D[M,N] = act(A[M,K] * diag(a_scale[img]) * W[N,K]^T + bias) (+res)
// One launch executes a table of problems (grouped GEMM): 5 pyramid levels x 5 heads share a launch.
// ---------------------------------------------------------------------------------------------
struct GemmProb {
  const void* A;         // [M, lda] activations, NHWC rows
  const void* W;         // [N, K] weights (BN folded), K contiguous
  const float* bias;     // [N]
  const float* a_scale;  // [images, K] squeeze-excite gate applied to A's columns, or null
  const void* residual;  // [M, N] skip connection, or null
  void* out;
  int M, N, K, lda, ldo;
  int act;
  int rows_per_img;      // H*W of this feature map (a_scale row / header image lookup)
  // out_mode 0: T out[m*ldo + n].  out_mode 1: fp32 head tensor (B, Nanchors, P) written in the
  // reference's permute(0,2,3,1).view(B,-1,P) order: channel n = a*p_src + p of pixel m goes to
  // out + (m / rows_per_img)*img_stride + (m % rows_per_img)*pix_stride + (n / p_src)*p_dst + p_off + n % p_src
  int out_mode, p_src, p_dst, p_off, pix_stride;
  long long img_stride;
  int m_tiles, n_tiles, bn, tile_start;
  int accumulate;        // out_mode 1 only: add to the head tensor instead of overwriting it (iterative refinement delta)
  int b_res;             // tcgen05 path: the [bn x K] weight panel of an n tile stays resident in shared memory
};

struct __align__(64) TcProb {
  CUtensorMap tmA;  // A: dims {K, M}, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmB;  // W: dims {K, N}, box {64, bn}, SWIZZLE_128B
  GemmProb p;
};

// Depthwise stencil group (one launch runs a table: backbone block = 1 group, head layer = 25 groups)
struct DwGroup {
  const void* in; void* out;
  const float* w;      // [k*k][C] tap-major (BN folded)
  const float* bias;   // [C] or null
  float* se_partial;   // [B][tiles_per_img][C] per-tile channel sums of the OUTPUT (squeeze), or null
  int H, W, Ho, Wo, C, k, stride, pad;
  int act;
  int tiles_per_img, cv_chunks, cvb, block_start, nblocks;
  // v2 tiling: a block = cvb channel vectors x sw strips (4 output pixels each) x sh rows, looping over
  // `rep` vertically adjacent tiles; tiles_per_img = tiles_x * ceil(tiles_y / rep) squeeze partials
  int sw, sh, tiles_x, tiles_y, rep;
  // v3 (shared-memory tiled): block = th x tw output pixels x cb channel vectors; ns = th*tw/4 strips
  int th, tw, cb;
  const CUtensorMap* tmap;   // 4-D map of the input (C, W, H, B), box (cb*V, iwd, ih, 1): one TMA per block tile
  // fused BiFPN node input (dw2 FUSED variant): in = swish(w0*in + w1*resample(fb) + w2*resample(fc))
  const void* fb; const void* fc;
  int mode_b, mode_c;
  float w0, w1, w2;
  // squeeze-excite folded into the depthwise kernel (dw3): the LAST block of an image to finish (se_counter) reduces
  // the per-tile sums and evaluates the two FC layers of efficientnet/model.py:88-93 into se_gate[b][C]
  int* se_counter;        // [B], zero between launches (the last block resets it), or null
  const float* se_wr; const float* se_br;    // se_reduce [Cse][C], [Cse]
  const float* se_weT; const float* se_be;   // se_expand transposed [Cse][C], [C]
  float* se_gate;
  int se_cse;
  float se_inv_hw;
};

// BiFPN node input: out = swish(w0*a + w1*resample(b) + w2*resample(c))  (efficientdet/model.py:215-264)
enum Resample { RS_NONE = 0, RS_SAME = 1, RS_UP2 = 2, RS_POOL = 3 };
struct FuseArgs {
  const void* a; const void* b; const void* c; void* out;
  int B, H, W, C, mode_b, mode_c;
  float w0, w1, w2;
};

// Input of an iterative refinement sub-net (hmdegopose/model.py:76-80,147-150,214-218): torch.cat((feat, estimate), 1)
// as an NHWC tensor with the channel count padded to a multiple of 8 / 32 (zeros)
struct ConcatProb {
  const void* feat;     // [npix][64] trunk output, activation type
  const float* head;    // head tensor (B, N_anchors, Pw) fp32, already offset to this level
  void* out;            // [npix][Cpad] activation type
  int npix, HW, Cpad, P;   // P = 9 * Pw estimate channels per pixel
  int trans;            // translation: channels are xy (18: a*2 + j) then z (9: a), rows are (x, y, z) per anchor
  long long img_stride; // N_anchors * Pw
  int blk_start;
};

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace hp
