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

// Stem weights + bias as a kernel parameter (stem_kernel, kernels_simt.cuh): [tap (ky*3+kx)*3+ci][co], [co]
struct StemW {
  float w[27 * 32];
  float b[32];
};

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
  // out_mode 2 (sepconv_kernel, EfficientDet-d0 classifier header): the (B, N_anchors, C) scores are NEVER stored; the
  // epilogue keeps max / first arg-max over the p_src = C classes of every anchor and appends the anchors above
  // d0_thr as sort keys (what d0_max_kernel does from the materialised tensor; utils/utils.py:93-94,104-108)
  unsigned long long* d0_keys;   // [B][d0_cap]
  int* d0_cand_cls;              // [B][d0_ntot]
  int* d0_cand_count;            // [B], zeroed before the launch
  int d0_cap, d0_ntot, d0_anchor0;   // key capacity per image, anchors per image, first anchor of this pyramid level
  float d0_thr;
  int w_img_rows;        // > 0: W holds one [w_img_rows x K] panel PER IMAGE (squeeze-excite gate folded into the weights by
                         // se3_kernel: W'[img] = W . diag(gate[img])); rows_per_img must be a multiple of the 128-row m tile
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

// ---------------------------------------------------------------------------------------------
// Fused MBConv block for the small feature maps (mbconv_tc.cuh): kernel argument + shared-memory / TMEM plan
// ---------------------------------------------------------------------------------------------
constexpr int MB_CL_MAX = 8;        // CTAs per image (cluster size) is chosen per block: MbSpec::cl
constexpr int MB_WORKERS = 512;     // 16 worker warps (epilogues, stencil, squeeze-excite, reduction)
constexpr int MB_THREADS = MB_WORKERS + 32;   // + one warp that only issues TMA loads and tcgen05.mma
constexpr int MB_SLICE = 64;        // expanded channels per slice = one 128-byte swizzle row of fp16
constexpr int MB_MAX_MINE = 3;      // slices per CTA
constexpr int MB_STAGE_WARP_BYTES = 32 * 36 * 4;   // per-warp 32 x 36 fp32 transpose tile of the split-K partials

struct MbSpec {
  const __half* x; __half* out;
  const __half* w_exp; const float* b_exp;      // [cexp][cin], [cexp]
  const float* w_dw; const float* b_dw;         // [k*k][cexp], [cexp]
  const float* se_wr; const float* se_br;       // [cse][cexp], [cse]
  const float* se_weT; const float* se_be;      // [cse][cexp], [cexp]
  const __half* w_proj; const float* b_proj;    // [cout][cexp], [cout]
  __half* dbg_exp; __half* dbg_dw; float* gate_out;   // optional copies of the intermediates (HMDPOSE_KEEP_ALL / tests)
  const CUtensorMap* tm;                        // [3]: x {cin, B*P} box {64, min(P,128)}; W_exp {cin, cexp} box {64, 64};
                                                // W_proj {cexp, cout} box {64, cout or cout/2}  (all SWIZZLE_128B)
  float* part;                                  // [B][cl][Po][cout] fp32 split-K partials of the project GEMM (L2-resident scratch)
  int H, W, Ho, Wo, cin, cexp, cout, cse, k, stride, pad, skip;
  float inv_hw;
  // ---- plan (mb_plan) ----
  int P, Po, MT, MTo, KB1, nsl, nmine;
  int pitch, pitch_o;                 // bytes between M tiles of the x/exp and of the A2 operands (8 KB when <= 64 rows)
  int a2_slice_bytes, w2_slice_bytes;
  int off_w, off_exp, off_a2, off_dw, off_misc, smem_bytes;
  int cl;                             // cluster size: CTAs per image
  int rows_own;                       // output rows summed and written by each CTA of the cluster
  int d2_col0, d2_pitch, tmem_cols;
};

// fills the plan fields; false when the block does not fit this kernel (the launch plan then keeps the four-launch path)
inline bool mb_plan(MbSpec& s) {
  s.P = s.H * s.W; s.Po = s.Ho * s.Wo;
  if (s.H != s.W || (s.W != 8 && s.W != 16) || s.Ho != s.Wo || (s.Wo != 8 && s.Wo != 16)) return false;
  if (s.cin % 16 || s.cexp % 16 || s.cout % 16 || s.cout > 320 || s.cse > 64) return false;
  if (!((s.k == 3 || s.k == 5) && (s.stride == 1 || s.stride == 2))) return false;
  if (s.skip && (s.cin != s.cout || s.stride != 1)) return false;
  s.MT = (s.P + 127) / 128; s.MTo = (s.Po + 127) / 128;
  s.KB1 = (s.cin + 63) / 64;
  s.nsl = (s.cexp + MB_SLICE - 1) / MB_SLICE;
  // B200 runs at most 15 clusters of 8 CTAs at a time (cudaOccupancyMaxActiveClusters: one GPC has fewer than 16
  // SMs), a batch of 16 images would take two waves.  Clusters of <= 6 always fit in one wave: slices per CTA =
  // ceil(nsl / 6), then the smallest cluster that still reaches it (8 slices -> 4 CTAs x 2, 11 -> 6 x 2, 18 -> 6 x 3).
  const int cl_cap = s.cl > 0 ? s.cl : 6;
  s.nmine = (s.nsl + cl_cap - 1) / cl_cap;
  if (s.nmine > MB_MAX_MINE) return false;
  s.cl = (s.nsl + s.nmine - 1) / s.nmine;
  s.pitch = s.P <= 64 ? 8192 : 16384;
  s.pitch_o = s.Po <= 64 ? 8192 : 16384;
  s.a2_slice_bytes = s.MTo * s.pitch_o;
  s.w2_slice_bytes = s.cout * 128;
  int off = s.KB1 * s.MT * s.pitch;                   // x (an M = 128 MMA over a 64-row tile reads 8 KB past it: into the next region)
  // W_exp slices [64 rows][KB1 x 128 B] of all my slices; once the expand MMAs are done the W_proj slices land here
  s.off_w = off; off += s.nmine * (s.KB1 * 8192 > s.w2_slice_bytes ? s.KB1 * 8192 : s.w2_slice_bytes);
  s.off_exp = off; off += s.MT * s.pitch;             // expanded tile of the current slice
  s.off_a2 = off; off += s.nmine * s.a2_slice_bytes;  // depthwise outputs = A operand of the project GEMM
  if (s.pitch_o == 8192) off += 8192;                 // slack for the M = 128 over-read of the last slice
  s.off_dw = off; off += s.nmine * s.k * s.k * 64 * 4;   // depthwise taps of my slices
  s.off_misc = off; off += (MB_MAX_MINE * 64 * 6 + 16 * 64 + MB_CL_MAX * 64 + 64) * 4;
  s.smem_bytes = off + 1024;
  if (16 * MB_STAGE_WARP_BYTES > s.off_dw) return false;   // the transpose staging aliases the dead operands
  s.rows_own = (s.Po + s.cl - 1) / s.cl;
  s.d2_col0 = s.nmine * s.MT * 64;
  s.d2_pitch = ((s.cout + 31) / 32) * 32;
  const int cols = s.d2_col0 + s.MTo * s.d2_pitch;
  if (cols > 512) return false;
  s.tmem_cols = 32;
  while (s.tmem_cols < cols) s.tmem_cols <<= 1;
  return s.smem_bytes <= 227 * 1024;
}

inline size_t mb_part_bytes(const MbSpec& s, int batch) { return (size_t)batch * s.cl * s.Po * s.cout * 4; }

// ---------------------------------------------------------------------------------------------
// Split-K project GEMM of the small feature maps (projk_tc.cuh): kernel argument + plan
// ---------------------------------------------------------------------------------------------
constexpr int PK_WORKERS = 256;
constexpr int PK_THREADS = PK_WORKERS + 32;
constexpr int PK_MAX_MINE = 3;      // k-blocks of 64 per CTA

struct PkSpec {
  const __half* a;            // [M][K] depthwise output (fp16, NHWC rows)
  const float* gate;          // [images][K] squeeze-excite gate
  const float* bias;          // [N]
  const __half* residual;     // [M][N] or null
  __half* out;                // [M][N]
  float* part;                // [m_tiles][S][128][N] fp32 partial tiles (L2-resident scratch)
  const CUtensorMap* tm;      // [2]: A {K, M} box {64, 128}; W {K, N} box {64, N or N / 2}  (SWIZZLE_128B)
  int M, N, K, rows_per_img;
  // ---- plan (pk_plan) ----
  int S, nkb, nmine, m_tiles, rows_own, w_slice_bytes, off_w, off_gate, smem_bytes, d_pitch, tmem_cols;
};

inline bool pk_plan(PkSpec& s, int num_sms) {
  if (s.K % 16 || s.N % 16 || s.N > 320 || s.K < 384 || s.rows_per_img < 64) return false;   // a tile touches <= 3 images
  s.m_tiles = (s.M + 127) / 128;
  s.nkb = (s.K + 63) / 64;
  s.S = num_sms / s.m_tiles;
  if (s.S > 6) s.S = 6;
  if (s.S < 2) return false;
  s.nmine = (s.nkb + s.S - 1) / s.S;
  if (s.nmine > PK_MAX_MINE) return false;
  s.S = (s.nkb + s.nmine - 1) / s.nmine;         // smallest cluster that still reaches nmine k-blocks per CTA
  s.rows_own = (128 + s.S - 1) / s.S;
  s.w_slice_bytes = s.N * 128;
  int off = s.nmine * 16384;                      // A k-blocks [128 rows][128 B]
  s.off_w = off; off += s.nmine * s.w_slice_bytes;
  if (8 * MB_STAGE_WARP_BYTES > off) off = 8 * MB_STAGE_WARP_BYTES;   // the transpose staging aliases the dead operands
  s.off_gate = off; off += 3 * PK_MAX_MINE * 64 * 4;    // gate values of up to 3 images x my channels
  s.smem_bytes = off + 1024;
  s.d_pitch = ((s.N + 31) / 32) * 32;
  s.tmem_cols = 32;
  while (s.tmem_cols < s.d_pitch) s.tmem_cols <<= 1;
  return s.smem_bytes <= 226 * 1024 && s.tmem_cols <= 512;
}
inline size_t pk_part_bytes(const PkSpec& s) { return (size_t)s.m_tiles * s.S * 128 * s.N * 4; }

// ---------------------------------------------------------------------------------------------
// Fused expand + depthwise kernel for the LARGE feature maps (expdw_tc.cuh): kernel argument + plan
// ---------------------------------------------------------------------------------------------
constexpr int ED_WIN = 16;          // input window of a tile: 16 x 16 pixels = two UMMA M tiles of 128 rows
constexpr int ED_WORKERS = 384;     // 12 worker warps (152 registers each: the stencil keeps taps, inputs and accumulators in registers)
constexpr int ED_THREADS = ED_WORKERS + 32;   // + the issue warp

struct EdSpec {
  const __half* x; __half* out;                 // x [B][H][W][cin]; out = depthwise output [B][Ho][Wo][cexp]
  const __half* w_exp; const float* b_exp;      // [cexp][cin], [cexp]
  const float* w_dw; const float* b_dw;         // [k*k][cexp], [cexp]
  float* se_partial;                            // [B][tiles_per_img][cexp] per-tile channel sums of `out` (squeeze)
  __half* dbg_exp;                              // optional copy of the expanded tensor [B][H][W][cexp] (HMDPOSE_KEEP_ALL / tests)
  const CUtensorMap* tm;                        // [2]: x {cin, W, H, B} box {64, 16, 16, 1}; W_exp {cin, cexp} box {64, cexp}
  int B, H, W, Ho, Wo, cin, cexp, k, stride, pad;
  // ---- plan (ed_plan) ----
  int TO;                                       // output pixels per tile side: (16 - k) / stride + 1
  int tiles_x, tiles_y, tiles_per_img, total_tiles;
  int ksteps, nchunk, e_pitch;                  // UMMA k steps of 16; 16-byte channel chunks; bytes per pixel of the expanded tile
  int off_w1, off_e, off_dw, off_b, off_red, smem_bytes;   // off_red < 0: the squeeze scratch aliases the expanded tile
};

inline int ed_tile_out(int k, int stride) { return (ED_WIN - k) / stride + 1; }
inline int ed_tiles_per_img(int k, int stride, int Ho, int Wo) {
  const int to = ed_tile_out(k, stride);
  return ((Ho + to - 1) / to) * ((Wo + to - 1) / to);
}
inline bool ed_plan(EdSpec& s) {
  if (!((s.k == 3 || s.k == 5) && (s.stride == 1 || s.stride == 2))) return false;
  if (s.cin % 8 || s.cin > 64 || s.cexp % 16 || s.cexp > 256) return false;
  s.TO = ed_tile_out(s.k, s.stride);
  s.tiles_x = (s.Wo + s.TO - 1) / s.TO; s.tiles_y = (s.Ho + s.TO - 1) / s.TO;
  s.tiles_per_img = s.tiles_x * s.tiles_y;
  s.total_tiles = s.B * s.tiles_per_img;
  s.ksteps = (s.cin + 15) / 16;
  s.nchunk = s.cexp / 8;
  s.e_pitch = s.cexp * 2 + 16;                  // odd multiple of 16 bytes: consecutive pixels start in different bank groups
  int off = ED_WIN * ED_WIN * 128;              // x window: 256 pixel rows of 128 bytes (K-major SWIZZLE_128B operand)
  s.off_w1 = off; off += ((s.cexp * 128 + 1023) / 1024) * 1024;
  s.off_e = off; off += ED_WIN * ED_WIN * s.e_pitch;
  off = ((off + 127) / 128) * 128;
  s.off_dw = off; off += s.k * s.k * s.cexp * 4;
  s.off_b = off; off += 2 * s.cexp * 4;
  const int red_bytes = (ED_WORKERS / s.nchunk) * s.cexp * 4;
  s.off_red = -1;
  if (off + red_bytes + 1024 <= 227 * 1024) { s.off_red = off; off += red_bytes; }
  s.smem_bytes = off + 1024;
  return s.smem_bytes <= 227 * 1024;
}



}  // namespace hp
