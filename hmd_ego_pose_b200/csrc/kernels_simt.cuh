// CUDA-core kernels of the hot path: stem conv, depthwise stencils (+folded BN, swish, squeeze
// partial sums), squeeze-excite gate, BiFPN fusion/resampling, and the FFMA pointwise GEMM used by
// the fp32 parity mode.  Templated on the activation storage type T (float | __half); all
// arithmetic is fp32.  Layout: NHWC, 16-byte channel vectors.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "tma.cuh"

namespace hp {

// ---------------------------------------------------------------------------------------------
// Stem: 3x3 stride-2 conv 3->32, TF-SAME padding (0,1), folded BN, swish
// (efficientnet/model.py:141-142 applied at efficientdet/model.py:437-439; padding utils_extra.py:33-47)
// in : fp32 (B,3,S,S) with arbitrary element strides (NCHW or the reference's permuted NHWC view)
// w  : [27][32] tap-major ((ky*3+kx)*3+ci), bias [32];  out: T (B,S/2,S/2,32) NHWC
// One thread = one output pixel x 32 output channels.  The 27 x 32 weights and the bias are a KERNEL PARAMETER: every FFMA
// takes its weight straight from the constant bank (warp-uniform operand), so the inner loop has no load at all -- the
// shared-memory version issued one LDS.128 per four FFMAs and spent half of its stall samples waiting for them.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) stem_kernel(const float* __restrict__ in, long long sb, long long sc,
                                                   long long sh, long long sw, int B, int S,
                                                   const __grid_constant__ StemW wb, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int So = S / 2;
  const long long total = (long long)B * So * So;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int ox = (int)(pix % So);
  const int oy = (int)((pix / So) % So);
  const int b = (int)(pix / ((long long)So * So));
  // the 27 inputs of this pixel, all loads in flight together (zero beyond the bottom/right edge: SAME pad (0,1))
  float x[27];
  const float* ib = in + (long long)b * sb;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const int iy = 2 * oy + ky, ix = 2 * ox + kx;
        x[(ky * 3 + kx) * 3 + ci] = (iy < S && ix < S) ? __ldg(ib + ci * sc + iy * sh + ix * sw) : 0.f;
      }
  T* op = out + pix * 32;
  constexpr int V = VecN<T>::N;
#pragma unroll
  for (int g = 0; g < 4; ++g) {   // 8 output channels at a time
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = wb.b[g * 8 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[t], wb.w[t * 32 + g * 8 + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = apply_act<T>(acc[j], ACT_SWISH);
#pragma unroll
    for (int j = 0; j < 8; j += V) stv<T>(op + g * 8 + j, acc + j);
  }
}

// ---------------------------------------------------------------------------------------------
// Depthwise k x k stencil, stride 1|2, TF-SAME zero padding folded into the bounds test
// (MBConv: efficientnet/model.py:84-86 + _bn1 folded + swish; SeparableConvBlock dw: efficientdet/model.py:43).
// Block = cvb (channel vectors) x py (pixels); a block covers DW_TP consecutive output pixels of one
// image for cvb channel vectors.  If se_partial != null the block also emits the per-channel sum of its
// outputs (the squeeze of efficientnet/model.py:89), reduced in a fixed order -> deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int DW_TP = 64;       // output pixels per block tile
constexpr int DW_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(DW_THREADS) dw_kernel(const DwGroup* __restrict__ groups, int ngroups) {
  constexpr int V = VecN<T>::N;
  __shared__ float red[DW_THREADS * V];
  pdl_trigger();
  pdl_wait();
  int gi = 0;
  while (gi + 1 < ngroups && (int)blockIdx.x >= groups[gi + 1].block_start) ++gi;
  const DwGroup g = groups[gi];
  int blk = blockIdx.x - g.block_start;
  const int chunk = blk % g.cv_chunks; blk /= g.cv_chunks;
  const int tile = blk % g.tiles_per_img;
  const int b = blk / g.tiles_per_img;
  const int cvb = g.cvb;
  const int py = DW_THREADS / cvb;
  const int tx = threadIdx.x % cvb;
  const int ty = threadIdx.x / cvb;
  const int CV = g.C / V;
  const int cv = chunk * cvb + tx;
  const bool active = (ty < py) && (cv < CV);
  const int c0 = cv * V;
  const T* in = reinterpret_cast<const T*>(g.in) + (long long)b * g.H * g.W * g.C;
  T* out = reinterpret_cast<T*>(g.out) + (long long)b * g.Ho * g.Wo * g.C;
  const int npix = g.Ho * g.Wo;
  float se[V];
#pragma unroll
  for (int j = 0; j < V; ++j) se[j] = 0.f;
  if (active) {
    float bias[V];
#pragma unroll
    for (int j = 0; j < V; ++j) bias[j] = g.bias ? __ldg(g.bias + c0 + j) : 0.f;
    const int pend = min(npix, (tile + 1) * DW_TP);
    for (int p = tile * DW_TP + ty; p < pend; p += py) {
      const int oy = p / g.Wo, ox = p - oy * g.Wo;
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = bias[j];
      const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
      for (int ky = 0; ky < g.k; ++ky) {
        const int iy = iy0 + ky;
        if (iy < 0 || iy >= g.H) continue;
        for (int kx = 0; kx < g.k; ++kx) {
          const int ix = ix0 + kx;
          if (ix < 0 || ix >= g.W) continue;
          float v[V], wv[V];
          ldv<T>(in + ((long long)iy * g.W + ix) * g.C + c0, v);
          const float* wp = g.w + (ky * g.k + kx) * g.C + c0;
#pragma unroll
          for (int j = 0; j < V; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(wp + j));
            wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
          }
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = fmaf(v[j], wv[j], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        // round to the storage type first: the squeeze is taken over what the next layer reads
        acc[j] = to_f<T>(from_f<T>(apply_act<T>(acc[j], g.act)));
        se[j] += acc[j];
      }
      stv<T>(out + (long long)p * g.C + c0, acc);
    }
  }
  if (g.se_partial) {
#pragma unroll
    for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = se[j];
    __syncthreads();
    if (ty == 0 && cv < CV) {
      float s[V];
#pragma unroll
      for (int j = 0; j < V; ++j) s[j] = 0.f;
      for (int y = 0; y < py; ++y)
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] += red[(y * cvb + tx) * V + j];
      float* dst = g.se_partial + ((long long)b * g.tiles_per_img + tile) * g.C + c0;
#pragma unroll
      for (int j = 0; j < V; ++j) dst[j] = s[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Squeeze-excite gate: pooled = mean(x); r = swish(Wr*pooled + br); gate = sigmoid(We*r + be)
// (efficientnet/model.py:88-93).  One block per image; partial sums are added in tile order.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) se_kernel(const float* __restrict__ partial, int tiles, int C, int Cse,
                                                 float inv_hw, const float* __restrict__ wr,
                                                 const float* __restrict__ br, const float* __restrict__ we,
                                                 const float* __restrict__ be, float* __restrict__ gate) {
  __shared__ float pooled[1152];
  __shared__ float r[64];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const float* pp = partial + (long long)b * tiles * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < tiles; ++t) s += pp[(long long)t * C + c];
    pooled[c] = s * inv_hw;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cse; j += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(wr + (long long)j * C + c), pooled[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) r[j] = apply_act<T>(s + br[j], ACT_SWISH);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = be[c];
    for (int j = 0; j < Cse; ++j) s = fmaf(__ldg(we + (long long)c * Cse + j), r[j], s);
    gate[(long long)b * C + c] = sigmoid_t<T>(s);
  }
}

// ---------------------------------------------------------------------------------------------
// BiFPN resampling + fast-normalised fusion + swish (efficientdet/model.py:215-264)
//   RS_UP2 : nearest x2 upsample of a half-resolution map  (nn.Upsample, :96-99)
//   RS_POOL: MaxPool2dStaticSamePadding(3,2) of a double-resolution map: ZERO pad (0,1) then max
//            (utils_extra.py:72-86) -- the pad value takes part in the max at the right/bottom edge.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void fetch_rs(const T* src, int mode, int b, int y, int x, int H, int W, int C, int c0,
                                         float* v) {
  constexpr int V = VecN<T>::N;
  if (mode == RS_SAME) {
    ldv<T>(src + (((long long)b * H + y) * W + x) * C + c0, v);
  } else if (mode == RS_UP2) {
    const int Hs = H / 2, Ws = W / 2;
    ldv<T>(src + (((long long)b * Hs + (y >> 1)) * Ws + (x >> 1)) * C + c0, v);
  } else {  // RS_POOL
    const int Hs = H * 2, Ws = W * 2;
    bool first = true;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int sy = 2 * y + dy, sx = 2 * x + dx;
        float t[V];
        if (sy < Hs && sx < Ws) {
          ldv<T>(src + (((long long)b * Hs + sy) * Ws + sx) * C + c0, t);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) t[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = first ? t[j] : fmaxf(v[j], t[j]);
        first = false;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs a) {
  constexpr int V = VecN<T>::N;
  pdl_trigger();
  pdl_wait();
  const int CV = a.C / V;
  const long long total = (long long)a.B * a.H * a.W * CV;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % CV);
  long long pix = idx / CV;
  const int x = (int)(pix % a.W); pix /= a.W;
  const int y = (int)(pix % a.H);
  const int b = (int)(pix / a.H);
  const int c0 = cv * V;
  float va[V], acc[V];
  ldv<T>(reinterpret_cast<const T*>(a.a) + (((long long)b * a.H + y) * a.W + x) * a.C + c0, va);
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = a.w0 * va[j];
  if (a.mode_b != RS_NONE) {
    float vb[V];
    fetch_rs<T>(reinterpret_cast<const T*>(a.b), a.mode_b, b, y, x, a.H, a.W, a.C, c0, vb);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = acc[j] + a.w1 * vb[j];
  }
  if (a.mode_c != RS_NONE) {
    float vc[V];
    fetch_rs<T>(reinterpret_cast<const T*>(a.c), a.mode_c, b, y, x, a.H, a.W, a.C, c0, vc);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = acc[j] + a.w2 * vc[j];
  }
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = apply_act<T>(acc[j], ACT_SWISH);
  stv<T>(reinterpret_cast<T*>(a.out) + (((long long)b * a.H + y) * a.W + x) * a.C + c0, acc);
}

// torch.cat((feat, estimate), dim=1) for the refinement sub-nets, NHWC, zero padded to Cpad channels
template <typename T>
__global__ void __launch_bounds__(256) concat_kernel(const ConcatProb* __restrict__ probs, int nprobs) {
  constexpr int V = VecN<T>::N;
  pdl_trigger();
  pdl_wait();
  int pi = 0;
  while (pi + 1 < nprobs && (int)blockIdx.x >= probs[pi + 1].blk_start) ++pi;
  const ConcatProb p = probs[pi];
  const int CV = p.Cpad / V;
  const long long item = (long long)((int)blockIdx.x - p.blk_start) * 256 + threadIdx.x;
  const int pix = (int)(item / CV), cv = (int)(item - (long long)pix * CV);
  if (pix >= p.npix) return;
  const int c0 = cv * V;
  float v[V];
  if (c0 < 64) {
    ldv<T>(reinterpret_cast<const T*>(p.feat) + (long long)pix * 64 + c0, v);
  } else {
    const int img = pix / p.HW, px = pix - img * p.HW;
    const float* base = p.head + (long long)img * p.img_stride + (long long)px * p.P;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = c0 + j - 64;
      float x = 0.f;
      if (c < p.P) {
        const int k = p.trans ? (c < 18 ? (c >> 1) * 3 + (c & 1) : (c - 18) * 3 + 2) : c;
        x = base[k];
      }
      v[j] = x;
    }
  }
  stv<T>(reinterpret_cast<T*>(p.out) + (long long)pix * p.Cpad + c0, v);
}

// out(B,H,W,C) = MaxPool2dStaticSamePadding(3,2)(src(B,2H,2W,C))   (p5_to_p6 / p6_to_p7, model.py:119-125)
template <typename T>
__global__ void __launch_bounds__(256) pool_kernel(const T* __restrict__ src, T* __restrict__ out, int B, int H,
                                                   int W, int C) {
  constexpr int V = VecN<T>::N;
  pdl_trigger();
  pdl_wait();
  const int CV = C / V;
  const long long total = (long long)B * H * W * CV;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % CV);
  long long pix = idx / CV;
  const int x = (int)(pix % W); pix /= W;
  const int y = (int)(pix % H);
  const int b = (int)(pix / H);
  float v[V];
  fetch_rs<T>(src, RS_POOL, b, y, x, H, W, C, cv * V, v);
  stv<T>(out + (((long long)b * H + y) * W + x) * C + cv * V, v);
}

// ---------------------------------------------------------------------------------------------
// Epilogue store shared by both GEMM kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void gemm_store(const GemmProb& p, int m, int n, float v) {
  if (p.out_mode == 0) {
    const long long o = (long long)m * p.ldo + n;
    if (p.residual) v += to_f<T>(reinterpret_cast<const T*>(p.residual)[o]);
    reinterpret_cast<T*>(p.out)[o] = from_f<T>(v);
  } else {
    const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
    const int a = n / p.p_src, q = n - a * p.p_src;
    float* dstp = reinterpret_cast<float*>(p.out) + img * p.img_stride + (long long)pix * p.pix_stride + a * p.p_dst + p.p_off + q;
    *dstp = p.accumulate ? *dstp + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// FFMA pointwise GEMM (parity mode; also the cross-check for the tcgen05 kernel).
// 64x64 tile, BK=16, 256 threads, 4x4 register tile, fp32 accumulate in k order.
// ---------------------------------------------------------------------------------------------
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmProb* __restrict__ probs, int nprobs) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  pdl_trigger();
  pdl_wait();
  int pi = 0;
  while (pi + 1 < nprobs && (int)blockIdx.x >= probs[pi + 1].tile_start) ++pi;
  const GemmProb p = probs[pi];
  const int tile = blockIdx.x - p.tile_start;
  const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
  const int m0 = mt * SG_BM, n0 = nt * SG_BN;
  const T* A = reinterpret_cast<const T*>(p.A);
  const T* W = reinterpret_cast<const T*>(p.W);
  const int tid = threadIdx.x;
  const int lr = tid >> 2;         // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;    // 0,4,8,12 : k offset (4 consecutive k)
  const int tm = (tid >> 4) * 4;   // register tile origin
  const int tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int am = m0 + lr, wn = n0 + lr;
  const float* sc = nullptr;
  if (p.a_scale && am < p.M) sc = p.a_scale + (long long)(am / p.rows_per_img) * p.K;
  for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (k < p.K) {  // K is a multiple of 4
      if (am < p.M) {
#pragma unroll
        for (int j = 0; j < 4; ++j) av[j] = to_f<T>(A[(long long)am * p.lda + k + j]);
        if (sc) {
#pragma unroll
          for (int j = 0; j < 4; ++j) av[j] *= sc[k + j];
        }
      }
      if (wn < p.N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = to_f<T>(W[(long long)wn * p.K + k + j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lk + j][lr] = av[j]; Ws[lk + j][lr] = wv[j]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tn]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= p.N) continue;
      gemm_store<T>(p, m, n, apply_act<T>(acc[i][j] + p.bias[n], p.act));
    }
  }
}

}  // namespace hp

// =============================================================================================
// v2 kernels (throughput-oriented).  The v1 kernels above stay as the cross-check (HMDPOSE_V1=1).
// =============================================================================================
namespace hp {

template <typename T> struct RawVec;
template <> struct RawVec<float> { float4 r; };
template <> struct RawVec<__half> { uint4 r; };
template <typename T> __device__ __forceinline__ RawVec<T> ld_raw(const T* p);
template <> __device__ __forceinline__ RawVec<float> ld_raw<float>(const float* p) {
  RawVec<float> v; v.r = __ldg(reinterpret_cast<const float4*>(p)); return v;
}
template <> __device__ __forceinline__ RawVec<__half> ld_raw<__half>(const __half* p) {
  RawVec<__half> v; v.r = __ldg(reinterpret_cast<const uint4*>(p)); return v;
}
__device__ __forceinline__ void unpack(const RawVec<float>& v, float* f) { f[0] = v.r.x; f[1] = v.r.y; f[2] = v.r.z; f[3] = v.r.w; }
__device__ __forceinline__ void unpack(const RawVec<__half>& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v.r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// ---------------------------------------------------------------------------------------------
// Depthwise K x K stride S stencil, register sliding window: one thread = one channel vector x a strip
// of DW2_OWT output pixels of one row.  Per input row the (OWT-1)*S+K input vectors of the strip are
// loaded once (raw, all in flight), converted once and applied to every output / tap that needs them.
// Taps (BN folded) for the block's channel chunk sit in shared memory.  TF-SAME zero padding = bounds
// tests.  Epilogue: folded-BN bias, swish, rounded store, deterministic squeeze partial sums.
// FUSED: the input pixel is the BiFPN node input swish(w0*a + w1*rs(b) + w2*rs(c)) computed on the fly
// (efficientdet/model.py:215-264 fused into the following depthwise conv).
// ---------------------------------------------------------------------------------------------
constexpr int DW2_OWT = 4;

template <typename T, bool FUSED>
__device__ __forceinline__ void dw2_fetch(const DwGroup& g, const T* in, int b, int iy, int ix, int c0, float* v) {
  constexpr int V = VecN<T>::N;
  if (!FUSED) {
    const RawVec<T> r = ld_raw<T>(in + (((long long)b * g.H + iy) * g.W + ix) * g.C + c0);
    unpack(r, v);
  } else {
    float a[V];
    ldv<T>(in + (((long long)b * g.H + iy) * g.W + ix) * g.C + c0, a);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = g.w0 * a[j];
    if (g.mode_b != RS_NONE) {
      float t[V];
      fetch_rs<T>(reinterpret_cast<const T*>(g.fb), g.mode_b, b, iy, ix, g.H, g.W, g.C, c0, t);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = v[j] + g.w1 * t[j];
    }
    if (g.mode_c != RS_NONE) {
      float t[V];
      fetch_rs<T>(reinterpret_cast<const T*>(g.fc), g.mode_c, b, iy, ix, g.H, g.W, g.C, c0, t);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = v[j] + g.w2 * t[j];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = to_f<T>(from_f<T>(apply_act<T>(v[j], ACT_SWISH)));
  }
}

template <typename T, int K, int S, bool FUSED>
__global__ void __launch_bounds__(256, 2) dw2_kernel(const DwGroup* __restrict__ groups, int ngroups) {
  constexpr int V = VecN<T>::N;
  constexpr int OWT = DW2_OWT;
  constexpr int IW = (OWT - 1) * S + K;
  extern __shared__ float dw2_smem[];
  int gi = 0;
  while (gi + 1 < ngroups && (int)blockIdx.x >= groups[gi + 1].block_start) ++gi;
  const DwGroup g = groups[gi];
  const int cvb = g.cvb, cvbV = cvb * V;
  float* wsm = dw2_smem;                 // [K*K][cvb*V]
  float* red = dw2_smem + K * K * cvbV;  // [256][V]
  int blk = blockIdx.x - g.block_start;
  const int chunk = blk % g.cv_chunks; blk /= g.cv_chunks;
  const int txt = blk % g.tiles_x; blk /= g.tiles_x;
  const int ygroups = cdiv(g.tiles_y, g.rep);
  const int tyg = blk % ygroups;
  const int b = blk / ygroups;
  const int tid = threadIdx.x;
  const int tx = tid % cvb;
  const int s = (tid / cvb) % g.sw;
  const int r = tid / (cvb * g.sw);
  const int CV = g.C / V;
  const int cv = chunk * cvb + tx;
  const int c0 = cv * V;
  const bool active = (r < g.sh) && (cv < CV);
  pdl_trigger();
  for (int i = tid; i < K * K * cvbV; i += 256) {   // constants: may run ahead of the producer kernel
    const int tap = i / cvbV, cc = i - tap * cvbV;
    const int c = chunk * cvbV + cc;
    wsm[i] = c < g.C ? __ldg(g.w + tap * g.C + c) : 0.f;
  }
  __syncthreads();
  pdl_wait();
  const T* in = reinterpret_cast<const T*>(g.in);
  T* out = reinterpret_cast<T*>(g.out) + (long long)b * g.Ho * g.Wo * g.C;
  float se[V];
#pragma unroll
  for (int j = 0; j < V; ++j) se[j] = 0.f;
  float bias[V];
#pragma unroll
  for (int j = 0; j < V; ++j) bias[j] = (g.bias && active) ? __ldg(g.bias + c0 + j) : 0.f;
  const float* wt = wsm + tx * V;
  for (int rep = 0; rep < g.rep; ++rep) {
    const int tyt = tyg * g.rep + rep;
    const int oy = tyt * g.sh + r;
    const int ox0 = (txt * g.sw + s) * OWT;
    if (!active || tyt >= g.tiles_y || oy >= g.Ho || ox0 >= g.Wo) continue;
    float acc[OWT][V];
#pragma unroll
    for (int o = 0; o < OWT; ++o)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[o][j] = bias[j];
    const int iy0 = oy * S - g.pad, ix0 = ox0 * S - g.pad;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = iy0 + ky;
      if (iy < 0 || iy >= g.H) continue;
      if (!FUSED) {
        RawVec<T> raw[IW];
        const T* rowp = in + (((long long)b * g.H + iy) * g.W) * g.C + c0;
#pragma unroll
        for (int i = 0; i < IW; ++i) {
          const int ix = ix0 + i;
          if (ix >= 0 && ix < g.W) raw[i] = ld_raw<T>(rowp + (long long)ix * g.C);
        }
#pragma unroll
        for (int i = 0; i < IW; ++i) {
          const int ix = ix0 + i;
          if (ix < 0 || ix >= g.W) continue;
          float v[V];
          unpack(raw[i], v);
#pragma unroll
          for (int o = 0; o < OWT; ++o) {
            const int kx = i - o * S;
            if (kx >= 0 && kx < K) {
              const float* wp = wt + (ky * K + kx) * cvbV;
#pragma unroll
              for (int j = 0; j < V; ++j) acc[o][j] = fmaf(v[j], wp[j], acc[o][j]);
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < IW; ++i) {
          const int ix = ix0 + i;
          if (ix < 0 || ix >= g.W) continue;
          float v[V];
          dw2_fetch<T, true>(g, in, b, iy, ix, c0, v);
#pragma unroll
          for (int o = 0; o < OWT; ++o) {
            const int kx = i - o * S;
            if (kx >= 0 && kx < K) {
              const float* wp = wt + (ky * K + kx) * cvbV;
#pragma unroll
              for (int j = 0; j < V; ++j) acc[o][j] = fmaf(v[j], wp[j], acc[o][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < OWT; ++o) {
      if (ox0 + o < g.Wo) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          acc[o][j] = to_f<T>(from_f<T>(apply_act<T>(acc[o][j], g.act)));
          se[j] += acc[o][j];
        }
        stv<T>(out + ((long long)oy * g.Wo + ox0 + o) * g.C + c0, acc[o]);
      }
    }
  }
  if (g.se_partial) {
#pragma unroll
    for (int j = 0; j < V; ++j) red[tid * V + j] = se[j];
    __syncthreads();
    if (tid < cvb && cv < CV) {
      float sum[V];
#pragma unroll
      for (int j = 0; j < V; ++j) sum[j] = 0.f;
      const int np = g.sw * g.sh;
      for (int q = 0; q < np; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) sum[j] += red[(q * cvb + tx) * V + j];
      float* dst = g.se_partial + ((long long)b * g.tiles_per_img + (tyg * g.tiles_x + txt)) * g.C + c0;
#pragma unroll
      for (int j = 0; j < V; ++j) dst[j] = sum[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Depthwise stencil v3: shared-memory tiled.  A block owns th x tw output pixels x cb channel vectors.
// The input tile (with its (K-1) halo, TF-SAME zero padding produced by cp.async zero-fill) is copied
// global -> shared with 16-byte cp.async: every byte of the tile is in flight at once and no register is
// held across the DRAM/L2 latency (the v2 register-window kernel was latency-bound at 16 warps/SM, ncu in
// profiles/).  Each thread then computes one strip of 4 output pixels x one channel vector from shared memory.
// blockDim = cb * th * tw / 4.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Squeeze-excite gate of one image by ONE thread block of any size (multiple of 32): fixed-order reductions, so the
// result does not depend on which block happens to be last.  scratch: (nthreads*4 + C + 64) floats of shared memory.
template <typename T>
__device__ __forceinline__ void se_tail(const DwGroup& g, int b, float* scratch) {
  const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int C = g.C, C4 = C >> 2, Cse = g.se_cse, tiles = g.tiles_per_img;
  float4* red = reinterpret_cast<float4*>(scratch);            // [nthreads]
  float* pooled = scratch + nthreads * 4;                      // [C]
  float* r = pooled + C;                                       // [64]
  const float* pp = g.se_partial + (long long)b * tiles * C;
  // squeeze: columns in passes of `ncol` float4s; G groups of tiles per column, combined in a fixed order
  for (int col0 = 0; col0 < C4; col0 += nthreads) {
    const int ncol = min(C4 - col0, nthreads);
    const int G = max(1, min(tiles, nthreads / ncol));
    const int i = tid % ncol, gq = tid / ncol;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gq < G) {
#pragma unroll 4
      for (int t = gq; t < tiles; t += G) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(pp + (long long)t * C) + col0 + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    red[tid] = s;
    __syncthreads();
    if (tid < ncol) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < G; ++q) {
        const float4 v = red[q * ncol + tid];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      reinterpret_cast<float4*>(pooled)[col0 + tid] =
          make_float4(a.x * g.se_inv_hw, a.y * g.se_inv_hw, a.z * g.se_inv_hw, a.w * g.se_inv_hw);
    }
    __syncthreads();
  }
  // FC1 + swish: one (full) warp per squeezed channel
  if (nwarps == 0) {   // block smaller than a warp: every thread owns whole rows
    for (int j = tid; j < Cse; j += nthreads) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(__ldg(g.se_wr + (long long)j * C + c), pooled[c], s);
      r[j] = apply_act<T>(s + __ldg(g.se_br + j), ACT_SWISH);
    }
  } else if (warp < nwarps)
  for (int j = warp; j < Cse; j += nwarps) {
    const float4* wrow = reinterpret_cast<const float4*>(g.se_wr + (long long)j * C);
    float s = 0.f;
#pragma unroll 4
    for (int c4 = lane; c4 < C4; c4 += 32) {
      const float4 w = __ldg(wrow + c4);
      const float4 p = reinterpret_cast<const float4*>(pooled)[c4];
      s = fmaf(w.x, p.x, fmaf(w.y, p.y, fmaf(w.z, p.z, fmaf(w.w, p.w, s))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) r[j] = apply_act<T>(s + __ldg(g.se_br + j), ACT_SWISH);
  }
  __syncthreads();
  // FC2 + sigmoid: one thread per float4 column, squeezed channels in order
  for (int c4 = tid; c4 < C4; c4 += nthreads) {
    float4 a = __ldg(reinterpret_cast<const float4*>(g.se_be) + c4);
#pragma unroll 4
    for (int j = 0; j < Cse; ++j) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(g.se_weT + (long long)j * C) + c4);
      const float rj = r[j];
      a.x = fmaf(w.x, rj, a.x); a.y = fmaf(w.y, rj, a.y); a.z = fmaf(w.z, rj, a.z); a.w = fmaf(w.w, rj, a.w);
    }
    reinterpret_cast<float4*>(g.se_gate + (long long)b * C)[c4] =
        make_float4(sigmoid_t<T>(a.x), sigmoid_t<T>(a.y), sigmoid_t<T>(a.z), sigmoid_t<T>(a.w));
  }
}

template <typename T, int K, int S, int CB>
__global__ void __launch_bounds__(256, 3) dw3_kernel(const __grid_constant__ DwGroup g) {
  constexpr int V = VecN<T>::N;
  constexpr int OWT = 4;
  constexpr int IW = (OWT - 1) * S + K;
  extern __shared__ __align__(128) uint8_t dw3_smem[];
  __shared__ uint64_t tma_bar;
  // (the group is a kernel PARAMETER: a block's first instruction does not wait for a global load of its own description)
  constexpr int cb = CB;   // compile-time: shared-memory offsets of the stencil become immediates
  const int th = g.th, tw = g.tw;
  const int ih = (th - 1) * S + K, iwd = (tw - 1) * S + K;
  const int nthreads = blockDim.x;
  uint8_t* tile = dw3_smem;                                            // [ih][iwd][cb] 16-byte vectors
  float* wsm = reinterpret_cast<float*>(dw3_smem + (size_t)ih * iwd * cb * 16);   // [K*K][cb*V]
  float* red = wsm + K * K * cb * V;                                   // [nthreads][V]
  const int tiles_x = cdiv(g.Wo, tw), tiles_y = cdiv(g.Ho, th);
  int blk = blockIdx.x;
  const int chunk = blk % g.cv_chunks; blk /= g.cv_chunks;
  const int txt = blk % tiles_x; blk /= tiles_x;
  const int tyt = blk % tiles_y;
  const int b = blk / tiles_y;
  const int tid = threadIdx.x;
  const int CV = g.C / V;
  const int cbV = cb * V;
  pdl_trigger();
  if (tid == 0) {
    mbar_init(&tma_bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < K * K * cbV; i += nthreads) {   // taps are constants: load ahead of the producer kernel
    const int tap = i / cbV, cc = i - tap * cbV;
    const int c = chunk * cbV + cc;
    wsm[i] = c < g.C ? __ldg(g.w + tap * g.C + c) : 0.f;
  }
  __syncthreads();
  pdl_wait();
  // ---- input tile + halo -> shared memory: ONE 4-D TMA per block; out-of-bounds = TF-SAME zero padding ----
  const int oy0 = tyt * th, ox0 = txt * tw;
  const int iy0 = oy0 * S - g.pad, ix0 = ox0 * S - g.pad;
  if (tid == 0) {
    mbar_expect_tx(&tma_bar, (uint32_t)(ih * iwd * cb * 16));
    tma_load_4d(tile, g.tmap, &tma_bar, chunk * cbV, ix0, iy0, b);
  }
  // ---- compute one strip per thread ----
  const int cv = tid % cb;
  const int strip = tid / cb;
  const int spr = tw / OWT;
  const int ry = strip / spr, sx = strip - ry * spr;
  const int oy = oy0 + ry, ox = ox0 + sx * OWT;
  const int gcv = chunk * cb + cv;
  const int c0 = gcv * V;
  const bool active = gcv < CV && oy < g.Ho && ox < g.Wo;
  float se[V], bias_r[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    se[j] = 0.f;
    bias_r[j] = (active && g.bias) ? __ldg(g.bias + c0 + j) : 0.f;   // constant: in flight while the tile lands
  }
  mbar_wait(&tma_bar, 0);
  if (active) {
    float acc[OWT][V];
#pragma unroll
    for (int o = 0; o < OWT; ++o)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[o][j] = bias_r[j];
    const float* wt = wsm + cv * V;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const uint8_t* rowp = tile + ((size_t)((ry * S + ky) * iwd + sx * OWT * S) * cb + cv) * 16;
#pragma unroll
      for (int i = 0; i < IW; ++i) {
        RawVec<T> raw;
        raw.r = *reinterpret_cast<const decltype(raw.r)*>(rowp + (size_t)i * cb * 16);
        float v[V];
        unpack(raw, v);
#pragma unroll
        for (int o = 0; o < OWT; ++o) {
          const int kx = i - o * S;
          if (kx >= 0 && kx < K) {
            const float* wp = wt + (ky * K + kx) * cbV;
#pragma unroll
            for (int j = 0; j < V; ++j) acc[o][j] = fmaf(v[j], wp[j], acc[o][j]);
          }
        }
      }
    }
    T* out = reinterpret_cast<T*>(g.out) + (long long)b * g.Ho * g.Wo * g.C;
#pragma unroll
    for (int o = 0; o < OWT; ++o) {
      if (ox + o < g.Wo) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          acc[o][j] = to_f<T>(from_f<T>(apply_act<T>(acc[o][j], g.act)));
          se[j] += acc[o][j];
        }
        stv<T>(out + ((long long)oy * g.Wo + ox + o) * g.C + c0, acc[o]);
      }
    }
  }
  if (g.se_partial) {
#pragma unroll
    for (int j = 0; j < V; ++j) red[tid * V + j] = se[j];
    __syncthreads();
    if (tid < cb && gcv < CV) {
      float sum[V];
#pragma unroll
      for (int j = 0; j < V; ++j) sum[j] = 0.f;
      const int ns = nthreads / cb;
      for (int q = 0; q < ns; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) sum[j] += red[(q * cb + cv) * V + j];
      float* dst = g.se_partial + ((long long)b * g.tiles_per_img + (tyt * tiles_x + txt)) * g.C + c0;
#pragma unroll
      for (int j = 0; j < V; ++j) dst[j] = sum[j];
      if (g.se_counter) __threadfence();   // publish the sums before this block is counted
    }
    if (g.se_counter) {
      __shared__ int s_last;
      __syncthreads();
      if (tid == 0) {
        const int blocks_per_img = tiles_x * tiles_y * g.cv_chunks;
        const int prev = atomicAdd(g.se_counter + b, 1);
        s_last = prev == blocks_per_img - 1;
        if (s_last) g.se_counter[b] = 0;   // ready for the next launch (nobody else touches it any more)
        __threadfence();
      }
      __syncthreads();
      if (s_last) se_tail<T>(g, b, reinterpret_cast<float*>(dw3_smem));   // tile / taps / red are dead by now
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Squeeze-excite gate v2 (efficientnet/model.py:88-93): one block per image; the partial sums are
// reduced by all 256 threads in two fixed-order levels; se_expand weights are stored transposed
// ([Cse][C]) so the second FC reads coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int SE2_THREADS = 320;  // 10 warps: one FC2 pass covers C/4 <= 320 float4 columns (C <= 1280)

template <typename T>
__global__ void __launch_bounds__(SE2_THREADS) se2_kernel(const float* __restrict__ partial, int tiles, int C, int Cse,
                                                          float inv_hw, const float* __restrict__ wr,
                                                          const float* __restrict__ br, const float* __restrict__ weT,
                                                          const float* __restrict__ be, float* __restrict__ gate) {
  __shared__ __align__(16) float part[8 * 1152];
  __shared__ __align__(16) float pooled[1152];
  __shared__ float r[64];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pp = partial + (long long)b * tiles * C;
  const int C4 = C >> 2;  // C is a multiple of 8
  // level 1: up to 8 interleaved groups of tiles, float4 per thread, all loads independent
  const int G = tiles < 8 ? tiles : 8;
  for (int idx = tid; idx < C4 * G; idx += SE2_THREADS) {
    const int gidx = idx / C4, c4 = idx - gidx * C4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int t = gidx; t < tiles; t += G) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pp + (long long)t * C) + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(part + gidx * C)[c4] = s;
  }
  __syncthreads();
  for (int c = tid; c < C; c += SE2_THREADS) {
    float s = 0.f;
    for (int q = 0; q < G; ++q) s += part[q * C + c];
    pooled[c] = s * inv_hw;
  }
  __syncthreads();
  // FC1: warp w owns squeezed channels w, w+10, ...; every lane keeps all its rows' partial sums so that the
  // weight loads of all rows are independent and in flight together; one shuffle reduction per row at the end
  constexpr int NW = SE2_THREADS / 32, MAXR = 5;  // Cse <= 48 -> at most 5 rows per warp
  const int warp = tid >> 5, lane = tid & 31;
  float sums[MAXR];
#pragma unroll
  for (int q = 0; q < MAXR; ++q) sums[q] = 0.f;
  for (int c4 = lane; c4 < C4; c4 += 32) {
    const float4 p = reinterpret_cast<const float4*>(pooled)[c4];
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
      const int j = warp + NW * q;
      if (j < Cse) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + (long long)j * C) + c4);
        sums[q] = fmaf(w.x, p.x, fmaf(w.y, p.y, fmaf(w.z, p.z, fmaf(w.w, p.w, sums[q]))));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MAXR; ++q) {
    float s = sums[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int j = warp + NW * q;
    if (lane == 0 && j < Cse) r[j] = apply_act<T>(s + br[j], ACT_SWISH);
  }
  __syncthreads();
  // FC2: four consecutive channels per thread, transposed weights [Cse][C] -> coalesced float4 rows
  for (int c4 = tid; c4 < C4; c4 += SE2_THREADS) {
    float4 s = __ldg(reinterpret_cast<const float4*>(be) + c4);
#pragma unroll 8
    for (int j = 0; j < Cse; ++j) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(weT + (long long)j * C) + c4);
      const float rj = r[j];
      s.x = fmaf(w.x, rj, s.x); s.y = fmaf(w.y, rj, s.y); s.z = fmaf(w.z, rj, s.z); s.w = fmaf(w.w, rj, s.w);
    }
    float4 g;
    g.x = sigmoid_t<T>(s.x); g.y = sigmoid_t<T>(s.y); g.z = sigmoid_t<T>(s.z); g.w = sigmoid_t<T>(s.w);
    reinterpret_cast<float4*>(gate + (long long)b * C)[c4] = g;
  }
}

// ---------------------------------------------------------------------------------------------
// Squeeze-excite gate v3: one thread-block CLUSTER of 8 CTAs per image.  Each phase (partial-sum reduction,
// FC1, FC2) is split 8 ways and its result is broadcast to the whole cluster through distributed shared
// memory, so a CTA pulls 1/8 of the (up to 442 KB) FC weights from L2 instead of all of them -- the single-CTA
// version was bound by one SM's L2 bandwidth and by three serial L2 round trips.
// ---------------------------------------------------------------------------------------------
constexpr int SE3_CL = 8;
constexpr int SE3_THREADS = 256;

template <typename T>
__global__ void __cluster_dims__(SE3_CL, 1, 1) __launch_bounds__(SE3_THREADS)
se3_kernel(const float* __restrict__ partial, int tiles, int C, int Cse, float inv_hw, const float* __restrict__ wr,
           const float* __restrict__ br, const float* __restrict__ weT, const float* __restrict__ be,
           float* __restrict__ gate, T* __restrict__ x, int HW, const T* __restrict__ wproj, T* __restrict__ wgated,
           int wN) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ __align__(16) float pooled[1152];
  __shared__ __align__(16) float red[SE3_THREADS * 4];
  __shared__ __align__(16) float gate_s[1152];
  __shared__ float r[64];
  pdl_trigger();
  // every CTA of the cluster must have STARTED before a peer writes its shared memory (compute-sanitizer racecheck:
  // "located in a block that might not have entered yet"): arrive here, wait right before the first remote store
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / SE3_CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C4 = C >> 2;
  const int ncol = (C4 - rank + SE3_CL - 1) / SE3_CL;   // float4 columns c4 = rank + 8*i owned by this CTA (<= 36)
  const float* pp = partial + (long long)b * tiles * C;
  // The FC weights are constants: every weight this thread will need is fetched into registers BEFORE the wait for the
  // previous kernel, so that the three phases below are separated by barriers only, not by L2 round trips.
  constexpr int W1N = 9, W2N = 8;      // C <= 1152: 288 float4 per FC1 row = 9 per lane; FC2: <= ceil(48 / 7) rows per thread
  const int j1 = rank + SE3_CL * warp;   // FC1 row of this warp
  float4 w1r[W1N];
#pragma unroll
  for (int k = 0; k < W1N; ++k) {
    const int c4 = lane + 32 * k;
    w1r[k] = (j1 < Cse && c4 < C4) ? __ldg(reinterpret_cast<const float4*>(wr + (long long)j1 * C) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float br1 = j1 < Cse ? __ldg(br + j1) : 0.f;
  const int P2 = SE3_THREADS / max(ncol, 1);
  const int i2 = tid % max(ncol, 1), p2 = tid / max(ncol, 1);
  const bool on2 = ncol > 0 && p2 < P2;
  float4 w2r[W2N];
#pragma unroll
  for (int k = 0; k < W2N; ++k) {
    const int j = p2 + k * P2;
    w2r[k] = (on2 && j < Cse) ? __ldg(reinterpret_cast<const float4*>(weT + (long long)j * C) + rank + SE3_CL * i2) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 be4 = tid < ncol ? __ldg(reinterpret_cast<const float4*>(be) + rank + SE3_CL * tid) : make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_wait();

  // phase 1: squeeze.  thread = (column i, tile group g); groups are summed in a fixed order
  {
    const int G = max(1, min(tiles, SE3_THREADS / max(ncol, 1)));
    const int i = tid % max(ncol, 1), g = tid / max(ncol, 1);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ncol > 0 && g < G) {
      const int c4 = rank + SE3_CL * i;
#pragma unroll 4
      for (int t = g; t < tiles; t += G) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pp + (long long)t * C) + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    reinterpret_cast<float4*>(red)[tid] = s;
    __syncthreads();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // all peers are running (see top)
    if (tid < ncol) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < G; ++q) {
        const float4 v = reinterpret_cast<const float4*>(red)[q * ncol + tid];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      a.x *= inv_hw; a.y *= inv_hw; a.z *= inv_hw; a.w *= inv_hw;
      const int c4 = rank + SE3_CL * tid;
#pragma unroll
      for (int d = 0; d < SE3_CL; ++d) reinterpret_cast<float4*>(cluster.map_shared_rank(pooled, d))[c4] = a;
    }
  }
  cluster.sync();
  // phase 2: FC1 rows j = rank, rank + 8, ... (<= 6 rows): one warp per row
  {
    const int j = j1;
    if (j < Cse) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < W1N; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < C4) {
          const float4 w = w1r[k];
          const float4 p = reinterpret_cast<const float4*>(pooled)[c4];
          s = fmaf(w.x, p.x, fmaf(w.y, p.y, fmaf(w.z, p.z, fmaf(w.w, p.w, s))));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane < SE3_CL) cluster.map_shared_rank(r, lane)[j] = apply_act<T>(s + br1, ACT_SWISH);
    }
  }
  cluster.sync();
  // phase 3: FC2 for this CTA's columns.  thread = (column i, part p of the squeezed channels)
  if (ncol > 0) {
    const int P = P2;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on2) {
#pragma unroll
      for (int k = 0; k < W2N; ++k) {
        const int j = p2 + k * P;
        if (j < Cse) {
          const float4 w = w2r[k];
          const float rj = r[j];
          s.x = fmaf(w.x, rj, s.x); s.y = fmaf(w.y, rj, s.y); s.z = fmaf(w.z, rj, s.z); s.w = fmaf(w.w, rj, s.w);
        }
      }
    }
    reinterpret_cast<float4*>(red)[tid] = s;
    __syncthreads();
    if (tid < ncol) {
      const int c4 = rank + SE3_CL * tid;
      float4 a = be4;
      for (int q = 0; q < P; ++q) {
        const float4 v = reinterpret_cast<const float4*>(red)[q * ncol + tid];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      float4 g;
      g.x = sigmoid_t<T>(a.x); g.y = sigmoid_t<T>(a.y); g.z = sigmoid_t<T>(a.z); g.w = sigmoid_t<T>(a.w);
      reinterpret_cast<float4*>(gate + (long long)b * C)[c4] = g;
      if (wgated != nullptr) reinterpret_cast<float4*>(gate_s)[tid] = g;   // own columns, local: see the weight scaling below
      if (x != nullptr) {
#pragma unroll
        for (int d = 0; d < SE3_CL; ++d) reinterpret_cast<float4*>(cluster.map_shared_rank(gate_s, d))[c4] = g;
      }
    }
  }
  // Gate folded into the project weights: W'[b][n][k] = W[n][k] * gate[b][k] for this CTA's columns k (one [N x K] panel
  // per image), so that the project convolution of the large maps is an UN-gated TMA -> tcgen05 GEMM
  // (`sigmoid(x_squeezed) * x` followed by the 1x1 conv, efficientnet/model.py:93-97, is linear in x)
  if (wgated != nullptr && ncol > 0) {
    __syncthreads();
    T* wb = wgated + (long long)b * wN * C;
    for (int it = tid; it < wN * ncol; it += SE3_THREADS) {
      const int n = it / ncol, i = it - n * ncol;
      const int c = (rank + SE3_CL * i) * 4;
      const float4 g = reinterpret_cast<const float4*>(gate_s)[i];
      // four consecutive weights = one 8-byte (fp16) / 16-byte (fp32) vector
      struct alignas(4 * sizeof(T)) Vec4 { T e[4]; };
      Vec4 w4 = *reinterpret_cast<const Vec4*>(wproj + (long long)n * C + c);
      w4.e[0] = from_f<T>(to_f<T>(w4.e[0]) * g.x); w4.e[1] = from_f<T>(to_f<T>(w4.e[1]) * g.y);
      w4.e[2] = from_f<T>(to_f<T>(w4.e[2]) * g.z); w4.e[3] = from_f<T>(to_f<T>(w4.e[3]) * g.w);
      *reinterpret_cast<Vec4*>(wb + (long long)n * C + c) = w4;
    }
  }
  // Small feature maps (H*W <= 256: blocks 5..15): apply the gate here, `sigmoid(x_squeezed) * x`
  // (efficientnet/model.py:93), in place on this CTA's 1/8 of the image, so that the deep-K project GEMM runs as
  // a pure TMA -> tcgen05 pipeline.  Large maps keep the gate fused into the GEMM's A-tile (1-3 k-blocks).
  if (x != nullptr) {
    cluster.sync();
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int p0 = (HW * rank) / SE3_CL, p1 = (HW * (rank + 1)) / SE3_CL;
    T* xb = x + (long long)b * HW * C;
    for (int i = p0 * CV + tid; i < p1 * CV; i += SE3_THREADS) {
      const int cv = i % CV;
      float v[V];
      ldv<T>(xb + (long long)i * V, v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] *= gate_s[cv * V + j];
      stv<T>(xb + (long long)i * V, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gather-before-head for the hand sub-network: FilterDetections keeps at most
// max_detections rows of the (B, N, 63) hand tensor (hmdegopose/layers.py:374), so on the detection path the
// 64 -> 567 header (hmdegopose/model.py:113,146-151: dw3x3 -> pw(+bias), 62 % of all header work and 3.1 MB
// of fp32 output per frame) is evaluated only at the kept anchors.  One warp per detection slot:
// anchor -> (level, pixel, a); depthwise 3x3 at that pixel from the hand trunk output; rows a*63..a*63+62 of
// the pointwise matrix.  Empty slots are padded with -1 (layers.py:384).
// ---------------------------------------------------------------------------------------------
struct HandGatherArgs {
  const void* trunk[5];  // hand trunk output per level, [B,side,side,64]
  int side[5];
  int lvl_off[6];        // first anchor row of each level (+ total)
  const float* dw_w;     // [9][64]
  const void* pw_w;      // [567][64], storage type T
  const float* bias;     // [567]
  const int* det_idx;    // [B][D] kept anchor rows, -1 = empty
  float* det_hand;       // [B][D][63]
  int B, D;
};

template <typename T>
__global__ void __launch_bounds__(256) hand_gather_kernel(HandGatherArgs a) {
  __shared__ float dwv[8][64];
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int det = blockIdx.x * 8 + warp;
  if (det >= a.B * a.D) return;
  const int b = det / a.D;
  const int anchor = a.det_idx[det];
  float* out = a.det_hand + (long long)det * 63;
  if (anchor < 0) {
    out[lane] = -1.0f;
    if (lane + 32 < 63) out[lane + 32] = -1.0f;
    return;
  }
  int l = 0;
  while (l < 4 && anchor >= a.lvl_off[l + 1]) ++l;
  const int rel = anchor - a.lvl_off[l];
  const int pix = rel / 9, aa = rel - pix * 9;
  const int side = a.side[l];
  const int y = pix / side, x = pix - y * side;
  const T* f = reinterpret_cast<const T*>(a.trunk[l]) + (long long)b * side * side * 64;
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + dy - 1, xx = x + dx - 1;
      if (yy < 0 || yy >= side || xx < 0 || xx >= side) continue;
      const T* px = f + ((long long)yy * side + xx) * 64 + 2 * lane;
      const float* w = a.dw_w + (dy * 3 + dx) * 64 + 2 * lane;
      acc0 = fmaf(to_f<T>(px[0]), w[0], acc0);
      acc1 = fmaf(to_f<T>(px[1]), w[1], acc1);
    }
  }
  dwv[warp][2 * lane] = to_f<T>(from_f<T>(acc0));       // the fused kernel feeds the GEMM in the storage type
  dwv[warp][2 * lane + 1] = to_f<T>(from_f<T>(acc1));
  __syncwarp();
  const T* W = reinterpret_cast<const T*>(a.pw_w);
  constexpr int V = VecN<T>::N;
  for (int p = lane; p < 63; p += 32) {
    const int row = aa * 63 + p;
    const T* wr = W + (long long)row * 64;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 64; k += V) {
      float wv[V];
      ldv<T>(wr + k, wv);
#pragma unroll
      for (int j = 0; j < V; ++j) s = fmaf(dwv[warp][k + j], wv[j], s);
    }
    out[p] = s + a.bias[row];
  }
}

// ---------------------------------------------------------------------------------------------
// Gather-before-head for ALL pose headers: FilterDetections keeps at most max_detections rows of the
// rotation / translation / hand tensors (hmdegopose/layers.py:369-374), so on the detection path the three headers
// (hmdegopose/model.py:55-90, 127-156, 191-228: dw3x3 -> pw(+bias)) are evaluated only at the kept anchors, followed
// by the translation recovery of layers.py:142-249 in the reference's op order (explicit IEEE mul / add / div: this
// translation unit is compiled with FMA contraction).  One warp per detection slot; empty slots are padded with -1.
// ---------------------------------------------------------------------------------------------
struct PoseGatherArgs {
  const void* trunk[3][5];   // rotation / translation / hand trunk outputs per level, [B,side,side,64]
  int side[5];
  int lvl_off[6];
  const float* dw_w[4];      // [9][64]: rotation, translation xy, translation z, hand
  const void* pw_w[4];       // [27][64], [18][64], [9][64], [567][64], storage type T
  const float* bias[4];
  const float* tanchors;     // (N,3) cx, cy, stride
  const float* cam;          // [B][6]
  const int* det_idx;        // [B][D] kept anchor rows, -1 = empty
  float* det_rot; float* det_trans; float* det_hand;   // [B][D][3], [B][D][3], [B][D][63]
  int B, D;
};

template <typename T>
__global__ void __launch_bounds__(256) pose_gather_kernel(PoseGatherArgs a) {
  __shared__ float dwv[8][4][64];
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int det = blockIdx.x * 8 + warp;
  if (det >= a.B * a.D) return;
  const int b = det / a.D;
  const int anchor = a.det_idx[det];
  float* o_hand = a.det_hand + (long long)det * 63;
  if (anchor < 0) {
    o_hand[lane] = -1.0f;
    if (lane + 32 < 63) o_hand[lane + 32] = -1.0f;
    if (lane < 3) { a.det_rot[(long long)det * 3 + lane] = -1.0f; a.det_trans[(long long)det * 3 + lane] = -1.0f; }
    return;
  }
  int l = 0;
  while (l < 4 && anchor >= a.lvl_off[l + 1]) ++l;
  const int rel = anchor - a.lvl_off[l];
  const int pix = rel / 9, aa = rel - pix * 9;
  const int side = a.side[l];
  const int y = pix / side, x = pix - y * side;
  // depthwise 3x3 of the four headers at this pixel (translation xy / z share the translation trunk)
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int t = h == 0 ? 0 : (h == 3 ? 2 : 1);
    const T* f = reinterpret_cast<const T*>(a.trunk[t][l]) + (long long)b * side * side * 64;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        if (yy < 0 || yy >= side || xx < 0 || xx >= side) continue;
        const T* px = f + ((long long)yy * side + xx) * 64 + 2 * lane;
        const float* w = a.dw_w[h] + (dy * 3 + dx) * 64 + 2 * lane;
        acc0 = fmaf(to_f<T>(px[0]), w[0], acc0);
        acc1 = fmaf(to_f<T>(px[1]), w[1], acc1);
      }
    }
    dwv[warp][h][2 * lane] = to_f<T>(from_f<T>(acc0));       // the dense kernels feed the GEMM in the storage type
    dwv[warp][h][2 * lane + 1] = to_f<T>(from_f<T>(acc1));
  }
  __syncwarp();
  constexpr int V = VecN<T>::N;
  auto dot = [&](int h, int row) {
    const T* wr = reinterpret_cast<const T*>(a.pw_w[h]) + (long long)row * 64;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 64; k += V) {
      float wv[V];
      ldv<T>(wr + k, wv);
#pragma unroll
      for (int j = 0; j < V; ++j) s = fmaf(dwv[warp][h][k + j], wv[j], s);
    }
    return s + a.bias[h][row];
  };
  // lanes 0-2: rotation (rows aa*3 + q), lanes 3-4: translation xy (rows aa*2 + q), lane 5: translation z (row aa)
  float v = 0.f;
  if (lane < 3) v = dot(0, aa * 3 + lane);
  else if (lane < 5) v = dot(1, aa * 2 + (lane - 3));
  else if (lane == 5) v = dot(2, aa);
  if (lane < 3) a.det_rot[(long long)det * 3 + lane] = v;
  const float dx_ = __shfl_sync(0xffffffffu, v, 3), dy_ = __shfl_sync(0xffffffffu, v, 4), dz_ = __shfl_sync(0xffffffffu, v, 5);
  if (lane == 0) {
    // translation_transform_inv (layers.py:142-166) + CalculateTxTy (layers.py:212-249), op order as written
    const float* ta = a.tanchors + 3 * anchor;
    const float* cam = a.cam + 6 * b;
    const float stride = ta[2];
    float tx = __fadd_rn(ta[0], __fmul_rn(dx_, stride));
    float ty = __fadd_rn(ta[1], __fmul_rn(dy_, stride));
    const float fx = cam[0], fy = cam[1], px = cam[2], py = cam[3], tzs = cam[4], ims = cam[5];
    tx = __fdiv_rn(tx, ims);
    ty = __fdiv_rn(ty, ims);
    const float tz = __fmul_rn(dz_, tzs);
    tx = __fsub_rn(tx, px);
    ty = __fsub_rn(ty, py);
    float* o = a.det_trans + (long long)det * 3;
    o[0] = __fdiv_rn(__fmul_rn(tx, tz), fx);
    o[1] = __fdiv_rn(__fmul_rn(ty, tz), fy);
    o[2] = tz;
  }
  for (int p = lane; p < 63; p += 32) o_hand[p] = dot(3, aa * 63 + p);
}

}  // namespace hp
