// One MBConv block of the small feature maps (<= 256 pixels per image) as ONE kernel (fast mode, fp16):
//
//   expand 1x1 + BN + swish -> depthwise kxk (stride 1/2, TF-SAME zero padding) + BN + swish -> squeeze-excite
//   (avg-pool, 1x1 + swish, 1x1 + sigmoid, gate) -> project 1x1 + BN (+ identity skip)      efficientnet/model.py:69-104
//
// replaces four launches (expand GEMM, dw3_kernel, se3_kernel, gated project GEMM) and three HBM/L2 round trips of the
// 6x-expanded tensor, which now never leaves the SM.
//
// Decomposition: a thread-block CLUSTER of 8 CTAs owns one image and splits the EXPANDED CHANNELS into slices of 64
// (one 128-byte swizzle row of fp16 per pixel); CTA `rank` takes slices rank, rank + 8, rank + 16.  The depthwise
// conv and both BN/swish are per channel, so a channel slice needs no halo and no exchange:
//   1. x (all pixels x cin) -> smem as a K-major SWIZZLE_128B UMMA operand (cp.async, 16-byte chunks)
//   2. per slice: tcgen05.mma  D1[pixels x 64] = x . W_exp[slice]^T  (accumulator in TMEM)
//      epilogue: tcgen05.ld -> + bias -> swish -> fp16 -> smem tile [pixel][64] (swizzled: conflict-free for the stencil)
//      depthwise stencil from smem, 4x8 / 2x8 register strips -> + bias -> swish -> fp16 straight into the A operand
//      of the project GEMM (same swizzled layout), channel sums for the squeeze on the way
//   3. squeeze-excite: FC1 partial sums over the CTA's channels -> all-reduce over the cluster through distributed
//      shared memory -> FC2 for the CTA's channels -> gate applied to the A operand in place
//   4. project GEMM split-K over the cluster: tcgen05.mma  D2[pixels x cout] += A2[slice] . W_proj[:, slice]^T
//   5. reduce-scatter of the fp32 partial tiles through distributed shared memory (each CTA owns pixels/8 rows and
//      receives 8 partials, conflict-free float4 remote stores), + bias (+ skip) -> fp16 -> global, coalesced.
// Weights are constants: they are prefetched before griddepcontrol.wait (programmatic dependent launch), so the
// prologue overlaps the tail of the previous block's kernel.
#pragma once
#include <cooperative_groups.h>

#include "gemm_tc.cuh"

namespace hp {

__device__ __forceinline__ void mb_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mb_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mb_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// debug timeline of CTA 0, thread 0 of the last launch (hmdpose_debug_read("__mb_timeline"), microseconds since entry)
__device__ unsigned long long g_mb_ts[32];
__device__ __forceinline__ void mb_stamp(int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mb_ts[i] = t;
  }
}

// byte offset of 16-byte chunk j of row r inside a K-major SWIZZLE_128B tile (rows of 128 bytes)
__device__ __forceinline__ uint32_t sw128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

template <int K, int S, int SP>
__global__ void __launch_bounds__(MB_THREADS, 1)
mbconv_fused_kernel(const MbSpec sp) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sX = smem;
  uint8_t* sW1 = smem + sp.off_w1;
  uint8_t* sExp = smem + sp.off_exp;
  uint8_t* sA2 = smem + sp.off_a2;
  uint8_t* sW2 = smem + sp.off_w2;
  float* sDw = reinterpret_cast<float*>(smem + sp.off_dw);            // [K*K][64] taps of the current slice
  float* pooled = reinterpret_cast<float*>(smem + sp.off_misc);       // [MB_MAX_MINE][64] squeezed means of this CTA's channels
  float* sGate = pooled + MB_MAX_MINE * 64;                           // [MB_MAX_MINE][64]
  float* s_part = sGate + MB_MAX_MINE * 64;                           // [8 warps][64]
  float* r_all = s_part + 8 * 64;                                     // [cl][64]  (written by every CTA of the cluster)
  float* r_s = r_all + MB_CL_MAX * 64;                                    // [64]
  float* recv = reinterpret_cast<float*>(smem);                       // aliases the operands once the MMAs are done

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  mb_stamp(0);
  const int rank = (int)cluster.block_rank();
  const int CL = sp.cl;
  const int img = blockIdx.x / CL;
  const int cin = sp.cin, cexp = sp.cexp, cout = sp.cout, P = sp.P, Po = sp.Po;
  const int nmine = rank < sp.nsl ? (sp.nsl - rank + CL - 1) / CL : 0;

  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();
  mb_stamp(1);
  cluster_arrive();   // matched by cluster_wait() before the first remote shared-memory access: every CTA has started
  uint32_t mma_phase = 0;

  // ---- constant prefetch (before griddepcontrol.wait): every W_proj slice of this CTA, W_exp + taps of slice 0 ----
  auto load_w1 = [&](int s) {
    const int wd = min(MB_SLICE, cexp - s * MB_SLICE);
    const int cpr = cin >> 3;
    for (int idx = tid; idx < wd * cpr; idx += MB_THREADS) {
      const int n = idx / cpr, cj = idx - n * cpr;
      mb_cp16(sW1 + (cj >> 3) * 8192 + sw128(n, cj & 7), sp.w_exp + (size_t)(s * MB_SLICE + n) * cin + cj * 8);
    }
  };
  auto load_dw = [&](int s, int buf) {
    const int wd = min(MB_SLICE, cexp - s * MB_SLICE);
    const int cpr = wd >> 2;
    for (int idx = tid; idx < K * K * cpr; idx += MB_THREADS) {
      const int tap = idx / cpr, c4 = idx - tap * cpr;
      mb_cp16(sDw + (buf * K * K + tap) * 64 + c4 * 4, sp.w_dw + (size_t)tap * cexp + s * MB_SLICE + c4 * 4);
    }
  };
  for (int li = 0; li < nmine; ++li) {
    const int s = rank + li * CL;
    const int wd = min(MB_SLICE, cexp - s * MB_SLICE);
    const int cpr = wd >> 3;
    for (int idx = tid; idx < cout * cpr; idx += MB_THREADS) {
      const int n = idx / cpr, j = idx - n * cpr;
      mb_cp16(sW2 + li * sp.w2_slice_bytes + sw128(n, j), sp.w_proj + (size_t)n * cexp + s * MB_SLICE + j * 8);
    }
  }
  if (nmine > 0) { load_w1(rank); load_dw(rank, 0); }
  mb_cp_commit();

  // ---- x: written by the previous kernel ----
  mb_stamp(2);
  pdl_wait();
  mb_stamp(3);
  if (nmine > 0) {
    const int cpr = cin >> 3;
    const __half* xi = sp.x + (size_t)img * P * cin;
    for (int idx = tid; idx < P * cpr; idx += MB_THREADS) {
      const int p = idx / cpr, cj = idx - p * cpr;
      mb_cp16(sX + ((cj >> 3) * sp.MT + (p >> 7)) * sp.pitch + sw128(p & 127, cj & 7), xi + (size_t)p * cin + cj * 8);
    }
  }
  mb_cp_commit();

  const int dj = lane & 7;            // depthwise: 16-byte channel chunk of this thread
  const int strip = lane >> 3;        // depthwise: strip of SP output pixels in the row (4 strips per row)
  const int q = warp & 3, g = warp >> 2;

  for (int li = 0; li < nmine; ++li) {
    const int s = rank + li * CL;
    const int c0 = s * MB_SLICE;
    const int wd = min(MB_SLICE, cexp - c0);
    mb_cp_wait_all();
    fence_async_smem();
    __syncthreads();
    mb_stamp(4 + 4 * li);
    // ---- expand GEMM of this slice ----
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(128, wd, 0);
      for (int mt = 0; mt < sp.MT; ++mt)
        for (int kb = 0; kb < sp.KB1; ++kb) {
          const int ksteps = min(4, (cin - kb * 64) >> 4);
          const uint32_t a = smem_u32(sX + (kb * sp.MT + mt) * sp.pitch), b = smem_u32(sW1 + kb * 8192);
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16(tmem_base + mt * 64, umma_desc_sw128(a + kk * 32), umma_desc_sw128(b + kk * 32), idesc,
                     (kb > 0 || kk > 0) ? 1u : 0u);
        }
      umma_commit(&mma_bar);
    }
    // biases of this warp's 32 columns / this thread's 8 stencil channels: in flight while the MMA runs
    float be[32], bd[8];
    {
      const int cb = min(c0 + g * 32, cexp - 32);
#pragma unroll
      for (int e4 = 0; e4 < 8; ++e4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(sp.b_exp + cb) + e4);
        be[4 * e4] = t.x; be[4 * e4 + 1] = t.y; be[4 * e4 + 2] = t.z; be[4 * e4 + 3] = t.w;
      }
      const int cd = min(c0 + dj * 8, cexp - 8);
      const float4 d0 = __ldg(reinterpret_cast<const float4*>(sp.b_dw + cd));
      const float4 d1 = __ldg(reinterpret_cast<const float4*>(sp.b_dw + cd + 4));
      bd[0] = d0.x; bd[1] = d0.y; bd[2] = d0.z; bd[3] = d0.w; bd[4] = d1.x; bd[5] = d1.y; bd[6] = d1.z; bd[7] = d1.w;
    }
    mbar_wait(&mma_bar, mma_phase, 0x4001);
    mma_phase ^= 1;
    tc_fence_after();
    mb_stamp(5 + 4 * li);
    // W_exp / taps of the next slice: the W1 buffer is free now, the other tap buffer since the last stencil
    if (li + 1 < nmine) load_w1(s + CL);
    mb_cp_commit();
    // ---- epilogue 1: TMEM -> + bias -> swish -> fp16 -> expanded tile in smem ----
    for (int mt = 0; mt < sp.MT; ++mt) {
      if (mt * 128 + q * 32 >= P || g * 32 >= wd) continue;   // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem_base + mt * 64 + g * 32 + ((uint32_t)(q * 32) << 16), v);
      const int row = mt * 128 + q * 32 + lane;
      if (row < P) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int col = g * 32 + jj * 8;
          if (col < wd) {
            const float* bb = be + jj * 8;
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              oh[e] = __floats2half2_rn(swish_t<__half>(__uint_as_float(v[jj * 8 + 2 * e]) + bb[2 * e]),
                                        swish_t<__half>(__uint_as_float(v[jj * 8 + 2 * e + 1]) + bb[2 * e + 1]));
            *reinterpret_cast<uint4*>(sExp + mt * sp.pitch + sw128(row & 127, g * 4 + jj)) = o;
            if (sp.dbg_exp) *reinterpret_cast<uint4*>(sp.dbg_exp + ((size_t)img * P + row) * cexp + c0 + col) = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    mb_stamp(6 + 4 * li);
    // ---- depthwise stencil + BN + swish, squeeze sums ----
    {
      const float* wt = sDw + dj * 8;
      float ssum[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ssum[e] = 0.f;
      if (dj * 8 < wd) {
        constexpr int NI = (SP - 1) * S + K;
        for (int oy = warp; oy < sp.Ho; oy += 8) {
          float acc[SP][8];
#pragma unroll
          for (int p = 0; p < SP; ++p)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[p][e] = 0.f;
          const int ix0 = strip * SP * S - sp.pad;
#pragma unroll
          for (int ty = 0; ty < K; ++ty) {
            const int iy = oy * S - sp.pad + ty;
            if (iy < 0 || iy >= sp.H) continue;
            float in[NI][8];
#pragma unroll
            for (int t = 0; t < NI; ++t) {
              const int ix = ix0 + t;
              if (ix >= 0 && ix < sp.W) {
                const int p = iy * sp.W + ix;
                const uint4 raw = *reinterpret_cast<const uint4*>(sExp + (p >> 7) * sp.pitch + sw128(p & 127, dj));
                const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h[e]);
                  in[t][2 * e] = f.x; in[t][2 * e + 1] = f.y;
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) in[t][e] = 0.f;
              }
            }
#pragma unroll
            for (int tx = 0; tx < K; ++tx) {
              const float4 w0 = lds128f(wt + (ty * K + tx) * 64);
              const float4 w1 = lds128f(wt + (ty * K + tx) * 64 + 4);
              const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int p = 0; p < SP; ++p)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[p][e] = fmaf(w[e], in[p * S + tx][e], acc[p][e]);
            }
          }
#pragma unroll
          for (int p = 0; p < SP; ++p) {
            const int po = oy * sp.Wo + strip * SP + p;
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float y0 = swish_t<__half>(acc[p][2 * e] + bd[2 * e]);
              const float y1 = swish_t<__half>(acc[p][2 * e + 1] + bd[2 * e + 1]);
              ssum[2 * e] += y0; ssum[2 * e + 1] += y1;
              oh[e] = __floats2half2_rn(y0, y1);
            }
            *reinterpret_cast<uint4*>(sA2 + li * sp.a2_slice_bytes + (po >> 7) * sp.pitch_o + sw128(po & 127, dj)) = o;
            if (sp.dbg_dw) *reinterpret_cast<uint4*>(sp.dbg_dw + ((size_t)img * Po + po) * cexp + c0 + dj * 8) = o;
          }
        }
      }
      // squeeze: sum over the 4 strips of the warp (fixed order), then over the 8 warps
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ssum[e] += __shfl_xor_sync(0xffffffffu, ssum[e], 8);
        ssum[e] += __shfl_xor_sync(0xffffffffu, ssum[e], 16);
      }
      if (strip == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) s_part[warp * 64 + dj * 8 + e] = ssum[e];
      }
      __syncthreads();
      mb_stamp(7 + 4 * li);
      if (li + 1 < nmine) { load_dw(s + CL, 0); mb_cp_commit(); }   // the tap buffer is free: next slice's taps
      if (tid < 64) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += s_part[w * 64 + tid];
        pooled[li * 64 + tid] = tid < wd ? a * sp.inv_hw : 0.f;
      }
    }
  }
  mb_cp_wait_all();
  __syncthreads();

  // ---- squeeze-excite (efficientnet/model.py:88-93): FC1 partials -> cluster all-reduce -> FC2 -> gate ----
  cluster_wait();
  mb_stamp(16);
  {
    // every weight load of the FC is issued before the first use (one L2 round trip instead of one per row)
    constexpr int MAXC = MB_MAX_MINE * 2;   // 32-channel groups of this CTA's channels
    float pv[MAXC];
    int chv[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int c = lane + 32 * k;
      const int ch = (rank + (c >> 6) * CL) * MB_SLICE + (c & 63);
      const bool ok = c < nmine * 64 && ch < cexp;
      pv[k] = ok ? pooled[c] : 0.f;
      chv[k] = ok ? ch : 0;
    }
    float a[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int j = min(warp + 8 * it, sp.cse - 1);
      const float* wrow = sp.se_wr + (size_t)j * cexp;
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < MAXC; ++k) t = fmaf(__ldg(wrow + chv[k]), pv[k], t);
      a[it] = t;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float t = a[it];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      const int j = warp + 8 * it;
      if (j < sp.cse && lane < CL) cluster.map_shared_rank(r_all, lane)[rank * 64 + j] = t;
    }
  }
  cluster.sync();
  mb_stamp(17);
  if (tid < sp.cse) {
    float a = __ldg(sp.se_br + tid);
    for (int d = 0; d < CL; ++d) a += r_all[d * 64 + tid];
    r_s[tid] = swish_t<__half>(a);
  }
  __syncthreads();
  if (tid < nmine * 64) {
    const int ch = (rank + (tid >> 6) * CL) * MB_SLICE + (tid & 63);
    float gt = 0.f;
    if (ch < cexp) {
      float a = __ldg(sp.se_be + ch);
      for (int j0 = 0; j0 < sp.cse; j0 += 16) {
        float w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) w[u] = __ldg(sp.se_weT + (size_t)min(j0 + u, sp.cse - 1) * cexp + ch);
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fmaf(w[u], j0 + u < sp.cse ? r_s[j0 + u] : 0.f, a);
      }
      gt = sigmoid_t<__half>(a);
      if (sp.gate_out) sp.gate_out[(size_t)img * cexp + ch] = gt;
    }
    sGate[tid] = gt;
  }
  __syncthreads();
  // gate applied to the A operand of the project GEMM in place (`sigmoid(x_squeezed) * x`, model.py:93)
  for (int idx = tid; idx < nmine * Po * 8; idx += MB_THREADS) {
    const int li = idx / (Po * 8), rem = idx - li * Po * 8;
    const int po = rem >> 3, j = rem & 7;
    uint4* ptr = reinterpret_cast<uint4*>(sA2 + li * sp.a2_slice_bytes + (po >> 7) * sp.pitch_o + sw128(po & 127, j));
    uint4 raw = *ptr;
    __half2* h = reinterpret_cast<__half2*>(&raw);
    const float* gp = sGate + li * 64 + j * 8;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      h[e] = __floats2half2_rn(f.x * gp[2 * e], f.y * gp[2 * e + 1]);
    }
    *ptr = raw;
  }
  fence_async_smem();
  __syncthreads();
  mb_stamp(18);

  // ---- project GEMM, split-K over the cluster: this CTA's channel slices ----
  if (tid == 0 && nmine > 0) {
    tc_fence_after();
    const int nsplit = cout > 256 ? 2 : 1;
    const int nn = cout / nsplit;
    const uint32_t idesc = umma_idesc_f16(128, nn, 0);
    for (int li = 0; li < nmine; ++li) {
      const int wd = min(MB_SLICE, cexp - (rank + li * CL) * MB_SLICE);
      for (int mto = 0; mto < sp.MTo; ++mto)
        for (int h = 0; h < nsplit; ++h) {
          const uint32_t a = smem_u32(sA2 + li * sp.a2_slice_bytes + mto * sp.pitch_o);
          const uint32_t b = smem_u32(sW2 + li * sp.w2_slice_bytes + h * nn * 128);
          for (int kk = 0; kk < (wd >> 4); ++kk)
            umma_f16(tmem_base + sp.d2_col0 + mto * sp.d2_pitch + h * nn, umma_desc_sw128(a + kk * 32),
                     umma_desc_sw128(b + kk * 32), idesc, (li > 0 || kk > 0) ? 1u : 0u);
        }
    }
    umma_commit(&mma_bar);
  }
  if (nmine > 0) {
    mbar_wait(&mma_bar, mma_phase, 0x4002);
    mma_phase ^= 1;
    tc_fence_after();
  }
  mb_stamp(19);
  cluster.sync();   // every CTA's operands are dead: their smem becomes the receive buffer

  mb_stamp(20);
  // ---- reduce-scatter: push this CTA's partial rows to their owners ----
  if (nmine > 0) {
    const int nchunks = sp.d2_pitch >> 5;
    for (int mto = (sp.MTo == 2 ? g : 0); mto < sp.MTo; mto += 2) {
      if (mto * 128 + q * 32 >= Po) continue;   // warp-uniform
      const int row = mto * 128 + q * 32 + lane;
      const int owner = min(row, Po - 1) / sp.rows_own, rl = min(row, Po - 1) - owner * sp.rows_own;
      float* dst = cluster.map_shared_rank(recv, owner) + (size_t)(rank * sp.rows_own + rl) * sp.recv_pitch;
      for (int cc = (sp.MTo == 2 ? 0 : g); cc < nchunks; cc += (sp.MTo == 2 ? 1 : 2)) {
        uint32_t v[32];
        tmem_ld32(tmem_base + sp.d2_col0 + mto * sp.d2_pitch + cc * 32 + ((uint32_t)(q * 32) << 16), v);
        if (row < Po) {
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const int col = cc * 32 + g4 * 4;
            if (col < cout)
              *reinterpret_cast<float4*>(dst + col) = make_float4(__uint_as_float(v[g4 * 4]), __uint_as_float(v[g4 * 4 + 1]),
                                                                  __uint_as_float(v[g4 * 4 + 2]), __uint_as_float(v[g4 * 4 + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  mb_stamp(21);
  cluster.sync();
  mb_stamp(22);

  // ---- owner: sum the partials in rank order, + bias (+ skip) -> fp16 -> global ----
  {
    const int half_n = cout >> 1;
    const int nsrc = min(sp.nsl, CL);   // CTAs that hold a slice
    const int my_rows = max(0, min(sp.rows_own, Po - rank * sp.rows_own));
    for (int idx = tid; idx < my_rows * half_n; idx += MB_THREADS) {
      const int rl = idx / half_n, col = (idx - rl * half_n) * 2;
      const int row = rank * sp.rows_own + rl;
      float2 a = __ldg(reinterpret_cast<const float2*>(sp.b_proj + col));
      for (int d = 0; d < nsrc; ++d) {
        const float2 v = *reinterpret_cast<const float2*>(recv + (size_t)(d * sp.rows_own + rl) * sp.recv_pitch + col);
        a.x += v.x; a.y += v.y;
      }
      if (sp.skip) {
        const float2 r = __half22float2(*reinterpret_cast<const __half2*>(sp.x + ((size_t)img * P + row) * cin + col));
        a.x += r.x; a.y += r.y;
      }
      *reinterpret_cast<__half2*>(sp.out + ((size_t)img * Po + row) * cout + col) = __floats2half2_rn(a.x, a.y);
    }
  }
  __syncthreads();
  mb_stamp(23);
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)sp.tmem_cols);
}

}  // namespace hp
