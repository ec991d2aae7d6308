// One MBConv block of the small feature maps (8x8 / 16x16 pixels per image) as ONE kernel (fast mode, fp16):
//
//   expand 1x1 + BN + swish -> depthwise kxk (stride 1/2, TF-SAME zero padding) + BN + swish -> squeeze-excite
//   (avg-pool, 1x1 + swish, 1x1 + sigmoid, gate) -> project 1x1 + BN (+ identity skip)      efficientnet/model.py:69-104
//
// replaces four launches (expand GEMM, dw3_kernel, se3_kernel, gated project GEMM) and three HBM/L2 round trips of the
// 6x-expanded tensor, which now never leaves the SM.
//
// Decomposition: a thread-block CLUSTER of `cl` (4 or 6) CTAs owns one image and splits the EXPANDED CHANNELS into
// slices of 64 (one 128-byte swizzle row of fp16 per pixel); CTA `rank` takes slices rank, rank + cl, rank + 2 cl.
// The depthwise conv and both BN/swish are per channel, so a channel slice needs no halo and no exchange.
//
// Roles: 16 worker warps + 1 issue warp (one elected lane): everything asynchronous is issued by that lane and
// tracked by single-use mbarriers, so the workers never wait on anything but data.
//   issue lane : TMA of every W_exp slice of this CTA (constants: BEFORE griddepcontrol.wait) -> TMA of x ->
//                tcgen05.mma  D1[slice][pixels x 64] = x . W_exp[slice]^T for ALL slices back to back (each slice has
//                its own TMEM columns and its own commit barrier) -> once the last expand MMA has read its operands
//                the W_proj slices are TMA-loaded over the dead W_exp region -> project MMAs when A2 is gated
//   workers    : per slice: tcgen05.ld -> + bias -> swish -> fp16 -> expanded tile in smem (swizzled: conflict-free
//                for the stencil); depthwise stencil from smem, one output row (or half row) per warp -> + bias ->
//                swish -> fp16 straight into the A operand of the project GEMM, channel sums for the squeeze on the way
//   squeeze-excite: FC1 partial sums over the CTA's channels -> all-reduce over the cluster through distributed
//                shared memory -> FC2 for the CTA's channels -> gate applied to the A operand in place.  The FC weights
//                are fetched into registers ahead of the barriers they would otherwise wait behind.
//   project GEMM split-K over the cluster: D2[pixels x cout] = A2[my slices] . W_proj[:, my slices]^T; the fp32
//                partial tiles go through a per-warp smem transpose to an L2-resident scratch with full-line
//                coalesced stores (an SM moves 64 B/clk to L2 but only ~20 B/clk to a peer's shared memory), one cluster
//                barrier, then every CTA sums the `cl` partials of its own rows in rank order, + bias (+ skip) -> fp16.
// Every reduction runs in a fixed order: results are bitwise reproducible.
#pragma once
#include <cooperative_groups.h>

#include "gemm_tc.cuh"

namespace hp {

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the 16 worker warps only (the issue warp never joins)
__device__ __forceinline__ void mb_workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MB_WORKERS) : "memory"); }

// debug timeline of CTA 0 of the last launch (hmdpose_debug_read("__mb_timeline"), microseconds since entry);
// slots 0..23 are stamped by worker thread 0, slots 24..31 by the issue lane
__device__ unsigned long long g_mb_ts[32];
__device__ __forceinline__ void mb_stamp(int i) {
  if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == MB_WORKERS)) {
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
  __shared__ uint64_t bar_x, bar_w1[MB_MAX_MINE], bar_w2, bar_d1[MB_MAX_MINE], bar_e1, bar_a2, bar_d2;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sX = smem;
  uint8_t* sW = smem + sp.off_w;                                      // W_exp slices, later W_proj slices
  uint8_t* sExp = smem + sp.off_exp;
  uint8_t* sA2 = smem + sp.off_a2;
  float* sDw = reinterpret_cast<float*>(smem + sp.off_dw);            // [nmine][K*K][64] taps
  float* sBe = reinterpret_cast<float*>(smem + sp.off_misc);          // [MB_MAX_MINE*64] expand bias / 2
  float* sBd = sBe + MB_MAX_MINE * 64;                                // depthwise bias / 2
  float* pooled = sBd + MB_MAX_MINE * 64;                             // squeezed means of this CTA's channels
  float* sGate = pooled + MB_MAX_MINE * 64;
  float* fc2p = sGate + MB_MAX_MINE * 64;                             // [2][MB_MAX_MINE*64] halves of the FC2 dot products
  float* s_part = fc2p + 2 * MB_MAX_MINE * 64;                        // [16 warps][64]
  float* r_all = s_part + 16 * 64;                                    // [cl][64]  (written by every CTA of the cluster)
  float* r_s = r_all + MB_CL_MAX * 64;                                // [64]
  uint8_t* sStage = smem;                                             // aliases the operands once the MMAs are done

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster.block_rank();
  const int CL = sp.cl;
  const int img = blockIdx.x / CL;
  const int cin = sp.cin, cexp = sp.cexp, cout = sp.cout, P = sp.P, Po = sp.Po, MT = sp.MT, KB1 = sp.KB1;
  const int nmine = rank < sp.nsl ? (sp.nsl - rank + CL - 1) / CL : 0;

  if (tid == 0) {
    mbar_init(&bar_x, 1);
    mbar_init(&bar_w2, 1);
    mbar_init(&bar_a2, 16);
    mbar_init(&bar_d2, 1);
    mbar_init(&bar_e1, 1);
    for (int i = 0; i < MB_MAX_MINE; ++i) { mbar_init(&bar_w1[i], 1); mbar_init(&bar_d1[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc(&tmem_slot, (uint32_t)sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();      // this CTA holds everything it will ever need (shared memory, TMEM columns)
  cluster_arrive();   // matched by cluster_wait() before the first remote shared-memory access: every CTA has started

  if (warp == 16) {
    // ===================== issue warp =====================
    if (lane == 0) {
      tma_prefetch_desc(sp.tm + 0);
      tma_prefetch_desc(sp.tm + 1);
      tma_prefetch_desc(sp.tm + 2);
      for (int li = 0; li < nmine; ++li) {   // constants: in flight before the previous kernel has finished
        mbar_expect_tx(&bar_w1[li], (uint32_t)(KB1 * 8192));
        for (int kb = 0; kb < KB1; ++kb)
          tma_load_2d(sW + (li * KB1 + kb) * 8192, sp.tm + 1, &bar_w1[li], kb * 64, (rank + li * CL) * MB_SLICE);
      }
      pdl_wait();
      mb_stamp(24);
      if (nmine > 0) {
        mbar_expect_tx(&bar_x, (uint32_t)(KB1 * MT * sp.pitch));
        for (int kb = 0; kb < KB1; ++kb)
          for (int mt = 0; mt < MT; ++mt)
            tma_load_2d(sX + (kb * MT + mt) * sp.pitch, sp.tm + 0, &bar_x, kb * 64, img * P + mt * 128);
        mbar_wait(&bar_x, 0, 0x4001);
        mb_stamp(25);
        for (int li = 0; li < nmine; ++li) {
          const int wd = min(MB_SLICE, cexp - (rank + li * CL) * MB_SLICE);
          mbar_wait(&bar_w1[li], 0, 0x4002);
          tc_fence_after();
          const uint32_t idesc = umma_idesc_f16(128, wd, 0);
          for (int mt = 0; mt < MT; ++mt)
            for (int kb = 0; kb < KB1; ++kb) {
              const int ksteps = min(4, (cin - kb * 64) >> 4);
              const uint32_t a = smem_u32(sX + (kb * MT + mt) * sp.pitch), b = smem_u32(sW + (li * KB1 + kb) * 8192);
              for (int kk = 0; kk < ksteps; ++kk)
                umma_f16(tmem_base + (li * MT + mt) * 64, umma_desc_sw128(a + kk * 32), umma_desc_sw128(b + kk * 32), idesc,
                         (kb > 0 || kk > 0) ? 1u : 0u);
            }
          umma_commit(&bar_d1[li]);
        }
        // W_proj slices over the W_exp region: every expand MMA has read its operands once this commit completes
        umma_commit(&bar_e1);
        mbar_wait(&bar_e1, 0, 0x4003);
        mb_stamp(26);
        const int nsplit = cout > 256 ? 2 : 1, nn = cout / nsplit;
        mbar_expect_tx(&bar_w2, (uint32_t)(nmine * sp.w2_slice_bytes));
        for (int li = 0; li < nmine; ++li)
          for (int h = 0; h < nsplit; ++h)
            tma_load_2d(sW + li * sp.w2_slice_bytes + h * nn * 128, sp.tm + 2, &bar_w2, (rank + li * CL) * MB_SLICE, h * nn);
      }
    }
    __syncwarp();
    cluster_wait();                      // start barrier
    cluster_arrive(); cluster_wait();    // squeeze-excite all-reduce barrier
    if (lane == 0 && nmine > 0) {
      mbar_wait(&bar_w2, 0, 0x4004);
      mbar_wait(&bar_a2, 0, 0x4005);
      tc_fence_after();
      mb_stamp(27);
      const int nsplit = cout > 256 ? 2 : 1, nn = cout / nsplit;
      const uint32_t idesc = umma_idesc_f16(128, nn, 0);
      for (int li = 0; li < nmine; ++li) {
        const int wd = min(MB_SLICE, cexp - (rank + li * CL) * MB_SLICE);
        for (int mto = 0; mto < sp.MTo; ++mto)
          for (int h = 0; h < nsplit; ++h) {
            const uint32_t a = smem_u32(sA2 + li * sp.a2_slice_bytes + mto * sp.pitch_o);
            const uint32_t b = smem_u32(sW + li * sp.w2_slice_bytes + h * nn * 128);
            for (int kk = 0; kk < (wd >> 4); ++kk)
              umma_f16(tmem_base + sp.d2_col0 + mto * sp.d2_pitch + h * nn, umma_desc_sw128(a + kk * 32),
                       umma_desc_sw128(b + kk * 32), idesc, (li > 0 || kk > 0) ? 1u : 0u);
          }
      }
      umma_commit(&bar_d2);
    }
    __syncwarp();
    cluster_arrive(); cluster_wait();    // partial tiles barrier
  } else {
    // ===================== worker warps =====================
    mb_stamp(0);
    // constants -> shared memory (before griddepcontrol.wait): taps and the two bias vectors of my slices
    for (int i = tid; i < nmine * K * K * 16; i += MB_WORKERS) {
      const int li = i / (K * K * 16), r = i - li * (K * K * 16);
      const int tap = r >> 4, c4 = r & 15;
      const int ch = (rank + li * CL) * MB_SLICE + c4 * 4;
      const float4 v = ch < cexp ? __ldg(reinterpret_cast<const float4*>(sp.w_dw + (size_t)tap * cexp + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sDw + (li * K * K + tap) * 64 + c4 * 4) = v;
    }
    for (int i = tid; i < nmine * 64; i += MB_WORKERS) {
      const int ch = (rank + (i >> 6) * CL) * MB_SLICE + (i & 63);
      sBe[i] = ch < cexp ? 0.5f * __ldg(sp.b_exp + ch) : 0.f;    // halved: swish(x) = h + h tanh(h), h = x / 2
      sBd[i] = ch < cexp ? 0.5f * __ldg(sp.b_dw + ch) : 0.f;
    }
    pdl_wait();   // skip connection, partial scratch and the output belong to the previous kernels until here
    mb_workers_sync();
    mb_stamp(1);

    const int dj = lane & 7;            // depthwise: 16-byte channel chunk of this thread
    const int strip = lane >> 3;        // depthwise: strip of SP output pixels (4 strips per warp)
    const int q = warp & 3, g = warp >> 2;

    for (int li = 0; li < nmine; ++li) {
      const int c0s = (rank + li * CL) * MB_SLICE;
      const int wd = min(MB_SLICE, cexp - c0s);
      mbar_wait(&bar_d1[li], 0, 0x4010);
      tc_fence_after();
      if (li == 0) mb_stamp(2);
      // ---- epilogue 1: TMEM -> + bias -> swish -> fp16 -> expanded tile in smem ----
      {
        const int mt = MT == 2 ? (g & 1) : 0;
        const int c0 = MT == 2 ? (g >> 1) * 32 : g * 16;
        const int ncol = MT == 2 ? 32 : 16;
        if (mt * 128 + q * 32 < P && c0 < wd) {   // warp-uniform
          uint32_t v[32];
          const uint32_t ta = tmem_base + (li * MT + mt) * 64 + c0 + ((uint32_t)(q * 32) << 16);
          if (MT == 2) tmem_ld_cols<32>(ta, v); else tmem_ld_cols<16>(ta, v);
          const int row = mt * 128 + q * 32 + lane;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int col = c0 + jj * 8;
            if (jj * 8 < ncol && col < wd) {
              const float4 b0 = lds128f(sBe + li * 64 + col), b1 = lds128f(sBe + li * 64 + col + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float t = fmaf(__uint_as_float(v[jj * 8 + e]), 0.5f, bb[e]);
                y[e] = fmaf(t, tanh_approx(t), t);
              }
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(y[2 * e], y[2 * e + 1]);
              if (row < P) {
                *reinterpret_cast<uint4*>(sExp + mt * sp.pitch + sw128(row & 127, col >> 3)) = o;
                if (sp.dbg_exp) *reinterpret_cast<uint4*>(sp.dbg_exp + ((size_t)img * P + row) * cexp + c0s + col) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mb_workers_sync();
      if (li == 0) mb_stamp(3);
      // ---- depthwise stencil + BN + swish, squeeze sums: one strip of SP output pixels x 8 channels per thread ----
      {
        constexpr int WPR_PX = 4 * SP;               // output pixels of a row covered by one warp
        const int wpr = sp.Wo / WPR_PX;              // warps per output row
        const float* wt = sDw + li * K * K * 64 + dj * 8;
        float ssum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ssum[e] = 0.f;
        if (dj * 8 < wd) {
          const float4 d0 = lds128f(sBd + li * 64 + dj * 8), d1 = lds128f(sBd + li * 64 + dj * 8 + 4);
          const float bd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          constexpr int NI = (SP - 1) * S + K;
          for (int oy = warp / wpr; oy < sp.Ho; oy += 16 / wpr) {
            const int ox0 = (warp % wpr) * WPR_PX + strip * SP;
            float acc[SP][8];
#pragma unroll
            for (int p = 0; p < SP; ++p)
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[p][e] = 0.f;
            const int ix0 = ox0 * S - sp.pad;
#pragma unroll
            for (int ty = 0; ty < K; ++ty) {
              const int iy = oy * S - sp.pad + ty;
              if (iy < 0 || iy >= sp.H) continue;
              float w[K][8];
#pragma unroll
              for (int tx = 0; tx < K; ++tx) {
                const float4 w0 = lds128f(wt + (ty * K + tx) * 64), w1 = lds128f(wt + (ty * K + tx) * 64 + 4);
                w[tx][0] = w0.x; w[tx][1] = w0.y; w[tx][2] = w0.z; w[tx][3] = w0.w;
                w[tx][4] = w1.x; w[tx][5] = w1.y; w[tx][6] = w1.z; w[tx][7] = w1.w;
              }
#pragma unroll
              for (int t = 0; t < NI; ++t) {
                const int ix = ix0 + t;
                if (ix >= 0 && ix < sp.W) {
                  const int pi = iy * sp.W + ix;
                  const uint4 raw = lds128(sExp + (pi >> 7) * sp.pitch + sw128(pi & 127, dj));
                  const __half2* h = reinterpret_cast<const __half2*>(&raw);
                  float in[8];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h[e]);
                    in[2 * e] = f.x; in[2 * e + 1] = f.y;
                  }
#pragma unroll
                  for (int p = 0; p < SP; ++p) {
                    const int tx = t - p * S;
                    if (tx >= 0 && tx < K) {
#pragma unroll
                      for (int e = 0; e < 8; ++e) acc[p][e] = fmaf(w[tx][e], in[e], acc[p][e]);
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int p = 0; p < SP; ++p) {
              const int po = oy * sp.Wo + ox0 + p;
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float t0 = fmaf(acc[p][2 * e], 0.5f, bd[2 * e]), t1 = fmaf(acc[p][2 * e + 1], 0.5f, bd[2 * e + 1]);
                const float y0 = fmaf(t0, tanh_approx(t0), t0), y1 = fmaf(t1, tanh_approx(t1), t1);
                oh[e] = __floats2half2_rn(y0, y1);
                const float2 r = __half22float2(oh[e]);   // the squeeze averages the STORED (fp16) activations
                ssum[2 * e] += r.x; ssum[2 * e + 1] += r.y;
              }
              *reinterpret_cast<uint4*>(sA2 + li * sp.a2_slice_bytes + (po >> 7) * sp.pitch_o + sw128(po & 127, dj)) = o;
              if (sp.dbg_dw) *reinterpret_cast<uint4*>(sp.dbg_dw + ((size_t)img * Po + po) * cexp + c0s + dj * 8) = o;
            }
          }
        }
        // squeeze: sum over the 4 strips of the warp (fixed order), then over the 16 warps
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ssum[e] += __shfl_xor_sync(0xffffffffu, ssum[e], 8);
          ssum[e] += __shfl_xor_sync(0xffffffffu, ssum[e], 16);
        }
        if (strip == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) s_part[warp * 64 + dj * 8 + e] = ssum[e];
        }
        mb_workers_sync();   // also: every read of the expanded tile is done
        if (li == 0) mb_stamp(4);
        if (tid < 64) {
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < 16; ++w) a += s_part[w * 64 + tid];
          pooled[li * 64 + tid] = tid < wd ? a * sp.inv_hw : 0.f;
        }
      }
    }
    mb_stamp(5);

    // ---- squeeze-excite (efficientnet/model.py:88-93): FC1 partials -> cluster all-reduce -> FC2 -> gate ----
    constexpr int MAXC = MB_MAX_MINE * 2;   // 32-channel groups of this CTA's channels
    const int nch = nmine * 64;
    {
      // FC1 weights of rows warp, warp + 16, warp + 32 for this lane's channels: in flight across the barrier below
      float w1r[3][MAXC];
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int j = warp + 16 * it;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
          const int c = lane + 32 * k;
          const int ch = (rank + (c >> 6) * CL) * MB_SLICE + (c & 63);
          w1r[it][k] = (j < sp.cse && c < nch && ch < cexp) ? __ldg(sp.se_wr + (size_t)j * cexp + ch) : 0.f;
        }
      }
      mb_workers_sync();   // pooled[] complete
      float pv[MAXC];
#pragma unroll
      for (int k = 0; k < MAXC; ++k) pv[k] = lane + 32 * k < nch ? pooled[lane + 32 * k] : 0.f;
      cluster_wait();      // every CTA of the cluster has started: its shared memory may be written
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) t = fmaf(w1r[it][k], pv[k], t);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const int j = warp + 16 * it;
        if (j < sp.cse && lane < CL) cluster.map_shared_rank(r_all, lane)[rank * 64 + j] = t;
      }
    }
    // FC2: two threads per channel, each half of the squeezed rows; the weights are fetched before the barrier
    const int jh = (sp.cse + 1) >> 1;
    const int f_half = tid / max(nch, 1), f_chi = tid - f_half * nch;
    const bool f_on = nch > 0 && f_half < 2;
    const int f_ch = f_on ? (rank + (f_chi >> 6) * CL) * MB_SLICE + (f_chi & 63) : 0;
    float w2r[24];
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      const int j = f_half * jh + u;
      w2r[u] = (f_on && f_ch < cexp && u < jh && j < sp.cse) ? __ldg(sp.se_weT + (size_t)j * cexp + f_ch) : 0.f;
    }
    cluster_arrive(); cluster_wait();
    mb_stamp(6);
    if (tid < sp.cse) {
      float a = __ldg(sp.se_br + tid);
      for (int d = 0; d < CL; ++d) a += r_all[d * 64 + tid];
      r_s[tid] = swish_t<__half>(a);
    }
    mb_workers_sync();
    if (f_on) {
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < 24; ++u) {
        const int j = f_half * jh + u;
        a = fmaf(w2r[u], (u < jh && j < sp.cse) ? r_s[j] : 0.f, a);
      }
      fc2p[f_half * (MB_MAX_MINE * 64) + f_chi] = a;
    }
    mb_workers_sync();
    if (tid < nch) {
      const int ch = (rank + (tid >> 6) * CL) * MB_SLICE + (tid & 63);
      float gt = 0.f;
      if (ch < cexp) {
        gt = sigmoid_t<__half>(__ldg(sp.se_be + ch) + fc2p[tid] + fc2p[MB_MAX_MINE * 64 + tid]);
        if (sp.gate_out) sp.gate_out[(size_t)img * cexp + ch] = gt;
      }
      sGate[tid] = gt;
    }
    mb_workers_sync();
    // gate applied to the A operand of the project GEMM in place (`sigmoid(x_squeezed) * x`, model.py:93)
    for (int idx = tid; idx < nmine * Po * 8; idx += MB_WORKERS) {
      const int li = idx / (Po * 8), rem = idx - li * Po * 8;
      const int po = rem >> 3, j = rem & 7;
      uint4* ptr = reinterpret_cast<uint4*>(sA2 + li * sp.a2_slice_bytes + (po >> 7) * sp.pitch_o + sw128(po & 127, j));
      uint4 raw = *ptr;
      __half2* h = reinterpret_cast<__half2*>(&raw);
      const float4 g0 = lds128f(sGate + li * 64 + j * 8), g1 = lds128f(sGate + li * 64 + j * 8 + 4);
      float2 f;
      f = __half22float2(h[0]); h[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
      f = __half22float2(h[1]); h[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
      f = __half22float2(h[2]); h[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
      f = __half22float2(h[3]); h[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
      *ptr = raw;
    }
    fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_a2);
    mb_stamp(7);

    // ---- split-K partial tile: TMEM -> per-warp transpose -> L2 scratch, full 128-byte lines ----
    if (nmine > 0) {
      mbar_wait(&bar_d2, 0, 0x4011);
      tc_fence_after();
      mb_stamp(8);
      uint8_t* stg = sStage + warp * MB_STAGE_WARP_BYTES;
      const int nchunks = sp.d2_pitch >> 5;
      float* part = sp.part + ((size_t)(img * CL + rank) * Po) * cout;
      for (int u = g; u < sp.MTo * nchunks; u += 4) {
        const int mto = u / nchunks, cc = u - mto * nchunks;
        const int row0 = mto * 128 + q * 32;
        if (row0 >= Po) continue;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + sp.d2_col0 + mto * sp.d2_pitch + cc * 32 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128(stg + (lane * 36 + j4 * 4) * 4, make_uint4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
        __syncwarp();
        const int c = cc * 32 + (lane & 7) * 4;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = rr * 4 + (lane >> 3);
          const uint4 t = lds128(stg + (r * 36 + (lane & 7) * 4) * 4);
          if (row0 + r < Po && c < cout) *reinterpret_cast<uint4*>(part + (size_t)(row0 + r) * cout + c) = t;
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    mb_stamp(9);
    cluster_arrive(); cluster_wait();   // release / acquire at cluster scope: every partial tile of this image is visible
    mb_stamp(10);

    // ---- owner: sum the partials in rank order, + bias (+ skip) -> fp16 -> global ----
    {
      const int c4n = cout >> 2;
      const int my_rows = max(0, min(sp.rows_own, Po - rank * sp.rows_own));
      const float* pimg = sp.part + (size_t)img * CL * Po * cout;
      for (int idx = tid; idx < my_rows * c4n; idx += MB_WORKERS) {
        const int rl = idx / c4n, col = (idx - rl * c4n) * 4;
        const int row = rank * sp.rows_own + rl;
        float4 a = __ldg(reinterpret_cast<const float4*>(sp.b_proj + col));
        for (int d = 0; d < CL; ++d) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(pimg + ((size_t)d * Po + row) * cout + col));
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (sp.skip) {
          const uint2 rv = *reinterpret_cast<const uint2*>(sp.x + ((size_t)img * P + row) * cin + col);
          const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&rv.x));
          const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&rv.y));
          a.x += r0.x; a.y += r0.y; a.z += r1.x; a.w += r1.y;
        }
        uint2 o;
        *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(a.x, a.y);
        *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(a.z, a.w);
        *reinterpret_cast<uint2*>(sp.out + ((size_t)img * Po + row) * cout + col) = o;
      }
    }
    mb_stamp(11);
  }
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, (uint32_t)sp.tmem_cols);
}

}  // namespace hp
