// Expand 1x1 + BN + swish  ->  depthwise kxk (stride 1/2, TF-SAME zero padding) + BN + swish + squeeze sums
// of an MBConv block on the LARGE feature maps as ONE kernel (fast mode, fp16)          efficientnet/model.py:76-89
//
// The 6x-expanded tensor is the largest tensor of a block (59 MB at batch 16 for block 1) and the four-launch path
// writes it once and reads it once.  Here it never leaves the SM: a tile's expanded pixels are recomputed from the
// (6x smaller) block input, halo included, and consumed by the stencil straight from shared memory.
//
// Tile = 16 x 16 INPUT pixels (halo included) -> TO x TO outputs, TO = (16 - k) / stride + 1 (14 / 12 / 7 / 6).
//   issue lane : one 4-D TMA box {64 ch, 16, 16, 1} per tile (out-of-image pixels and channels >= cin are zero-filled)
//                lands as 256 pixel rows of 128 bytes = two K-major SWIZZLE_128B UMMA A tiles; W_exp (one TMA for the
//                whole kernel) is the B operand; 2 x ksteps tcgen05.mma (M = 128, N = cexp <= 256) -> TMEM
//   workers    : epilogue: tcgen05.ld -> + bias -> swish -> ZERO for pixels outside the image (the reference pads the
//                EXPANDED tensor with zeros, and swish(bias) != 0) -> fp16 -> expanded tile [256 px][cexp] in smem;
//                stencil: thread = (8-channel chunk, row strip), taps from smem -> + bias -> swish -> fp16 -> global,
//                per-tile channel sums for the squeeze (fixed order: bitwise reproducible) -> se_partial.
// Persistent CTAs (one per SM); the TMA of tile i+1 is issued as soon as the MMAs of tile i have read the window and
// the MMAs of tile i+1 run while the workers are still in the stencil of tile i, so the workers never wait for data.
#pragma once
#include "mbconv_tc.cuh"

namespace hp {

__device__ __forceinline__ void ed_workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ED_WORKERS) : "memory"); }

template <int K, int S, int SP>
__global__ void __launch_bounds__(ED_THREADS, 1)
expdw_kernel(const EdSpec sp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_w1, bar_x, bar_mma, bar_tfree;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sX = smem;
  uint8_t* sW1 = smem + sp.off_w1;
  uint8_t* sE = smem + sp.off_e;
  float* sDw = reinterpret_cast<float*>(smem + sp.off_dw);   // [K*K][2][nchunk][4] taps: channels 0-3 / 4-7 of every chunk apart (conflict-free float4 reads)
  float* sBe = reinterpret_cast<float*>(smem + sp.off_b);    // [cexp] expand bias / 2
  float* sBd = sBe + sp.cexp;                                // [cexp] depthwise bias / 2
  const bool red_alias = sp.off_red < 0;                     // no room for a separate scratch: alias the expanded tile between tiles
  float* red = reinterpret_cast<float*>(red_alias ? sE : smem + sp.off_red);   // [groups][cexp]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cexp = sp.cexp, nchunk = sp.nchunk, e_pitch = sp.e_pitch;
  constexpr int TO = (ED_WIN - K) / S + 1;
  constexpr int NSTRIP = (TO + SP - 1) / SP;
  constexpr int NRS = TO * NSTRIP;          // (output row, strip) pairs of a tile
  constexpr int NI = (SP - 1) * S + K;      // input pixels of a strip row

  if (tid == 0) {
    mbar_init(&bar_w1, 1);
    mbar_init(&bar_x, 1);
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tfree, ED_WORKERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ED_WORKERS / 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();

  if (warp == ED_WORKERS / 32) {
    // ===================== issue warp =====================
    if (lane == 0) {
      tma_prefetch_desc(sp.tm + 0);
      tma_prefetch_desc(sp.tm + 1);
      mbar_expect_tx(&bar_w1, (uint32_t)(cexp * 128));   // constants: before the previous kernel has finished
      tma_load_2d(sW1, sp.tm + 1, &bar_w1, 0, 0);
      pdl_wait();
      const uint32_t idesc = umma_idesc_f16(128, cexp, 0);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < sp.total_tiles; t += gridDim.x, ++it) {
        const int b = t / sp.tiles_per_img, r = t - b * sp.tiles_per_img;
        const int ty = r / sp.tiles_x, tx = r - ty * sp.tiles_x;
        if (it > 0) mbar_wait(&bar_mma, (it - 1) & 1, 0x5001);   // the MMAs of the previous tile have read the window
        mbar_expect_tx(&bar_x, (uint32_t)(ED_WIN * ED_WIN * 128));
        tma_load_4d(sX, sp.tm + 0, &bar_x, 0, tx * TO * S - sp.pad, ty * TO * S - sp.pad, b);
        mbar_wait(&bar_x, it & 1, 0x5002);
        if (it == 0) mbar_wait(&bar_w1, 0, 0x5003);
        else mbar_wait(&bar_tfree, (it - 1) & 1, 0x5004);        // the epilogue of the previous tile has drained TMEM
        tc_fence_after();
        for (int mt = 0; mt < 2; ++mt)
          for (int kk = 0; kk < sp.ksteps; ++kk)
            umma_f16(tmem_base + mt * 256, umma_desc_sw128(smem_u32(sX + mt * 16384) + kk * 32),
                     umma_desc_sw128(smem_u32(sW1) + kk * 32), idesc, kk > 0 ? 1u : 0u);
        umma_commit(&bar_mma);
      }
    }
    __syncwarp();
  } else {
    // ===================== worker warps =====================
    for (int i = tid; i < K * K * cexp; i += ED_WORKERS) {
      const int tap = i / cexp, c = i - tap * cexp;
      sDw[(tap * 2 + ((c >> 2) & 1)) * (nchunk * 4) + (c >> 3) * 4 + (c & 3)] = __ldg(sp.w_dw + i);
    }
    for (int i = tid; i < cexp; i += ED_WORKERS) {
      sBe[i] = 0.5f * __ldg(sp.b_exp + i);   // halved: swish(x) = h + h tanh(h), h = x / 2
      sBd[i] = 0.5f * __ldg(sp.b_dw + i);
    }
    pdl_wait();   // the output / squeeze buffers may still be read by the previous kernels
    ed_workers_sync();

    // epilogue role: TMEM lane quadrant q; units (M tile, 16-column chunk) eg, eg + 3, ...
    const int q = warp & 3, eg = warp >> 2;
    const int nunit = 2 * (cexp >> 4);
    // stencil role: 8-channel chunk `ch` of (row, strip) pairs grp, grp + G, ...
    const int G = ED_WORKERS / nchunk;
    const int ch = tid % nchunk, grp = tid / nchunk;
    const bool s_on = grp < G;

    uint32_t it = 0;
    for (int t = blockIdx.x; t < sp.total_tiles; t += gridDim.x, ++it) {
      const int b = t / sp.tiles_per_img, r = t - b * sp.tiles_per_img;
      const int ty = r / sp.tiles_x, tx = r - ty * sp.tiles_x;
      const int oy0 = ty * TO, ox0 = tx * TO;
      const int iy0 = oy0 * S - sp.pad, ix0 = ox0 * S - sp.pad;
      mbar_wait(&bar_mma, it & 1, 0x5010);
      tc_fence_after();
      // ---- epilogue: TMEM -> + bias -> swish -> zero outside the image -> fp16 -> expanded tile ----
      for (int u = eg; u < nunit; u += 3) {
        const int mt = u & 1, c = u >> 1;
        const int erow = mt * 128 + q * 32 + lane;      // window pixel of this thread: (erow >> 4, erow & 15)
        const int iy = iy0 + (erow >> 4), ix = ix0 + (erow & 15);
        const bool inside = iy >= 0 && iy < sp.H && ix >= 0 && ix < sp.W;
        uint8_t* erowp = sE + (size_t)erow * e_pitch;
        __half* dbg = (sp.dbg_exp && inside) ? sp.dbg_exp + (((size_t)b * sp.H + iy) * sp.W + ix) * cexp : nullptr;
        uint32_t v[16];
        tmem_ld_cols<16>(tmem_base + mt * 256 + c * 16 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) {
          const float4 b0 = lds128f(sBe + c * 16 + j8 * 8), b1 = lds128f(sBe + c * 16 + j8 * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float h = fmaf(__uint_as_float(v[j8 * 8 + e]), 0.5f, bb[e]);
            y[e] = inside ? fmaf(h, tanh_approx(h), h) : 0.f;
          }
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(y[2 * e], y[2 * e + 1]);
          sts128(erowp + c * 32 + j8 * 16, o);
          if (dbg) *reinterpret_cast<uint4*>(dbg + c * 16 + j8 * 8) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree);
      ed_workers_sync();
      // ---- depthwise stencil + BN + swish + squeeze sums ----
      float ssum[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ssum[e] = 0.f;
      if (s_on) {
        const float* wt = sDw + ch * 4;
        const int wstep = nchunk * 4;   // floats between the two halves of a tap
        for (int rs = grp; rs < NRS; rs += G) {
          const int oy = rs / NSTRIP, oxl = (rs - oy * NSTRIP) * SP;
          float acc[SP][8];
#pragma unroll
          for (int p = 0; p < SP; ++p)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[p][e] = 0.f;
#pragma unroll
          for (int kty = 0; kty < K; ++kty) {
            float w[K][8];
#pragma unroll
            for (int ktx = 0; ktx < K; ++ktx) {
              const float4 w0 = lds128f(wt + (kty * K + ktx) * 2 * wstep), w1 = lds128f(wt + ((kty * K + ktx) * 2 + 1) * wstep);
              w[ktx][0] = w0.x; w[ktx][1] = w0.y; w[ktx][2] = w0.z; w[ktx][3] = w0.w;
              w[ktx][4] = w1.x; w[ktx][5] = w1.y; w[ktx][6] = w1.z; w[ktx][7] = w1.w;
            }
            const uint8_t* rowp = sE + (size_t)((oy * S + kty) * ED_WIN + oxl * S) * e_pitch + ch * 16;
#pragma unroll
            for (int ti = 0; ti < NI; ++ti) {
              if (oxl * S + ti < ED_WIN) {   // the last strip of a row may be partial
                const uint4 raw = lds128(rowp + (size_t)ti * e_pitch);
                const __half2* h = reinterpret_cast<const __half2*>(&raw);
                float in[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h[e]);
                  in[2 * e] = f.x; in[2 * e + 1] = f.y;
                }
#pragma unroll
                for (int p = 0; p < SP; ++p) {
                  const int ktx = ti - p * S;
                  if (ktx >= 0 && ktx < K) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[p][e] = fmaf(w[ktx][e], in[e], acc[p][e]);
                  }
                }
              }
            }
          }
          const int oyg = oy0 + oy;
          __half* orow = sp.out + (((size_t)b * sp.Ho + oyg) * sp.Wo + ox0 + oxl) * cexp + ch * 8;
          const float4 d0 = lds128f(sBd + ch * 8), d1 = lds128f(sBd + ch * 8 + 4);
          const float bd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int p = 0; p < SP; ++p) {
            if (oxl + p < TO && oyg < sp.Ho && ox0 + oxl + p < sp.Wo) {
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float t0 = fmaf(acc[p][2 * e], 0.5f, bd[2 * e]), t1 = fmaf(acc[p][2 * e + 1], 0.5f, bd[2 * e + 1]);
                oh[e] = __floats2half2_rn(fmaf(t0, tanh_approx(t0), t0), fmaf(t1, tanh_approx(t1), t1));
                const float2 rr = __half22float2(oh[e]);   // the squeeze averages the STORED (fp16) activations
                ssum[2 * e] += rr.x; ssum[2 * e + 1] += rr.y;
              }
              *reinterpret_cast<uint4*>(orow + (size_t)p * cexp) = o;
            }
          }
        }
      }
      // ---- squeeze: per-tile channel sums in a fixed order ----
      if (sp.se_partial) {
        if (red_alias) ed_workers_sync();   // every read of the expanded tile is done: its memory becomes the scratch
        if (s_on) {
          *reinterpret_cast<float4*>(red + grp * cexp + ch * 8) = make_float4(ssum[0], ssum[1], ssum[2], ssum[3]);
          *reinterpret_cast<float4*>(red + grp * cexp + ch * 8 + 4) = make_float4(ssum[4], ssum[5], ssum[6], ssum[7]);
        }
        ed_workers_sync();   // (separate scratch: this is also the barrier that frees the expanded tile)
        if (tid < cexp) {
          const int gu = G < NRS ? G : NRS;
          float a = 0.f;
          for (int g2 = 0; g2 < gu; ++g2) a += red[g2 * cexp + tid];
          sp.se_partial[((size_t)b * sp.tiles_per_img + r) * cexp + tid] = a;
        }
        if (red_alias) ed_workers_sync();   // before the next epilogue overwrites the scratch
      } else {
        ed_workers_sync();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ED_WORKERS / 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace hp
