// Fused depthwise-separable convolution for the 64-channel pyramid (BiFPN nodes and head layers), fast mode.
//
//   out = act( pw( dw3x3( x ) ) + bias ),   x = feature map, or the BiFPN node input
//   x = swish(w0*a + w1*resample(b) + w2*resample(c))  (efficientdet/model.py:215-264; SeparableConvBlock :42-52)
//
// One CTA = one tile of 128 consecutive output pixels (whole rows of one image, or several whole small images):
//   A  all threads build the input tile + one halo row above/below in shared memory (fp16), evaluating the
//      fast-normalised fusion, nearest upsample and zero-padded 3x3/2 max-pool on the fly;
//   B  all threads run the 3x3 depthwise stencil from shared memory (fp32 accumulate) and write the result as
//      the K-major, 128B-swizzled A operand of the pointwise GEMM (128 pixels x 64 channels = one k-block);
//   C  one thread issues tcgen05.mma (M=128, N=bn, K=64) against the TMA-loaded pointwise weights, the
//      accumulator lives in TMEM; all 8 warps run the epilogue (bias, activation, coalesced stores / head scatter).
// The depthwise output never touches HBM and a BiFPN node is one launch instead of three.
#pragma once
#include "gemm_tc.cuh"

namespace hp {

struct __align__(64) SepProb {
  CUtensorMap tmW;   // pointwise weights [N][64] fp16: dims {64, N}, box {64, bn}, SWIZZLE_128B
  GemmProb p;        // epilogue description (bias, out, M = B*H*W, N, ldo, act, out_mode..., bn, n_tiles = #n chunks)
  const void* in;    // a: [B,H,W,64] fp16
  const void* fb; const void* fc;
  const float* dw_w; // [9][64] fp32
  int H, W, Bn, fused, mode_b, mode_c;
  float w0, w1, w2;
};

constexpr int SEP_THREADS = 256;
constexpr int SEP_STAGE_BYTES = 34816;  // >= max staged tile (33 KB) and >= 8 epilogue staging tiles (33.8 KB)
constexpr int SEP_D0_STATE_BYTES = 2 * 128 * 9 * 8;   // out_mode 2: (max, arg-max) of 9 anchors x 128 pixels x 2 column halves
__host__ __device__ inline int sep_smem_bytes(int bn_max, bool d0_state = false) {   // bn_max <= 64 -> 75 KB -> three CTAs per SM
  return 1024 + TC_A_STAGE_BYTES + 2 * bn_max * 128 + SEP_STAGE_BYTES + 9 * 64 * 4 + 8 * 128 * 4 +
         (d0_state ? SEP_D0_STATE_BYTES : 0);
}
__device__ __forceinline__ uint32_t sep_float_sortable(float f) {   // = float_sortable of postprocess.cu
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// chain_len > 0: CHAIN launch.  CTA g walks probs[0..chain_len) in order, each restricted to the images
// [g*chain_nb, (g+1)*chain_nb): consecutive BiFPN nodes of the small pyramid levels (<= 128 pixels per chain_nb images)
// only depend on the same images, so one CTA runs them back to back with a block barrier in between instead of one
// launch per node (efficientdet/model.py:215-264: P6_up -> P5_up, and P5_out -> P6_out -> P7_out -> next cell's
// P6_up -> P5_up).
// (CHAIN only separates the two launch flavours in profiles; a deeper unroll / no register cap for the chain
// instantiation was measured slower.)
template <bool CHAIN>
__global__ void __launch_bounds__(SEP_THREADS, CHAIN ? 1 : 2) sepconv_kernel(const SepProb* __restrict__ probs, int nprobs,
                                                                             int bn_max, int chain_len, int chain_nb) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t wfull[2], mma_done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = sA + TC_A_STAGE_BYTES;
  const int w_buf = bn_max * 128;
  uint8_t* sStage = sW + 2 * w_buf;
  float* sDw = reinterpret_cast<float*>(sStage + SEP_STAGE_BYTES);
  float* sBias = sDw + 9 * 64;
  float* sD0max = sBias + 8 * 128;                              // out_mode 2 only: [2 halves][128 rows][9 anchors]
  int* sD0arg = reinterpret_cast<int*>(sD0max + 2 * 128 * 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&wfull[0], 1);
    mbar_init(&wfull[1], 1);
    mbar_init(&mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();   // after the TMEM allocation: dependents must never hold columns this grid still waits for (gemm_tc.cuh)
  uint32_t wcnt = 0;   // weight chunks consumed so far: buffer = wcnt & 1, barrier phase = (wcnt >> 1) & 1
  const int nsteps = CHAIN ? chain_len : 1;
  for (int step = 0; step < nsteps; ++step) {
  // problem lookup: every lane tests one table entry, one round trip instead of a dependent linear scan
  int pi = 0;
  if (CHAIN) {
    pi = step;
  } else {
    for (int base = 0; base < nprobs; base += 32) {
      const int q = base + lane;
      const bool le = q < nprobs && probs[q].p.tile_start <= (int)blockIdx.x;
      pi += __popc(__ballot_sync(0xffffffffu, le));
    }
    pi -= 1;
  }
  const SepProb* sp = probs + pi;
  const GemmProb& p = sp->p;
  const int H = sp->H, W = sp->W, HW = H * W, Bn = sp->Bn;
  const int bn = p.bn, n_chunks = p.n_tiles;

  // tile geometry: `segs` segments of `rps` rows each (one image per segment); W, H*W are powers of two
  const int lgW = 31 - __clz(W), lgHW = 31 - __clz(HW);
  const int rps = HW >= 128 ? (128 >> lgW) : H;
  const int segs = CHAIN ? chain_nb : (HW >= 128 ? 1 : (128 >> lgHW));
  const int P0 = CHAIN ? ((int)blockIdx.x * chain_nb) << lgHW : ((int)blockIdx.x - p.tile_start) * 128;
  const int b0 = P0 >> lgHW;
  const int row0 = HW >= 128 ? ((P0 - (b0 << lgHW)) >> lgW) : 0;
  const int M_lim = CHAIN ? min(p.M, P0 + (chain_nb << lgHW)) : p.M;   // rows this CTA may write
  const int valid_px = CHAIN ? (chain_nb << lgHW) : 128;                    // pixels of the tile that exist

  const bool stamp = CHAIN && step < 2 && tid == 0;
  if (stamp) s3_stamp(16 + step * 8 + 0);
  if (CHAIN && step > 0) {      // the previous node's outputs (written by this CTA) are this node's inputs
    __threadfence();
    __syncthreads();
  }
  if (stamp) s3_stamp(16 + step * 8 + 1);
  if (tid == 0) {
    mbar_expect_tx(&wfull[wcnt & 1], bn * 128);
    tma_load_2d(sW + (wcnt & 1) * w_buf, &sp->tmW, &wfull[wcnt & 1], 0, 0);
  }
  for (int i = tid; i < 9 * 64; i += SEP_THREADS) sDw[i] = __ldg(sp->dw_w + i);
  if (step == 0) pdl_wait();   // weights (TMA above, taps) are constants; the feature maps below come from the previous kernel

  // ---- phase A: input tile (+1 halo row each side) -> shared memory, fp16 [staged pixel][64] ----
  // thread = (element e of a staged row, row slot); rows advance by a constant stride so (seg, ry) are updated
  // incrementally.  Plain inputs go global -> shared with cp.async (everything in flight at once).
  {
    const __half* in = reinterpret_cast<const __half*>(sp->in);
    const int rows = rps + 2;
    const int total_rows = segs * rows;
    const int epr = W << 3;                                   // 16-byte vectors per staged row (power of two)
    const int lg_epr = lgW + 3;
    const int rpp = epr >= SEP_THREADS ? 1 : (SEP_THREADS >> lg_epr);   // staged rows per pass
    const int e0 = epr >= SEP_THREADS ? tid : (tid & (epr - 1));
    const int estep = epr >= SEP_THREADS ? SEP_THREADS : epr;
    int srow = epr >= SEP_THREADS ? 0 : (tid >> lg_epr);
    int seg = 0, ry = srow;
    while (ry >= rows) { ry -= rows; ++seg; }
    uint8_t* stage_a = sStage;
    const bool fused = sp->fused != 0;
    if (!fused) {
      for (; srow < total_rows; srow += rpp) {
        const int b = b0 + seg;
        const int y = row0 + ry - 1;
        const bool row_ok = y >= 0 && y < H && b < Bn;
        for (int e = e0; e < epr; e += estep) {
          const int cv = e & 7, x = e >> 3;
          uint8_t* dst = stage_a + (uint32_t)((srow << lg_epr) + e) * 16;
          const __half* src = in + ((((long long)b * H + y) << lgW) + x) * 64 + cv * 8;
          cp_async16(smem_u32(dst), row_ok ? (const void*)src : (const void*)in, row_ok ? 16 : 0);
        }
        ry += rpp;
        while (ry >= rows) { ry -= rows; ++seg; }
      }
    } else {
      // BiFPN node input: flat loop over the staged 16-byte vectors (independent iterations, plain stores, so the
      // compiler overlaps the resampled loads of consecutive items)
      const int items = total_rows << lg_epr;
      const float w0 = sp->w0, w1 = sp->w1, w2 = sp->w2;
      const int mode_b = sp->mode_b, mode_c = sp->mode_c;
      const __half* fb = reinterpret_cast<const __half*>(sp->fb);
      const __half* fc = reinterpret_cast<const __half*>(sp->fc);
#pragma unroll 2
      for (int it = tid; it < items; it += SEP_THREADS) {
        const int cv = it & 7;
        int px = it >> 3;
        const int x = px & (W - 1);
        px >>= lgW;
        const int sg = px / rows, r_ = px - sg * rows;
        const int b = b0 + sg;
        const int y = row0 + r_ - 1;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < H && b < Bn) {
          const int c0 = cv * 8;
          float a[8], v[8];
          ldv<__half>(in + ((((long long)b * H + y) << lgW) + x) * 64 + c0, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = w0 * a[j];
          if (mode_b != RS_NONE) {
            float t[8];
            fetch_rs<__half>(fb, mode_b, b, y, x, H, W, 64, c0, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] + w1 * t[j];
          }
          if (mode_c != RS_NONE) {
            float t[8];
            fetch_rs<__half>(fc, mode_c, b, y, x, H, W, 64, c0, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] + w2 * t[j];
          }
          __half2* h2 = reinterpret_cast<__half2*>(&val);
#pragma unroll
          for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(swish_t<__half>(v[2 * j]), swish_t<__half>(v[2 * j + 1]));
        }
        sts128(stage_a + (size_t)it * 16, val);
            }
    }
    cp_async_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (stamp) s3_stamp(16 + step * 8 + 2);

  // ---- phase B: 3x3 depthwise stencil from shared memory -> swizzled A operand ----
  {
    const int rows = rps + 2;
    const uint8_t* stage_b = sStage;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + SEP_THREADS * i;
      const int cv = idx & 7, pix = idx >> 3;   // pix: 0..127
      if (CHAIN && pix >= valid_px) continue;   // rows of the MMA tile that no image of this CTA owns stay garbage
      int seg, ly, x;
      if (HW >= 128) { seg = 0; ly = pix >> lgW; x = pix & (W - 1); }
      else { seg = pix >> lgHW; const int rem = pix & (HW - 1); ly = rem >> lgW; x = rem & (W - 1); }
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = x + dx - 1;
          if (xx < 0 || xx >= W) continue;
          const uint4 raw = lds128(stage_b + (uint32_t)(((((seg * rows + ly + dy) << lgW) + xx) << 3) + cv) * 16);
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
          const float4 w0 = *reinterpret_cast<const float4*>(sDw + (dy * 3 + dx) * 64 + cv * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sDw + (dy * 3 + dx) * 64 + cv * 8 + 4);
          float2 f;
          f = __half22float2(h2[0]); acc[0] = fmaf(f.x, w0.x, acc[0]); acc[1] = fmaf(f.y, w0.y, acc[1]);
          f = __half22float2(h2[1]); acc[2] = fmaf(f.x, w0.z, acc[2]); acc[3] = fmaf(f.y, w0.w, acc[3]);
          f = __half22float2(h2[2]); acc[4] = fmaf(f.x, w1.x, acc[4]); acc[5] = fmaf(f.y, w1.y, acc[5]);
          f = __half22float2(h2[3]); acc[6] = fmaf(f.x, w1.z, acc[6]); acc[7] = fmaf(f.y, w1.w, acc[7]);
        }
      }
      uint4 pk;
      __half2* o2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      sts128(sA + pix * 128 + ((cv ^ (pix & 7)) << 4), pk);   // SWIZZLE_128B, K-major
    }
  }
  if (p.out_mode == 2)
    for (int i = tid; i < 2 * 128 * 9; i += SEP_THREADS) { sD0max[i] = -INFINITY; sD0arg[i] = 0; }
  fence_async_smem();
  __syncthreads();

  if (stamp) s3_stamp(16 + step * 8 + 3);
  // ---- phase C: pointwise GEMM on the tensor core, chunk by chunk over the output channels ----
  const int q = warp & 3, h = warp >> 2;
  uint8_t* stg_a = sStage + warp * TC2_EPI_WARP_BYTES;
  float* bias_s = sBias + warp * 128;
  const float* bias_a = bias_s;
  const int M = M_lim, N = p.N, act = p.act;
  const float bias_sc = (p.out_mode == 0 && act == ACT_SWISH) ? 0.5f : 1.0f;   // epi_cols_f16 takes bias / 2 for swish
  const int mrow0 = P0 + q * 32;
  for (int c = 0; c < n_chunks; ++c) {
    const int n0 = c * bn;
    const uint32_t wc = wcnt + (uint32_t)c;   // running weight-chunk / MMA-commit index across chain steps
    if (tid == 0) {
      if (c + 1 < n_chunks) {
        mbar_expect_tx(&wfull[(wc + 1) & 1], bn * 128);
        tma_load_2d(sW + ((wc + 1) & 1) * w_buf, &sp->tmW, &wfull[(wc + 1) & 1], 0, (c + 1) * bn);
      }
      mbar_wait(&wfull[wc & 1], (wc >> 1) & 1);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(TC_BM, bn, 0);
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sW + (wc & 1) * w_buf);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, k > 0 ? 1u : 0u);
      umma_commit(&mma_done);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + lane + 32 * j;
      bias_s[lane + 32 * j] = (lane + 32 * j < bn && n < N) ? bias_sc * __ldg(p.bias + n) : 0.f;
    }
    __syncwarp();
    mbar_wait(&mma_done, wc & 1);
    tc_fence_after();
    if (stamp && c == 0) s3_stamp(16 + step * 8 + 4);
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nchunks32 = (bn + 31) >> 5;
    if (p.out_mode == 0) {
      const int ldo = p.ldo;
      const int rows_valid = M - mrow0;
      const long long row_step = 8LL * ldo;
      const long long o0 = (long long)(mrow0 + (lane >> 2)) * ldo + n0 + (lane & 3) * 8;
      __half* gout = reinterpret_cast<__half*>(p.out) + o0;
      for (int cc = h; cc < nchunks32; cc += 2) {
        const int c0 = cc * 32;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)c0, v);
        const bool cols_ok = (c0 + (lane & 3) * 8 < bn) && (n0 + c0 + (lane & 3) * 8 < N);
        if (act == ACT_SWISH)
          epi_cols_f16<32, ACT_SWISH, false>(v, bias_a + c0, stg_a, lane, gout + c0, nullptr, row_step, rows_valid, cols_ok);
        else
          epi_cols_f16<32, ACT_NONE, false>(v, bias_a + c0, stg_a, lane, gout + c0, nullptr, row_step, rows_valid, cols_ok);
      }
    } else if (p.out_mode == 2) {
      // EfficientDet-d0 classifier header: running max / first arg-max over the classes of every anchor, per row
      // (thread = pixel row q * 32 + lane; the two column halves h keep separate states, combined after the last chunk)
      const int C = p.p_src;
      float* st_m = sD0max + (h * 128 + q * 32 + lane) * 9;
      int* st_a = sD0arg + (h * 128 + q * 32 + lane) * 9;
      for (int cc = h; cc < nchunks32; cc += 2) {
        const int c0 = cc * 32;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)c0, v);
        int a = (n0 + c0) / C, k = (n0 + c0) - a * C;
        float cm = a < 9 ? st_m[a] : -INFINITY;
        int ca = a < 9 ? st_a[a] : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + j < bn && n0 + c0 + j < N) {
            const float val = apply_act<__half>(__uint_as_float(v[j]) + bias_s[c0 + j], act);
            if (val > cm) { cm = val; ca = k; }   // strict: the first maximum wins, like torch.max
            if (++k == C) {
              st_m[a] = cm; st_a[a] = ca;
              ++a; k = 0;
              cm = a < 9 ? st_m[a] : -INFINITY;
              ca = a < 9 ? st_a[a] : 0;
            }
          }
        }
        if (a < 9) { st_m[a] = cm; st_a[a] = ca; }
      }
    } else {
      float* tile_s = reinterpret_cast<float*>(sStage + warp * TC2_EPI_WARP_BYTES);
      float* outp = reinterpret_cast<float*>(p.out);
      for (int cc = h; cc < nchunks32; cc += 2) {
        const int c0 = cc * 32;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = apply_act<__half>(__uint_as_float(v[j]) + bias_s[c0 + j], act);
        __syncwarp();
        const int n = n0 + c0 + lane;
        if (c0 + lane < bn && n < N) {
          const int a = n / p.p_src, qq = n - a * p.p_src;
          const int coff = a * p.p_dst + p.p_off + qq;
          for (int r = 0; r < 32; ++r) {
            const int m = mrow0 + r;
            if (m < M) {
              const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
              { float* dstp = outp + img * p.img_stride + (long long)pix * p.pix_stride + coff; *dstp = p.accumulate ? *dstp + tile_s[r * 33 + lane] : tile_s[r * 33 + lane]; }
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();   // accumulator and weight buffer (wc & 1) are free again
    tc_fence_after();
  }
  wcnt += (uint32_t)n_chunks;
  if (p.out_mode == 2) {
    // combine the two column halves (ties -> lower class index), threshold, append the anchor as a sort key
    for (int item = tid; item < 128 * 9; item += SEP_THREADS) {
      const int r = item / 9, a = item - r * 9;
      const int m = P0 + r;
      if (m >= M) continue;
      float best = sD0max[r * 9 + a];
      int arg = sD0arg[r * 9 + a];
      const float v1 = sD0max[(128 + r) * 9 + a];
      const int k1 = sD0arg[(128 + r) * 9 + a];
      if (v1 > best || (v1 == best && k1 < arg)) { best = v1; arg = k1; }
      if (best > p.d0_thr) {
        const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
        const int n_a = p.d0_anchor0 + pix * 9 + a;
        const int pos = atomicAdd(p.d0_cand_count + img, 1);
        p.d0_keys[(long long)img * p.d0_cap + pos] = ((unsigned long long)(~sep_float_sortable(best)) << 32) | (unsigned)n_a;
        p.d0_cand_cls[(long long)img * p.d0_ntot + n_a] = arg;
      }
    }
  }
  if (stamp) s3_stamp(16 + step * 8 + 5);
  }   // chain steps
  if (warp == 1) tmem_dealloc(tmem_slot, 128);
}

}  // namespace hp
