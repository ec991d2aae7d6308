// tcgen05 / TMEM / TMA pointwise-conv GEMM in SPLIT PRECISION (3xTF32) for the `parity` mode: fp32 activations in
// HBM, tensor-core arithmetic, fp32-grade results.
//
//   D[M,N] = act( (A[M,K] . diag(gate[img])) * W[N,K]^T + bias ) (+ residual)          all tensors fp32
//
// The reference is fp32 end to end (efficientnet/model.py:69-104, efficientdet/model.py:42-52) and the synthetic
// network amplifies rounding ~100x, so the end-to-end bar (rel 1e-3, bit-exact kept indices) needs
// per-layer errors around 1e-5 -- out of reach for a single TF32 / fp16 product (2^-11).  Each operand is therefore
// split into two TF32 numbers, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) (22 significand bits
// together), and the product is accumulated in the fp32 TMEM accumulator as
//   A*W ~= A_lo*W_hi + A_hi*W_lo + A_hi*W_hi                                             (A_lo*W_lo ~ 2^-22 dropped)
// i.e. three tcgen05.mma.kind::tf32 per k-step.  The GEMMs of this network sit far below the tensor ridge
// (<= 165 FLOP/B), so the 3x tensor work is free; what the mode pays for is the fp32 activation traffic.
//
// CHUNKED ACCUMULATION.  The tensor core's fp32 accumulate TRUNCATES (measured, tools/tf32_probe.py: the error of a
// plain TMEM accumulation over K is a pure bias towards zero growing linearly with the number of MMAs, -7e-6 at
// K = 1152 against 6e-7 unbiased for FFMA, and the network amplifies that coherent bias to 5e-4..1.6e-3 end to end).
// So the long sum never lives in TMEM: the main term A_hi*W_hi of every CHUNK of k-blocks goes into a fresh
// accumulator of a small TMEM ring, the epilogue warps add the finished chunks into fp32 REGISTERS (round to nearest,
// like the reference's FMA chain), and the two small terms -- 2^-11 of the magnitude, their truncation is harmless --
// accumulate over the whole K in a per-tile accumulator that is added last.
//
// Pipeline per CTA (persistent, one CTA per SM, contiguous runs of tiles -- the structure of gemm_tc2_kernel):
//   warp 0       TMA producer: A tiles [128 rows x 32 fp32] (SWIZZLE_128B, OOB rows/columns zero-filled) into the
//                "hi" plane of a ring stage; W_hi / W_lo (split once on the device when the plan is built) either
//                resident per (problem, n tile) or through the ring
//   warps 10-13  split warps: landed A tile -> (x gate: `sigmoid(x_squeezed) * x`, efficientnet/model.py:93) ->
//                hi written in place, lo into the stage's second plane, same swizzled offsets
//   warp 1       one elected thread issues the 3 x (K/8) tcgen05.mma per k-block: small terms into the tile's S
//                accumulator (double-buffered across tiles), main term into the chunk ring
//   warps 2-9    epilogue: per chunk tcgen05.ld -> register accumulate; per tile + S + bias -> IEEE swish / sigmoid
//                -> smem transpose -> 128-byte coalesced fp32 rows (+ residual), or the (B, N_anchors, P) scatter
#pragma once
#include "gemm_tc.cuh"

namespace hp {

constexpr int T32_BK = 32;                          // one 128-byte swizzle row of fp32
constexpr int T32_PLANE_BYTES = TC_BM * 128;        // 16 KB: one plane (hi or lo) of an A stage
constexpr int T32_MAX_STAGES = 4;
constexpr int T32_EPI_WARPS = 8;
constexpr int T32_SPLIT_WARPS = 4;
constexpr int T32_SPLIT_THREADS = 32 * T32_SPLIT_WARPS;                       // 128
constexpr int T32_THREADS = 32 * (2 + T32_EPI_WARPS + T32_SPLIT_WARPS);       // 448
constexpr int T32_RES_MAX = 64 * 1024;              // largest resident weight panel (hi + lo)
constexpr int T32_EPI_WARP_BYTES = 32 * 33 * 4;
constexpr int T32_MAX_RING = 6;                     // chunk accumulators in flight (TMEM: 2 S buffers + the ring)

// TMEM plan for accumulators of `ncols` columns: S[2] | ring[nring]
__host__ __device__ inline int t32_ring(int ncols) { return min(T32_MAX_RING, (512 - 2 * ncols) / ncols); }

struct __align__(64) T32Prob {
  CUtensorMap tmA;    // A fp32: dims {K, M}, box {32, 128}, SWIZZLE_128B
  CUtensorMap tmBhi;  // W_hi fp32 (tf32-representable): dims {K, N}, box {32, bn}
  CUtensorMap tmBlo;  // W_lo
  GemmProb p;
};

__host__ __device__ inline int t32_smem_bytes(int stages, int b_ring_bytes, int b_res_bytes) {
  return 1024 + stages * (2 * T32_PLANE_BYTES + b_ring_bytes) + b_res_bytes + T32_EPI_WARPS * T32_EPI_WARP_BYTES +
         T32_EPI_WARPS * 128 * 4;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor: c_format F32 (1 << 4), a/b format TF32 (2) at [7,10)/[10,13), K-major A and B, N>>3 at
// [17,23), M>>4 at [24,29)  (cute::UMMA::InstrDescriptor)
__host__ __device__ inline uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// W -> (W_hi, W_lo), once per plan (weights are constants)
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i];
  const float h = tf32_rna(x);
  hi[i] = h;
  lo[i] = tf32_rna(x - h);
}

template <bool HEADOUT>
__global__ void __launch_bounds__(T32_THREADS, 1)
gemm_tf32_kernel(const T32Prob* __restrict__ probs, int nprobs, int total_tiles, int bn_max, int STAGES, int b_ring_bytes,
                 int b_res_bytes, int CH, float debias0, float debias1) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[T32_MAX_STAGES], ready_bar[T32_MAX_STAGES], empty_bar[T32_MAX_STAGES];
  __shared__ uint64_t sfull_bar[2], sempty_bar[2], mfull_bar[T32_MAX_RING], mempty_bar[T32_MAX_RING];
  __shared__ uint64_t bres_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;                                        // per stage: hi plane | lo plane
  uint8_t* sB = sA + STAGES * 2 * T32_PLANE_BYTES;           // per stage: W_hi tile | W_lo tile (ring problems)
  uint8_t* sBres = sB + STAGES * b_ring_bytes;               // resident: all k-blocks of W_hi, then of W_lo
  uint8_t* sEpi = sBres + b_res_bytes;
  float* sBias = reinterpret_cast<float*>(sEpi + T32_EPI_WARPS * T32_EPI_WARP_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ncols = 32;
  while ((int)ncols < bn_max) ncols <<= 1;
  const int NR = t32_ring((int)ncols);
  uint32_t alloc_cols = 32;
  while (alloc_cols < (2 + NR) * ncols) alloc_cols <<= 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], T32_SPLIT_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sfull_bar[i], 1); mbar_init(&sempty_bar[i], 32 * T32_EPI_WARPS); }
    for (int i = 0; i < NR; ++i) { mbar_init(&mfull_bar[i], 1); mbar_init(&mempty_bar[i], 32 * T32_EPI_WARPS); }
    mbar_init(&bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, alloc_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();   // after the TMEM allocation (see gemm_tc2_kernel)

  const int base_cnt = total_tiles / (int)gridDim.x, rem_cnt = total_tiles - base_cnt * (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * base_cnt + min((int)blockIdx.x, rem_cnt);
  const int t_end = t_begin + base_cnt + ((int)blockIdx.x < rem_cnt ? 1 : 0);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      TileCursor cur;
      uint32_t it = 0;
      int res_key = -1;
      for (int t = t_begin; t < t_end; ++t) {
        int m0, n0;
        cur.locate(probs, nprobs, t, m0, n0);
        const T32Prob* tp = probs + cur.pi;
        const int K = tp->p.K, bn = tp->p.bn;
        const int num_kb = (K + T32_BK - 1) / T32_BK;
        const bool res = tp->p.b_res != 0;
        const int wtile = bn * 128;                       // bytes of one [bn x 32] fp32 weight tile
        if (res) {
          const int key = (cur.pi << 12) | cur.nt;
          if (key != res_key) {
            res_key = key;
            for (uint32_t j = it > (uint32_t)STAGES ? it - STAGES : 0; j < it; ++j)
              mbar_wait(&empty_bar[j % STAGES], (j / STAGES) & 1, 0x3001);
            mbar_expect_tx(&bres_bar, (uint32_t)(2 * num_kb * wtile));
            for (int kb = 0; kb < num_kb; ++kb) {
              tma_load_2d(sBres + kb * wtile, &tp->tmBhi, &bres_bar, kb * T32_BK, n0);
              tma_load_2d(sBres + (num_kb + kb) * wtile, &tp->tmBlo, &bres_bar, kb * T32_BK, n0);
            }
          }
        }
        const uint32_t tx_bytes = T32_PLANE_BYTES + (res ? 0 : 2 * wtile);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1, 0x3002);
          mbar_expect_tx(&full_bar[s], tx_bytes);
          if (!res) {
            tma_load_2d(sB + s * b_ring_bytes, &tp->tmBhi, &full_bar[s], kb * T32_BK, n0);
            tma_load_2d(sB + s * b_ring_bytes + wtile, &tp->tmBlo, &full_bar[s], kb * T32_BK, n0);
          }
          if (it == 0) pdl_wait();   // weights are constants; activations only after the previous grid
          tma_load_2d(sA + s * 2 * T32_PLANE_BYTES, &tp->tmA, &full_bar[s], kb * T32_BK, m0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      TileCursor cur;
      uint32_t it = 0, i = 0, res_loads = 0, ch = 0;   // ch: running chunk count (ring position)
      int res_key = -1;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        int m0, n0;
        cur.locate(probs, nprobs, t, m0, n0);
        const GemmProb& p = probs[cur.pi].p;
        const int K = p.K, bn = p.bn;
        const bool res = p.b_res != 0;
        const int num_kb = (K + T32_BK - 1) / T32_BK;
        const int wtile = bn * 128;
        if (res) {
          const int key = (cur.pi << 12) | cur.nt;
          if (key != res_key) {
            res_key = key;
            mbar_wait(&bres_bar, res_loads & 1, 0x3003);
            ++res_loads;
          }
        }
        const uint32_t buf = i & 1;
        mbar_wait(&sempty_bar[buf], ((i >> 1) & 1) ^ 1, 0x3004);   // epilogue has drained this tile slot's S accumulator
        tc_fence_after();
        const uint32_t idesc = umma_idesc_tf32(TC_BM, bn);
        const uint32_t s_tmem = tmem_base + buf * ncols;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const int kc = kb % CH;                                   // position inside the chunk
          const uint32_t r = ch % NR;
          if (kc == 0) {
            mbar_wait(&mempty_bar[r], ((ch / NR) & 1) ^ 1, 0x3008);  // ring slot drained
            tc_fence_after();
          }
          mbar_wait(&ready_bar[s], ph, 0x3005);
          tc_fence_after();
          const uint32_t m_tmem = tmem_base + (2 + r) * ncols;
          const uint32_t ahi = smem_u32(sA + s * 2 * T32_PLANE_BYTES), alo = ahi + T32_PLANE_BYTES;
          const uint32_t bhi = res ? smem_u32(sBres + kb * wtile) : smem_u32(sB + s * b_ring_bytes);
          const uint32_t blo = res ? smem_u32(sBres + (num_kb + kb) * wtile) : bhi + wtile;
          const int krem = K - kb * T32_BK;
          const int ksteps = krem >= T32_BK ? T32_BK / 8 : (krem + 7) / 8;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t dah = umma_desc_sw128(ahi + k * 32), dal = umma_desc_sw128(alo + k * 32);
            const uint64_t dbh = umma_desc_sw128(bhi + k * 32), dbl = umma_desc_sw128(blo + k * 32);
            umma_tf32(s_tmem, dal, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);   // small terms: whole-K accumulator
            umma_tf32(s_tmem, dah, dbl, idesc, 1u);
            umma_tf32(m_tmem, dah, dbh, idesc, (kc > 0 || k > 0) ? 1u : 0u);   // main term: this chunk's accumulator
          }
          umma_commit(&empty_bar[s]);
          if (kc == CH - 1 || kb == num_kb - 1) {
            umma_commit(&mfull_bar[r]);
            ++ch;
          }
        }
        umma_commit(&sfull_bar[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= 2 + T32_EPI_WARPS) {
    // ===== split warps: hi/lo planes of every landed A tile (+ squeeze-excite gate) =====
    const int st = (warp - 2 - T32_EPI_WARPS) * 32 + lane;   // 0..127
    TileCursor cur;
    uint32_t it = 0;
    pdl_wait();   // the gate rows are written by the previous grid
    for (int t = t_begin; t < t_end; ++t) {
      int m0, n0;
      cur.locate(probs, nprobs, t, m0, n0);
      const GemmProb& p = probs[cur.pi].p;
      const int K = p.K;
      const int num_kb = (K + T32_BK - 1) / T32_BK;
      // work item = one 16-byte chunk (4 floats) of one row; thread st owns items st + 128*i, i = 0..7:
      // rows (st >> 3) + 16*i, physical chunk st & 7 -> consecutive threads touch consecutive 16-byte chunks
      const int pj = st & 7;
      const float* gate_r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (st >> 3) + 16 * i;
        const int img = min(m0 + row, p.M - 1) / p.rows_per_img;
        gate_r[i] = p.a_scale ? p.a_scale + (long long)img * K : nullptr;
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1, 0x3006);
        uint8_t* hi = sA + s * 2 * T32_PLANE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (st >> 3) + 16 * i;
          const int kbase = kb * T32_BK + ((pj ^ (row & 7)) << 2);   // logical k of this chunk (SWIZZLE_128B XOR)
          uint8_t* a = hi + row * 128 + pj * 16;
          float4 x = lds128f(a);
          if (gate_r[i] != nullptr && kbase < K) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gate_r[i] + kbase));
            x.x *= g.x; x.y *= g.y; x.z *= g.z; x.w *= g.w;
          }
          float4 h, l;
          h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
          l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y); l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
          *reinterpret_cast<float4*>(a) = h;
          *reinterpret_cast<float4*>(a + T32_PLANE_BYTES) = l;
        }
        fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready_bar[s]);
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quadrant q = warp & 3, column half h =====
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    float* tile_s = reinterpret_cast<float*>(sEpi + ew * T32_EPI_WARP_BYTES);
    float* bias_s = sBias + ew * 128;
    TileCursor cur;
    uint32_t i = 0, ch = 0;
    int bias_key = -1;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int t = t_begin; t < t_end; ++t, ++i) {
      int m0, n0;
      cur.locate(probs, nprobs, t, m0, n0);
      const GemmProb& p = probs[cur.pi].p;
      const int bn = p.bn, N = p.N, M = p.M, act = p.act;
      const uint32_t buf = i & 1;
      const int key = (cur.pi << 12) | cur.nt;
      if (key != bias_key) {
        bias_key = key;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + lane + 32 * j;
          bias_s[lane + 32 * j] = (lane + 32 * j < bn && n < N) ? __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
      const int nchunks = (bn + 31) >> 5;               // 32-column groups of this tile; this warp owns h and h + 2
      const int num_kb = (p.K + T32_BK - 1) / T32_BK;
      const int kchunks = (num_kb + CH - 1) / CH;
      float acc[2][32];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[ci][j] = 0.f;
      // main term: add every finished chunk accumulator in fp32 registers (round to nearest)
      for (int kc = 0; kc < kchunks; ++kc, ++ch) {
        const uint32_t r = ch % NR;
        mbar_wait(&mfull_bar[r], (ch / NR) & 1, 0x3009);
        tc_fence_after();
        const uint32_t m_addr = tmem_base + (2 + r) * ncols + lane_off;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = h + 2 * ci;
          if (c < nchunks) {
            uint32_t v[32];
            tmem_ld32(m_addr + (uint32_t)(c * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[ci][j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(&mempty_bar[r]);
      }
      {
        // De-bias: every MMA truncates its fp32 result towards zero (expected loss 0.72 * 2^-24 of the partial sum
        // per operation for log-uniform significands); with n MMAs per chunk the main sum comes out short by
        // (debias0 + debias1 * n) * 2^-24 on average (measured, tools/tf32_probe.py).  Add the expected loss back:
        // the error that is left is zero-mean and no longer compounds coherently through the layers.
        const int total_ksteps = (p.K + 7) >> 3;
        const float delta = (debias0 + debias1 * (float)total_ksteps / (float)kchunks) * 5.9604645e-8f;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[ci][j] = fmaf(acc[ci][j], delta, acc[ci][j]);
      }
      // small terms
      mbar_wait(&sfull_bar[buf], (i >> 1) & 1, 0x3007);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + buf * ncols + lane_off;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = h + 2 * ci;
        if (c < nchunks) {
          uint32_t v[32];
          tmem_ld32(s_addr + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[ci][j] += __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      mbar_arrive(&sempty_bar[buf]);
      if (i == 0) pdl_wait();   // before this warp's first residual read / global store
      const int mrow0 = m0 + q * 32;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = h + 2 * ci;
        if (c >= nchunks) continue;
        const int c0 = c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = apply_act<float>(acc[ci][j] + bias_s[c0 + j], act);
        __syncwarp();
        const int n = n0 + c0 + lane;
        if (c0 + lane < bn && n < N) {
          if (!HEADOUT) {
            float* outp = reinterpret_cast<float*>(p.out) + n;
            const float* resp = p.residual ? reinterpret_cast<const float*>(p.residual) + n : nullptr;
            const int ldo = p.ldo;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int m = mrow0 + r;
              if (m < M) {
                float x = tile_s[r * 33 + lane];
                if (resp) x += __ldg(resp + (long long)m * ldo);
                outp[(long long)m * ldo] = x;
              }
            }
          } else {
            // fp32 head tensors (B, N_anchors, P): scatter in the reference's permute/view order
            const int a = n / p.p_src, qq = n - a * p.p_src;
            const int coff = a * p.p_dst + p.p_off + qq;
            float* outp = reinterpret_cast<float*>(p.out);
            for (int r = 0; r < 32; ++r) {
              const int m = mrow0 + r;
              if (m < M) {
                const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
                float* dstp = outp + img * p.img_stride + (long long)pix * p.pix_stride + coff;
                *dstp = p.accumulate ? *dstp + tile_s[r * 33 + lane] : tile_s[r * 33 + lane];
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, alloc_cols);
}

}  // namespace hp
