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
constexpr int T32_EPI_WARPS = 16;                    // lane quadrant q = warp & 3, 32-column group h = 0..3
constexpr int T32_SPLIT_WARPS = 4;
constexpr int T32_SPLIT_THREADS = 32 * T32_SPLIT_WARPS;                       // 128
constexpr int T32_THREADS = 32 * (2 + T32_EPI_WARPS + T32_SPLIT_WARPS);       // 704
constexpr int T32_RES_MAX = 64 * 1024;              // largest resident weight panel (hi + lo)
constexpr int T32_EPI_PITCH = 20;                     // floats per transposed row of 16 columns: 16-byte aligned, conflict-free float4 writes
constexpr int T32_EPI_WARP_BYTES = 32 * T32_EPI_PITCH * 4;
constexpr int T32_GATE_IMGS = 3;                      // images one 128-row tile may touch with its gate rows cached
constexpr int T32_GATE_BYTES = T32_GATE_IMGS * 1152 * 4;
constexpr int T32_MAX_RING = 6;                     // chunk accumulators in flight (TMEM: 2 S buffers + the ring)

// TMEM plan for accumulators of `ncols` columns: S[2] | ring[nring]
__host__ __device__ inline int t32_ring(int ncols) { return min(T32_MAX_RING, (512 - 2 * ncols) / ncols); }

struct __align__(64) T32Prob {
  CUtensorMap tmA;    // A fp32: dims {K, M}, box {32, 128}, SWIZZLE_128B
  CUtensorMap tmBhi;  // W_hi fp32 (tf32-representable): dims {K, N}, box {32, bn}
  CUtensorMap tmBlo;  // W_lo
  GemmProb p;
};

__host__ __device__ inline int t32_smem_bytes(int stages, int b_ring_bytes, int b_res_bytes, bool gated) {
  return 1024 + stages * (2 * T32_PLANE_BYTES + b_ring_bytes) + b_res_bytes + T32_EPI_WARPS * T32_EPI_WARP_BYTES +
         T32_EPI_WARPS * 32 * 4 + (gated ? T32_GATE_BYTES : 0);
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

// Epilogue activation of the parity GEMM.  swish = x / (1 + e^-x) with the accurate expf and ONE approximate division
// (<= 2 ulp, no slow-path branches): the IEEE `x * (1 / (1 + expf(-x)))` spent more issue slots in the division's
// fix-up code than the whole rest of the epilogue (ncu: 12 M BRA of 80 M warp instructions on the 262144 x 96 x 16
// expand).  The class scores keep the IEEE sigmoid.
__device__ __forceinline__ float t32_act(float x, int act) {
  if (act == ACT_SWISH) return __fdividef(x, 1.0f + expf(-x));
  if (act == ACT_SIGMOID) return sigmoid_t<float>(x);
  return x;
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
  float* sBias = reinterpret_cast<float*>(sEpi + T32_EPI_WARPS * T32_EPI_WARP_BYTES);   // [warp][32]
  float* sGate = sBias + T32_EPI_WARPS * 32;   // gate rows of the current tile's images (gated launches only)

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
      if (p.a_scale != nullptr) {
        // gate rows of the images this tile touches -> shared memory once per tile: the k loop never waits on a
        // dependent global load between the TMA landing and the MMA (it cost ~1 us per k-block of the deep-K projects)
        const int img0 = m0 / p.rows_per_img;
        const int img1 = min(m0 + TC_BM - 1, p.M - 1) / p.rows_per_img;
        const int nimg = img1 - img0 + 1;
        const bool cached = nimg <= T32_GATE_IMGS && K <= 1152;
        __syncwarp();
        asm volatile("bar.sync 2, %0;" ::"n"(T32_SPLIT_THREADS) : "memory");   // previous tile's readers are done with sGate
        if (cached) {
          const float4* src = reinterpret_cast<const float4*>(p.a_scale + (long long)img0 * K);
          float4* dst = reinterpret_cast<float4*>(sGate);
          for (int i4 = st; i4 < nimg * K / 4; i4 += T32_SPLIT_THREADS) dst[i4] = __ldg(src + i4);
        }
        __syncwarp();
        asm volatile("bar.sync 2, %0;" ::"n"(T32_SPLIT_THREADS) : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (st >> 3) + 16 * i;
          const int img = min(m0 + row, p.M - 1) / p.rows_per_img;
          gate_r[i] = cached ? sGate + (img - img0) * K : p.a_scale + (long long)img * K;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) gate_r[i] = nullptr;
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
            const float4 g = *reinterpret_cast<const float4*>(gate_r[i] + kbase);
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
    // ===== epilogue warps 2..17: TMEM lane quadrant q = warp & 3, 32-column group h = 0..3 =====
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    float* tile_s = reinterpret_cast<float*>(sEpi + ew * T32_EPI_WARP_BYTES);
    float* bias_s = sBias + ew * 32;
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
      const int c0 = h * 32;
      const bool mine = c0 < bn;                        // this warp owns a column group of this tile (warp-uniform)
      if (key != bias_key) {
        bias_key = key;
        __syncwarp();
        const int n = n0 + c0 + lane;
        bias_s[lane] = (c0 + lane < bn && n < N) ? __ldg(p.bias + n) : 0.f;
        __syncwarp();
      }
      const int num_kb = (p.K + T32_BK - 1) / T32_BK;
      const int kchunks = (num_kb + CH - 1) / CH;
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      // main term: add every finished chunk accumulator in fp32 registers (round to nearest)
      for (int kc = 0; kc < kchunks; ++kc, ++ch) {
        const uint32_t r = ch % NR;
        mbar_wait(&mfull_bar[r], (ch / NR) & 1, 0x3009);
        tc_fence_after();
        if (mine) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (2 + r) * ncols + lane_off + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
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
        for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], delta, acc[j]);
      }
      // small terms
      mbar_wait(&sfull_bar[buf], (i >> 1) & 1, 0x3007);
      tc_fence_after();
      if (mine) {
        uint32_t v[32];
        tmem_ld32(tmem_base + buf * ncols + lane_off + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(&sempty_bar[buf]);
      if (i == 0) pdl_wait();   // before this warp's first residual read / global store
      if (!mine) continue;
      const int mrow0 = m0 + q * 32;
      // bias + activation in registers
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j4 * 4);
        acc[j4 * 4] = t32_act(acc[j4 * 4] + b4.x, act); acc[j4 * 4 + 1] = t32_act(acc[j4 * 4 + 1] + b4.y, act);
        acc[j4 * 4 + 2] = t32_act(acc[j4 * 4 + 2] + b4.z, act); acc[j4 * 4 + 3] = t32_act(acc[j4 * 4 + 3] + b4.w, act);
      }
      if (!HEADOUT && (N & 3) == 0 && (p.ldo & 3) == 0) {
        // transpose through shared memory 16 columns at a time with 16-byte accesses: thread = row writes 4 float4,
        // then 4 lanes read one row -> 8 rows x 64 contiguous bytes per warp store instruction
        const int cq = (lane & 3) * 4;
        const int ldo = p.ldo;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(tile_s + lane * T32_EPI_PITCH + j4 * 4) =
                make_float4(acc[half * 16 + j4 * 4], acc[half * 16 + j4 * 4 + 1], acc[half * 16 + j4 * 4 + 2], acc[half * 16 + j4 * 4 + 3]);
          __syncwarp();
          const int n = n0 + c0 + half * 16 + cq;
          const bool col_ok = c0 + half * 16 + cq < bn && n < N;
          float* outp = reinterpret_cast<float*>(p.out) + n;
          const float* resp = p.residual ? reinterpret_cast<const float*>(p.residual) + n : nullptr;
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 8) {
            const int r = r0 + (lane >> 2);
            const int m = mrow0 + r;
            if (col_ok && m < M) {
              float4 x = *reinterpret_cast<const float4*>(tile_s + r * T32_EPI_PITCH + cq);
              if (resp) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(resp + (long long)m * ldo));
                x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
              }
              *reinterpret_cast<float4*>(outp + (long long)m * ldo) = x;
            }
          }
        }
      } else {
        // ragged N / head tensors: scalar path, 16 columns per pass (lanes 0..15 = columns, two rows per instruction)
        const int cl = lane & 15, rsel = lane >> 4;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(tile_s + lane * T32_EPI_PITCH + j4 * 4) =
                make_float4(acc[half * 16 + j4 * 4], acc[half * 16 + j4 * 4 + 1], acc[half * 16 + j4 * 4 + 2], acc[half * 16 + j4 * 4 + 3]);
          __syncwarp();
          const int n = n0 + c0 + half * 16 + cl;
          if (c0 + half * 16 + cl < bn && n < N) {
            if (!HEADOUT) {
              float* outp = reinterpret_cast<float*>(p.out) + n;
              const float* resp = p.residual ? reinterpret_cast<const float*>(p.residual) + n : nullptr;
              const int ldo = p.ldo;
#pragma unroll 4
              for (int r = rsel; r < 32; r += 2) {
                const int m = mrow0 + r;
                if (m < M) {
                  float x = tile_s[r * T32_EPI_PITCH + cl];
                  if (resp) x += __ldg(resp + (long long)m * ldo);
                  outp[(long long)m * ldo] = x;
                }
              }
            } else {
              // fp32 head tensors (B, N_anchors, P): scatter in the reference's permute/view order
              const int a = n / p.p_src, qq = n - a * p.p_src;
              const int coff = a * p.p_dst + p.p_off + qq;
              float* outp = reinterpret_cast<float*>(p.out);
              for (int r = rsel; r < 32; r += 2) {
                const int m = mrow0 + r;
                if (m < M) {
                  const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
                  float* dstp = outp + img * p.img_stride + (long long)pix * p.pix_stride + coff;
                  *dstp = p.accumulate ? *dstp + tile_s[r * T32_EPI_PITCH + cl] : tile_s[r * T32_EPI_PITCH + cl];
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, alloc_cols);
}

}  // namespace hp
