// tcgen05 / TMEM / TMA pointwise-conv GEMM for sm_100a (fast mode: fp16 operands, fp32 accumulate).
//
//   D[M,N] = act( (A[M,K] . diag(gate[img])) * W[N,K]^T + bias ) (+ residual)
//
// A = NHWC activations (K = C_in contiguous), W = BN-folded 1x1 weights (K contiguous): both are
// K-major operands, loaded by TMA (cp.async.bulk.tensor, SWIZZLE_128B, out-of-bounds rows/columns
// zero-filled so ragged M, N and K need no padding in HBM).  One CTA computes one 128 x BN output
// tile: UMMA M=128, N=BN (multiple of 16, <= 128), K=16 per instruction, accumulator in TMEM.
//
// Warp roles (192 threads):
//   warp 0      TMA producer (one elected lane), mbarrier set-up
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  (a) squeeze-excite gate: scale the landed A tile in shared memory in place
//                   (efficientnet/model.py:93 `sigmoid(x_squeezed) * x` fused into the project conv),
//               (b) epilogue: tcgen05.ld TMEM -> registers -> bias/activation -> per-warp smem
//                   transpose -> coalesced global stores (+ residual, or the head-tensor scatter).
// Several CTAs are co-resident per SM (<= ~82 KB smem, <= 128 TMEM columns each), which overlaps one
// tile's epilogue with the next tile's loads; the kernel is HBM-bound by construction.
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "kernels_simt.cuh"

namespace hp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // one 128-byte swizzle row of fp16
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_A_STAGE_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_EPI_BYTES = 4 * 32 * 33 * 4;        // per-warp 32x33 fp32 transpose tiles

__host__ __device__ inline int tc_smem_bytes(int stages, int bn_max) {
  return 1024 /*align slack*/ + stages * (TC_A_STAGE_BYTES + bn_max * TC_BK * 2) + TC_EPI_BYTES;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16/bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32 accumulator: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format F16 (0) or
// BF16 (1) at [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ inline uint32_t umma_idesc_f16(int m, int n, int bf16) {
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ---- the kernel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const TcProb* __restrict__ probs, int nprobs, int stages,
                                                             int bn_max) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES], ready_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + stages * TC_A_STAGE_BYTES;
  const int b_stage_bytes = bn_max * TC_BK * 2;
  float* sEpi = reinterpret_cast<float*>(sB + stages * b_stage_bytes);

  int pi = 0;
  while (pi + 1 < nprobs && (int)blockIdx.x >= probs[pi + 1].p.tile_start) ++pi;
  const TcProb* tp = probs + pi;
  const GemmProb p = tp->p;
  const int tile = blockIdx.x - p.tile_start;
  const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
  const int m0 = mt * TC_BM, n0 = nt * p.bn;
  const int bn = p.bn;
  const int num_kb = (p.K + TC_BK - 1) / TC_BK;
  const bool gated = p.a_scale != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ncols = 32;
  while ((int)ncols < bn) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tp->tmA);
    tma_prefetch_desc(&tp->tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();   // only after this CTA owns its TMEM columns (see gemm_tc2_kernel)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = TC_A_STAGE_BYTES + bn * TC_BK * 2;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], tx_bytes);
        tma_load_2d(sA + s * TC_A_STAGE_BYTES, &tp->tmA, &full_bar[s], kb * TC_BK, m0);
        tma_load_2d(sB + s * b_stage_bytes, &tp->tmB, &full_bar[s], kb * TC_BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(TC_BM, bn, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(gated ? &ready_bar[s] : &full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * TC_A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * b_stage_bytes);
        const int krem = p.K - kb * TC_BK;
        const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
        for (int k = 0; k < ksteps; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                   (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(&accum_bar);       // accumulator complete
    }
    __syncwarp();
  } else {
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;      // tile row owned by this thread
    if (gated) {
      // scale A's columns by the squeeze-excite gate of the row's image, in place in the swizzled tile
      const int m = min(m0 + row, p.M - 1);
      const float* gate = p.a_scale + (long long)(m / p.rows_per_img) * p.K;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        uint8_t* rowp = sA + s * TC_A_STAGE_BYTES + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pj = (j + row) & 7;            // physical 16-byte chunk (rotated: conflict-free)
          const int kbase = kb * TC_BK + ((pj ^ (row & 7)) << 3);  // logical k of that chunk (SW128 XOR)
          if (kbase < p.K) {
            uint4 raw = *reinterpret_cast<uint4*>(rowp + pj * 16);
            __half2* h = reinterpret_cast<__half2*>(&raw);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + kbase));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + kbase + 4));
            float2 f;
            f = __half22float2(h[0]); h[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
            f = __half22float2(h[1]); h[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
            f = __half22float2(h[2]); h[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
            f = __half22float2(h[3]); h[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
            *reinterpret_cast<uint4*>(rowp + pj * 16) = raw;
          }
        }
        fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready_bar[s]);
      }
    }
    // ---- epilogue ----
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    float* tile_s = sEpi + (warp - 2) * (32 * 33);
    const int mrow0 = m0 + q * 32;
    for (int c0 = 0; c0 < bn; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        float x = __uint_as_float(v[j]);
        if (c0 + j < bn && n < p.N) x = apply_act<__half>(x + __ldg(p.bias + n), p.act);
        tile_s[lane * 33 + j] = x;
      }
      __syncwarp();
      if (p.out_mode == 0) {
        // fp16 NHWC rows: 16 lanes x half2 = 64 contiguous bytes per row, two rows per instruction
        const int cpair = (lane & 15) * 2;
        const int n = n0 + c0 + cpair;
        const bool nok = (c0 + cpair < bn) && (n < p.N);  // N and bn are even
        __half* outp = reinterpret_cast<__half*>(p.out);
        const __half* resp = reinterpret_cast<const __half*>(p.residual);
#pragma unroll 4
        for (int r = (lane >> 4); r < 32; r += 2) {
          const int m = mrow0 + r;
          if (nok && m < p.M) {
            float x0 = tile_s[r * 33 + cpair], x1 = tile_s[r * 33 + cpair + 1];
            const long long o = (long long)m * p.ldo + n;
            if (resp) {
              const float2 rr = __half22float2(*reinterpret_cast<const __half2*>(resp + o));
              x0 += rr.x; x1 += rr.y;
            }
            *reinterpret_cast<__half2*>(outp + o) = __floats2half2_rn(x0, x1);
          }
        }
      } else {
        const int n = n0 + c0 + lane;
        if (c0 + lane < bn && n < p.N) {
          const int a = n / p.p_src, qq = n - a * p.p_src;
          const int coff = a * p.p_dst + p.p_off + qq;
          float* outp = reinterpret_cast<float*>(p.out);
          for (int r = 0; r < 32; ++r) {
            const int m = mrow0 + r;
            if (m < p.M) {
              const int img = m / p.rows_per_img, pix = m - img * p.rows_per_img;
              { float* dstp = outp + img * p.img_stride + (long long)pix * p.pix_stride + coff; *dstp = p.accumulate ? *dstp + tile_s[r * 33 + lane] : tile_s[r * 33 + lane]; }
            }
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

}  // namespace hp

// =============================================================================================
// v2: persistent, warp-specialised, double-buffered TMEM accumulators.
//
// grid = min(#tiles, resident CTAs); each CTA walks tiles blockIdx.x, +gridDim.x, ... (n-tile fastest, so the
// CTAs that share an A tile run at the same time and the re-reads hit L2).  Three pipelines run
// concurrently inside a CTA: TMA -> smem ring (full/empty mbarriers, continuous across tiles),
// tcgen05.mma -> TMEM accumulator ping-pong (accf/acce mbarriers), and the epilogue of tile i overlapping
// the loads and MMAs of tile i+1.
// Roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue (two warps per TMEM lane
// quadrant, alternating 32-column chunks), warps 10..13 (gated launches only) squeeze-excite gate applied in
// place to the landed A tile.
// Epilogue (fp16 NHWC out): TMEM -> regs -> bias -> swish via tanh.approx (one MUFU) -> half2 pack ->
// per-warp padded smem tile -> 64-byte coalesced row segments to global (+ residual).  The epilogue is the
// issue-bound part of the kernel (ncu: profiles/), hence ACT is a template parameter of the chunk routine,
// shared memory is addressed through explicit ld/st.shared PTX and row pointers are strength-reduced.
// =============================================================================================
namespace hp {

constexpr int TC2_MAX_STAGES = 6;   // smem ring depth is chosen per launch (deep for K = 1152, shallow for K <= 128)
constexpr int TC2_EPI_WARPS = 8;
constexpr int TC2_THREADS = 32 * (2 + TC2_EPI_WARPS);         // 320
constexpr int TC2_GATE_WARPS = 2;                             // each gate thread scales two rows of the A tile
constexpr int TC2_THREADS_GATED = TC2_THREADS + 32 * TC2_GATE_WARPS;   // 384
constexpr int TC2_WIDE_GATE_THREADS = 32 * (TC2_EPI_WARPS + TC2_GATE_WARPS);   // 320: epilogue + gate warps
constexpr int TC2_EPI_PITCH = 80;                             // bytes per staged row: 64 + 16 pad
constexpr int TC2_EPI_WARP_BYTES = 32 * 33 * 4;               // fp32 head path: 32x33 floats per warp
constexpr int TC2_EPI_WARP_BYTES_F16 = 32 * TC2_EPI_PITCH;    // fp16 path: 32 rows x 80 bytes per warp
constexpr int TC2_RES_MAX = 48 * 1024;                        // largest weight panel kept resident in shared memory
constexpr int TC2_BIAS_BYTES = TC2_EPI_WARPS * 128 * 4;
constexpr int TC2_GATE_IMGS = 3;                               // images whose gate rows are cached per tile
constexpr int TC2_GATE_BYTES = TC2_GATE_IMGS * 1152 * 4;       // K <= 1152

__host__ __device__ inline int tc2_smem_bytes(int stages, int b_ring_bytes, int b_res_bytes, bool gated, bool headout) {
  return 1024 + stages * (TC_A_STAGE_BYTES + b_ring_bytes) + b_res_bytes +
         TC2_EPI_WARPS * (headout ? TC2_EPI_WARP_BYTES : TC2_EPI_WARP_BYTES_F16) + TC2_BIAS_BYTES +
         (gated ? TC2_GATE_BYTES : 0);
}

// optional timeline of CTA 0 (debug): globaltimer stamps at fixed points of the sepconv kernels
// (slots 0..15: sepconv3_kernel, 16..31: sepconv_kernel), read back through hmdpose_debug_read("__s3_timeline")
__device__ unsigned long long g_s3_ts[32];
__device__ __forceinline__ void s3_stamp(int i) {
  if (blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_s3_ts[i] = t;
  }
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fast-mode swish: x*sigmoid(x) = h + h*tanh(h), h = x/2  (one MUFU instead of ex2 + rcp)
__device__ __forceinline__ float swish_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
// 16-byte shared-memory accesses.  The dynamic shared buffer is aligned by adding an OFFSET to the extern array
// (never by casting through an integer), so these plain accesses keep their shared-memory provenance and compile
// to LDS.128 / STS.128 that the scheduler is free to interleave with independent work.
__device__ __forceinline__ void sts128(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ float4 lds128f(const void* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* raw) {
  const uint32_t a = smem_u32(raw);
  return raw + (((a + 1023u) & ~1023u) - a);
}

// Walks a contiguous run of tiles of a problem table (t advances by exactly one per call).  Inside a problem the
// order is n tile outer / m tile inner, so a CTA keeps the same weight panel for a long run of tiles.
struct TileCursor {
  int pi = 0, mt = 0, nt = 0, m_tiles = 1, next_start = -1;
  template <typename P>   // TcProb or T32Prob (any table entry with a GemmProb member `p`)
  __device__ __forceinline__ void locate(const P* probs, int nprobs, int t, int& m0, int& n0) {
    if (t >= next_start) {
      while (pi + 1 < nprobs && t >= probs[pi + 1].p.tile_start) ++pi;
      const GemmProb& p = probs[pi].p;
      m_tiles = p.m_tiles;
      const int local = t - p.tile_start;
      nt = local / m_tiles;
      mt = local - nt * m_tiles;
      next_start = p.tile_start + m_tiles * p.n_tiles;
    } else if (++mt == m_tiles) {
      mt = 0;
      ++nt;
    }
    m0 = mt * TC_BM;
    n0 = nt * probs[pi].p.bn;
  }
};

template <int NC> __device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v);
template <> __device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t* v) { tmem_ld32(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// NC (32 or 16) columns of the fp16 epilogue for one warp (32 rows).
//   v        accumulator values of this thread's row
//   bias_a   shared address of the NC bias values of this chunk (pre-halved for swish: t = acc/2 + bias/2)
//   stg_a    this warp's staging tile (row pitch TC2_EPI_PITCH)
//   gout     global pointer to (first row this lane writes, first column this lane writes)
//   gres     same position in the residual tensor or null
template <int NC, int ACT, bool FULL>
__device__ __forceinline__ void epi_cols_f16(const uint32_t* v, const float* bias_a, uint8_t* stg_a, int lane,
                                             __half* gout, const __half* gres, long long row_step, int rows_valid,
                                             bool cols_ok) {
  uint8_t* myrow = stg_a + lane * TC2_EPI_PITCH;
#pragma unroll
  for (int j8 = 0; j8 < NC / 8; ++j8) {
    const float4 b0 = lds128f(bias_a + j8 * 8);
    const float4 b1 = lds128f(bias_a + j8 * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = __uint_as_float(v[j8 * 8 + e]);
      if (ACT == ACT_SWISH) {
        const float t = fmaf(a, 0.5f, bb[e]);       // (acc + bias) / 2
        x[e] = fmaf(t, tanh_approx(t), t);          // x * sigmoid(x) with one MUFU
      } else if (ACT == ACT_SIGMOID) {
        x[e] = sigmoid_t<__half>(a + bb[e]);
      } else {
        x[e] = a + bb[e];
      }
    }
    uint4 pk;
    __half2* hp2 = reinterpret_cast<__half2*>(&pk);
    hp2[0] = __floats2half2_rn(x[0], x[1]); hp2[1] = __floats2half2_rn(x[2], x[3]);
    hp2[2] = __floats2half2_rn(x[4], x[5]); hp2[3] = __floats2half2_rn(x[6], x[7]);
    sts128(myrow + j8 * 16, pk);
  }
  __syncwarp();
  // write-out: NC/8 lanes cover one row's NC*2 bytes; 32/(NC/8) rows per instruction
  constexpr int LPR = NC / 8, RPI = 32 / LPR;
  const int r0 = lane / LPR;
  const uint8_t* rd = stg_a + r0 * TC2_EPI_PITCH + (lane % LPR) * 16;
#pragma unroll
  for (int rr = 0; rr < 32 / RPI; ++rr) {
    if (FULL || (cols_ok && r0 + rr * RPI < rows_valid)) {
      uint4 pk = lds128(rd + rr * RPI * TC2_EPI_PITCH);
      if (gres) {
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(gres + rr * row_step));
        __half2* a2 = reinterpret_cast<__half2*>(&pk);
        const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __half22float2(a2[e]), fr = __half22float2(r2[e]);
          a2[e] = __floats2half2_rn(fa.x + fr.x, fa.y + fr.y);
        }
      }
      *reinterpret_cast<uint4*>(gout + rr * row_step) = pk;
    }
  }
  __syncwarp();
}

// one chunk: TMEM -> registers, optional release of the accumulator buffer, epilogue maths + stores
template <int NC>
__device__ __forceinline__ void epi_run_chunk(uint32_t taddr, uint64_t* release_bar, int act, const float* bias_a,
                                              uint8_t* stg_a, int lane, __half* out_base, const __half* res_base,
                                              int ldo, int rows_valid, int cols_valid) {
  uint32_t v[NC];
  tmem_ld_cols<NC>(taddr, v);
  if (release_bar) {   // last TMEM read of this warp for this tile
    tc_fence_before();
    mbar_arrive(release_bar);
  }
  constexpr int LPR = NC / 8, RPI = 32 / LPR;
  const long long row_step = (long long)RPI * ldo;
  const long long o0 = (long long)(lane / LPR) * ldo + (lane % LPR) * 8;
  __half* gout = out_base + o0;
  const __half* gres = res_base ? res_base + o0 : nullptr;
  const bool cols_ok = (lane % LPR) * 8 < cols_valid;
  const bool full = rows_valid >= 32 && cols_valid >= NC;   // warp-uniform
  if (act == ACT_SWISH) {
    if (full) epi_cols_f16<NC, ACT_SWISH, true>(v, bias_a, stg_a, lane, gout, gres, row_step, rows_valid, cols_ok);
    else epi_cols_f16<NC, ACT_SWISH, false>(v, bias_a, stg_a, lane, gout, gres, row_step, rows_valid, cols_ok);
  } else if (act == ACT_NONE) {
    if (full) epi_cols_f16<NC, ACT_NONE, true>(v, bias_a, stg_a, lane, gout, gres, row_step, rows_valid, cols_ok);
    else epi_cols_f16<NC, ACT_NONE, false>(v, bias_a, stg_a, lane, gout, gres, row_step, rows_valid, cols_ok);
  } else {
    epi_cols_f16<NC, ACT_SIGMOID, false>(v, bias_a, stg_a, lane, gout, gres, row_step, rows_valid, cols_ok);
  }
}

// Squeeze-excite gate of ONE tile applied by TC2_WIDE_GATE_THREADS threads (gid = 0..319): the gate rows of the
// images the tile touches are cached in shared memory, then every k-block that lands is scaled in place
// (`sigmoid(x_squeezed) * x`, efficientnet/model.py:93) and handed to the MMA warp through ready_bar.
// Work item = one 16-byte chunk of one row: consecutive threads take consecutive chunks (conflict-free).
// (noinline: the epilogue warps and the gate warps reach the named barrier below through ONE instruction, which is also
// what compute-sanitizer's synccheck expects of a barrier)
__device__ __noinline__ void gate_tile_wide(const GemmProb& p, int m0, uint8_t* sA, float* sGate, uint64_t* full_bar,
                                               uint64_t* ready_bar, int stages, int gid) {
  if (p.a_scale == nullptr) return;
  const int K = p.K;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int img0 = m0 / p.rows_per_img;
  const int img1 = min(m0 + TC_BM - 1, p.M - 1) / p.rows_per_img;
  const int nimg = img1 - img0 + 1;
  const bool cached = nimg <= TC2_GATE_IMGS && K <= 1152;
  if (cached) {
    const float4* src = reinterpret_cast<const float4*>(p.a_scale + (long long)img0 * K);
    float4* dst = reinterpret_cast<float4*>(sGate);
    const int n4 = nimg * K / 4;
    for (int i4 = gid; i4 < n4; i4 += TC2_WIDE_GATE_THREADS) dst[i4] = __ldg(src + i4);
  }
  __syncwarp();   // the fill loop has a per-lane trip count: reconverge before the (warp-aligned) named barrier
  asm volatile("bar.sync 1, %0;" ::"n"(TC2_WIDE_GATE_THREADS) : "memory");
  constexpr int ITEMS = (TC_BM * 8 + TC2_WIDE_GATE_THREADS - 1) / TC2_WIDE_GATE_THREADS;   // 4
  const float* gate_r[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int row = min((gid + TC2_WIDE_GATE_THREADS * i) >> 3, TC_BM - 1);
    const int my_img = min(m0 + row, p.M - 1) / p.rows_per_img;
    gate_r[i] = cached ? sGate + (my_img - img0) * K : p.a_scale + (long long)my_img * K;
  }
  for (int kb = 0; kb < num_kb; ++kb) {   // single tile per CTA: ring position == k-block index
    const int s = kb % stages;
    mbar_wait(&full_bar[s], (kb / stages) & 1);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int item = gid + TC2_WIDE_GATE_THREADS * i;
      if (item < TC_BM * 8) {
        const int row = item >> 3, pj = item & 7;
        const int kbase = kb * TC_BK + ((pj ^ (row & 7)) << 3);
        if (kbase < K) {
          uint8_t* a = sA + s * TC_A_STAGE_BYTES + item * 16;
          uint4 raw = lds128(a);
          __half2* h = reinterpret_cast<__half2*>(&raw);
          const float4 g0 = *reinterpret_cast<const float4*>(gate_r[i] + kbase);
          const float4 g1 = *reinterpret_cast<const float4*>(gate_r[i] + kbase + 4);
          float2 f;
          f = __half22float2(h[0]); h[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
          f = __half22float2(h[1]); h[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
          f = __half22float2(h[2]); h[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
          f = __half22float2(h[3]); h[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
          sts128(a, raw);
        }
      }
    }
    fence_async_smem();
    mbar_arrive(&ready_bar[s]);
  }
}

// Shared-memory layout (after 1 KB alignment): A ring | B ring (launches with a non-resident problem) | resident
// weight panel | epilogue staging | bias | gate cache.  A problem is RESIDENT (GemmProb::b_res) when the whole
// [bn x K] weight panel of an n tile fits TC2_RES_MAX bytes: the producer then loads it once per (problem, n tile)
// run instead of once per tile -- the per-tile reload of the same few KB by every CTA serialises on a handful of
// L2 sectors and was the bound of the shallow-K expand convolutions (profiles/r1_gemm_diag.txt).
template <bool GATED, bool HEADOUT>
__global__ void __launch_bounds__(GATED ? TC2_THREADS_GATED : TC2_THREADS, 2)
gemm_tc2_kernel(const TcProb* __restrict__ probs, int nprobs, int total_tiles, int bn_max, int TC2_STAGES,
                int b_ring_bytes, int b_res_bytes) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC2_MAX_STAGES], ready_bar[TC2_MAX_STAGES], empty_bar[TC2_MAX_STAGES], accf_bar[2], acce_bar[2];
  __shared__ uint64_t bres_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + TC2_STAGES * TC_A_STAGE_BYTES;
  uint8_t* sBres = sB + TC2_STAGES * b_ring_bytes;
  uint8_t* sEpi = sBres + b_res_bytes;
  float* sBias = reinterpret_cast<float*>(sEpi + TC2_EPI_WARPS * (HEADOUT ? TC2_EPI_WARP_BYTES : TC2_EPI_WARP_BYTES_F16));
  float* sGate = sBias + TC2_EPI_WARPS * 128;   // only allocated for gated launches

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ncols = 32;
  while ((int)ncols < bn_max) ncols <<= 1;   // columns per accumulator buffer

  // gated launches whose CTAs own a single tile (the deep-K project convolutions of the small feature maps): the eight
  // epilogue warps are idle during the k loop, so all 320 non-producer threads apply the squeeze-excite gate
  const bool wide_gate = GATED && total_tiles <= (int)gridDim.x;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], wide_gate ? TC2_WIDE_GATE_THREADS : 32 * TC2_GATE_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&accf_bar[i], 1); mbar_init(&acce_bar[i], 32 * TC2_EPI_WARPS); }
    mbar_init(&bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 2 * ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // The dependents of this grid may only become resident once EVERY CTA of it owns its TMEM columns: tcgen05.alloc
  // blocks while the SM's 512 columns are taken, and a dependent CTA that got its (smaller) allocation first would
  // then spin on this grid's completion while holding the columns this grid is waiting for -- with several streams
  // of PDL chains in flight that is a cross-stream deadlock (the watchdog trap of round 1's 8-GPU run).  Invariant:
  // "resident and triggered" implies "holds every resource it will ever need".
  pdl_trigger();
  // Programmatic dependent launch: everything above touches only this CTA's shared memory / TMEM.  Weights, bias and
  // the problem table are constants: the producer fetches the first weight tile BEFORE waiting for the previous grid;
  // every role executes pdl_wait before its first access to activations (A tiles, gate, residual, output).

  // contiguous run of tiles of this CTA
  const int base_cnt = total_tiles / (int)gridDim.x, rem_cnt = total_tiles - base_cnt * (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * base_cnt + min((int)blockIdx.x, rem_cnt);
  const int t_end = t_begin + base_cnt + ((int)blockIdx.x < rem_cnt ? 1 : 0);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      TileCursor cur;
      uint32_t it = 0;
      int res_key = -1;
      // per-image weight panels are written by the previous kernel (se3_kernel): no weight fetch ahead of it
      if (probs[0].p.w_img_rows > 0) pdl_wait();
      for (int t = t_begin; t < t_end; ++t) {
        int m0, n0;
        cur.locate(probs, nprobs, t, m0, n0);
        const TcProb* tp = probs + cur.pi;
        const int K = tp->p.K, bn = tp->p.bn;
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        const bool res = tp->p.b_res != 0;
        const int img = tp->p.w_img_rows > 0 ? m0 / tp->p.rows_per_img : 0;
        const int wrow = n0 + img * tp->p.w_img_rows;   // row of this tile's weight panel in the (per-image) weight matrix
        if (res) {
          const int key = (cur.pi << 24) | (img << 8) | cur.nt;
          if (key != res_key) {
            res_key = key;
            // every MMA that read the previous panel has completed once all issued stages have been released
            for (uint32_t j = it > (uint32_t)TC2_STAGES ? it - TC2_STAGES : 0; j < it; ++j)
              mbar_wait(&empty_bar[j % TC2_STAGES], (j / TC2_STAGES) & 1);
            mbar_expect_tx(&bres_bar, (uint32_t)(num_kb * bn * TC_BK * 2));
            for (int kb = 0; kb < num_kb; ++kb)
              tma_load_2d(sBres + kb * bn * TC_BK * 2, &tp->tmB, &bres_bar, kb * TC_BK, wrow);
          }
        }
        const uint32_t tx_bytes = TC_A_STAGE_BYTES + (res ? 0 : bn * TC_BK * 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % TC2_STAGES;
          const uint32_t ph = (it / TC2_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], tx_bytes);
          if (!res) tma_load_2d(sB + s * b_ring_bytes, &tp->tmB, &full_bar[s], kb * TC_BK, wrow);
          if (it == 0) pdl_wait();   // the first weight tile is in flight; activations only after the previous grid
          tma_load_2d(sA + s * TC_A_STAGE_BYTES, &tp->tmA, &full_bar[s], kb * TC_BK, m0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      TileCursor cur;
      uint32_t it = 0, i = 0, res_loads = 0;
      int res_key = -1;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        int m0, n0;
        cur.locate(probs, nprobs, t, m0, n0);
        const GemmProb& p = probs[cur.pi].p;
        const int K = p.K, bn = p.bn;
        const bool gated = p.a_scale != nullptr;
        const bool res = p.b_res != 0;
        if (res) {
          const int img = p.w_img_rows > 0 ? m0 / p.rows_per_img : 0;
          const int key = (cur.pi << 24) | (img << 8) | cur.nt;
          if (key != res_key) {
            res_key = key;
            mbar_wait(&bres_bar, res_loads & 1);
            ++res_loads;
          }
        }
        const uint32_t buf = i & 1;
        mbar_wait(&acce_bar[buf], ((i >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(TC_BM, bn, 0);
        const uint32_t d_tmem = tmem_base + buf * ncols;
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % TC2_STAGES;
          const uint32_t ph = (it / TC2_STAGES) & 1;
          mbar_wait(gated ? &ready_bar[s] : &full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * TC_A_STAGE_BYTES);
          const uint32_t b_addr = res ? smem_u32(sBres + kb * bn * TC_BK * 2) : smem_u32(sB + s * b_ring_bytes);
          const int krem = K - kb * TC_BK;
          const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&accf_bar[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= 2 + TC2_EPI_WARPS) {
    // ===== squeeze-excite gate warps (gated launches only): thread t scales rows t and t + 64 =====
    const int gt = (warp - 2 - TC2_EPI_WARPS) * 32 + lane;   // 0..63
    TileCursor cur;
    uint32_t it = 0;
    pdl_wait();   // the gate rows are written by the previous grid
    if (wide_gate) {
      if (t_begin < t_end) {
        int m0, n0;
        cur.locate(probs, nprobs, t_begin, m0, n0);
        gate_tile_wide(probs[cur.pi].p, m0, sA, sGate, full_bar, ready_bar, TC2_STAGES, 32 * TC2_EPI_WARPS + gt);
      }
    } else
    for (int t = t_begin; t < t_end; ++t) {
      int m0, n0;
      cur.locate(probs, nprobs, t, m0, n0);
      const GemmProb& p = probs[cur.pi].p;
      const int K = p.K;
      const int num_kb = (K + TC_BK - 1) / TC_BK;
      if (p.a_scale == nullptr) { it += num_kb; continue; }
      // gate rows of the images this tile touches -> shared memory once per tile (the k loop then never waits
      // on global memory; the in-place scaling is the serial stage between TMA and MMA)
      const int img0 = m0 / p.rows_per_img;
      const int img1 = min(m0 + TC_BM - 1, p.M - 1) / p.rows_per_img;
      const int nimg = img1 - img0 + 1;
      const bool cached = nimg <= TC2_GATE_IMGS && K <= 1152;
      __syncwarp();
      asm volatile("bar.sync 1, 64;" ::: "memory");   // previous tile's readers are done with sGate
      if (cached) {
        const float4* src = reinterpret_cast<const float4*>(p.a_scale + (long long)img0 * K);
        float4* dst = reinterpret_cast<float4*>(sGate);
        const int n4 = nimg * K / 4;
        for (int i4 = gt; i4 < n4; i4 += 64) dst[i4] = __ldg(src + i4);
      }
      __syncwarp();   // per-lane trip count above: reconverge before the (warp-aligned) named barrier
      asm volatile("bar.sync 1, 64;" ::: "memory");
      const float* gate_r[2];
      int rows[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        rows[hh] = gt + 64 * hh;
        const int my_img = min(m0 + rows[hh], p.M - 1) / p.rows_per_img;
        gate_r[hh] = cached ? sGate + (my_img - img0) * K : p.a_scale + (long long)my_img * K;
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % TC2_STAGES;
        const uint32_t ph = (it / TC2_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = rows[hh];
          const float* gate = gate_r[hh];
          uint8_t* rowa = sA + s * TC_A_STAGE_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int pj = (j + row) & 7;
            const int kbase = kb * TC_BK + ((pj ^ (row & 7)) << 3);
            if (kbase < K) {
              uint4 raw = lds128(rowa + pj * 16);
              __half2* h = reinterpret_cast<__half2*>(&raw);
              const float4 g0 = *reinterpret_cast<const float4*>(gate + kbase);
              const float4 g1 = *reinterpret_cast<const float4*>(gate + kbase + 4);
              float2 f;
              f = __half22float2(h[0]); h[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
              f = __half22float2(h[1]); h[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
              f = __half22float2(h[2]); h[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
              f = __half22float2(h[3]); h[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
              sts128(rowa + pj * 16, raw);
            }
          }
        }
        fence_async_smem();
        mbar_arrive(&ready_bar[s]);
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quadrant q = warp & 3, column half h =====
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    uint8_t* stg_a = sEpi + ew * (HEADOUT ? TC2_EPI_WARP_BYTES : TC2_EPI_WARP_BYTES_F16);
    float* bias_s = sBias + ew * 128;
    TileCursor cur;
    uint32_t i = 0;
    int bias_key = -1;
    for (int t = t_begin; t < t_end; ++t, ++i) {
      int m0, n0;
      cur.locate(probs, nprobs, t, m0, n0);
      const GemmProb& p = probs[cur.pi].p;
      const int bn = p.bn, N = p.N, M = p.M, act = p.act;
      const uint32_t buf = i & 1;
      const int key = (cur.pi << 12) | cur.nt;
      if (key != bias_key) {   // bias of this (problem, n tile): once per run of tiles
        bias_key = key;
        const float sc = (!HEADOUT && act == ACT_SWISH) ? 0.5f : 1.0f;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + lane + 32 * j;
          bias_s[lane + 32 * j] = (lane + 32 * j < bn && n < N) ? sc * __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
      if (i == 0) pdl_wait();   // before this warp's first residual read / global store
      if (GATED && wide_gate) gate_tile_wide(p, m0, sA, sGate, full_bar, ready_bar, TC2_STAGES, ew * 32 + lane);
      mbar_wait(&accf_bar[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + buf * ncols + ((uint32_t)(q * 32) << 16);
      const int mrow0 = m0 + q * 32;
      if (!HEADOUT) {
        // columns in units of 16: half 0 takes ceil(units / 2), half 1 the rest
        const int units = bn >> 4;
        const int u0 = h == 0 ? 0 : (units + 1) >> 1;
        const int u1 = h == 0 ? (units + 1) >> 1 : units;
        const int ldo = p.ldo;
        const int rows_valid = M - mrow0;
        __half* orow = reinterpret_cast<__half*>(p.out) + (long long)mrow0 * ldo + n0;
        const __half* rrow = p.residual ? reinterpret_cast<const __half*>(p.residual) + (long long)mrow0 * ldo + n0 : nullptr;
        if (u0 >= u1) {   // this warp has no columns in this tile (bn == 16 and h == 1)
          tc_fence_before();
          mbar_arrive(&acce_bar[buf]);
        }
        for (int u = u0; u < u1;) {
          const int c0 = u * 16;
          const int cols_valid = min(N - n0, bn) - c0;   // may be <= 0 for a ragged last n tile
          if (u + 2 <= u1) {
            epi_run_chunk<32>(t_addr + (uint32_t)c0, u + 2 == u1 ? &acce_bar[buf] : nullptr, act, bias_s + c0, stg_a, lane,
                              orow + c0, rrow ? rrow + c0 : nullptr, ldo, rows_valid, cols_valid);
            u += 2;
          } else {
            epi_run_chunk<16>(t_addr + (uint32_t)c0, &acce_bar[buf], act, bias_s + c0, stg_a, lane, orow + c0,
                              rrow ? rrow + c0 : nullptr, ldo, rows_valid, cols_valid);
            u += 1;
          }
        }
      } else {
        // fp32 head tensors (B, N_anchors, P): scatter in the reference's permute/view order
        const int nchunks = (bn + 31) >> 5;
        bool released = false;
        float* tile_s = reinterpret_cast<float*>(stg_a);
        float* outp = reinterpret_cast<float*>(p.out);
        for (int c = h; c < nchunks; c += 2) {
          const int c0 = c * 32;
          uint32_t v[32];
          tmem_ld32(t_addr + (uint32_t)c0, v);
          if (c + 2 >= nchunks) {
            tc_fence_before();
            mbar_arrive(&acce_bar[buf]);
            released = true;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]) + bias_s[c0 + j];
            tile_s[lane * 33 + j] = apply_act<__half>(x, act);
          }
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
        if (!released) {   // this warp had no chunk in this tile (bn <= 32 and h == 1)
          tc_fence_before();
          mbar_arrive(&acce_bar[buf]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * ncols);
}

}  // namespace hp
