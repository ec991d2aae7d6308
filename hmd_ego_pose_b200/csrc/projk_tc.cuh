// Split-K project GEMM of the small feature maps (fast mode, fp16): the squeeze-excite gate multiply + project 1x1
// conv + BN (+ identity skip) of efficientnet/model.py:93-104 for the deep-K blocks (K = 480 ... 1152, 8x8 / 16x16 maps).
//
// Why a second GEMM kernel: gemm_tc2_kernel gives one CTA a whole [128 x K] row panel, and ONE SM ingests its K-major
// operands at ~0.75 us per k-block (128 + bn rows of 128 bytes, measured with and without the gate, at batch 1 and at
// batch 16 alike) -- 10-20 us per launch for 2-7 MB, with 2 ... 32 CTAs busy.  Here a CLUSTER of S <= 6 CTAs shares an
// m tile and splits K: CTA `rank` takes k-blocks rank, rank + S, ... (<= 3), so every SM streams a third to a sixth of
// the panel and all loads of a CTA are in flight at once.
//   issue lane : TMA of my W_proj k-blocks (constants: before griddepcontrol.wait), TMA of my A k-blocks, then -- once
//                the workers have gated the A tiles -- all tcgen05.mma (M = 128, N = cout <= 320) into one TMEM tile
//   workers    : gate `sigmoid(x_squeezed) * x` applied in place to the landed A k-blocks (gate rows of the <= 3 images a
//                tile touches cached in shared memory); fp32 partial tile TMEM -> per-warp transpose -> L2-resident
//                scratch with full-line stores; ONE cluster barrier; every CTA then sums the S partials of its own
//                128 / S rows in rank order (bitwise reproducible), + bias (+ skip) -> fp16 -> global
#pragma once
#include "mbconv_tc.cuh"

namespace hp {

__device__ __forceinline__ void pk_workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PK_WORKERS) : "memory"); }

__global__ void __launch_bounds__(PK_THREADS, 1) projk_kernel(const PkSpec sp) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_w, bar_a, bar_g, bar_d;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + sp.off_w;
  float* sGate = reinterpret_cast<float*>(smem + sp.off_gate);   // [3 images][PK_MAX_MINE][64]
  uint8_t* sStage = smem;                                        // aliases the operands once the MMAs are done

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = sp.S;
  const int rank = (int)cluster.block_rank();
  const int mt = blockIdx.x / S;
  const int m0 = mt * 128;
  const int N = sp.N, K = sp.K;
  const int nmine = rank < sp.nkb ? (sp.nkb - rank + S - 1) / S : 0;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_a, 1);
    mbar_init(&bar_g, PK_WORKERS / 32);
    mbar_init(&bar_d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PK_WORKERS / 32) tmem_alloc(&tmem_slot, (uint32_t)sp.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();   // this CTA holds everything it will ever need (shared memory, TMEM columns)

  const int nsplit = N > 256 ? 2 : 1, nn = N / nsplit;
  if (warp == PK_WORKERS / 32) {
    // ===================== issue warp =====================
    if (lane == 0 && nmine > 0) {
      tma_prefetch_desc(sp.tm + 0);
      tma_prefetch_desc(sp.tm + 1);
      mbar_expect_tx(&bar_w, (uint32_t)(nmine * sp.w_slice_bytes));
      for (int li = 0; li < nmine; ++li)
        for (int h = 0; h < nsplit; ++h)
          tma_load_2d(sW + li * sp.w_slice_bytes + h * nn * 128, sp.tm + 1, &bar_w, (rank + li * S) * 64, h * nn);
      pdl_wait();
      mbar_expect_tx(&bar_a, (uint32_t)(nmine * 16384));
      for (int li = 0; li < nmine; ++li) tma_load_2d(sA + li * 16384, sp.tm + 0, &bar_a, (rank + li * S) * 64, m0);
      mbar_wait(&bar_w, 0, 0x6001);
      mbar_wait(&bar_g, 0, 0x6002);   // A k-blocks landed and gated
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(128, nn, 0);
      for (int li = 0; li < nmine; ++li) {
        const int krem = K - (rank + li * S) * 64;
        const int ksteps = krem >= 64 ? 4 : (krem + 15) / 16;
        for (int h = 0; h < nsplit; ++h) {
          const uint32_t a = smem_u32(sA + li * 16384), b = smem_u32(sW + li * sp.w_slice_bytes + h * nn * 128);
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16(tmem_base + h * nn, umma_desc_sw128(a + kk * 32), umma_desc_sw128(b + kk * 32), idesc,
                     (li > 0 || kk > 0) ? 1u : 0u);
        }
      }
      umma_commit(&bar_d);
    }
    __syncwarp();
    cluster_arrive(); cluster_wait();   // partial tiles barrier
  } else {
    // ===================== worker warps =====================
    pdl_wait();   // gate rows, A, residual and the scratch belong to the previous kernels until here
    // gate rows of the images this tile touches, my channels only -> shared memory
    const int img0 = m0 / sp.rows_per_img;
    const int img1 = min(m0 + 127, sp.M - 1) / sp.rows_per_img;
    const int nimg = min(img1 - img0 + 1, 3);
    for (int i = tid; i < nimg * nmine * 64; i += PK_WORKERS) {
      const int im = i / (nmine * 64), r = i - im * (nmine * 64);
      const int li = r >> 6, c = (rank + li * S) * 64 + (r & 63);
      sGate[(im * PK_MAX_MINE + li) * 64 + (r & 63)] = c < K ? __ldcg(sp.gate + (size_t)(img0 + im) * K + c) : 0.f;
    }
    pk_workers_sync();
    if (nmine > 0) {
      mbar_wait(&bar_a, 0, 0x6010);
      // `sigmoid(x_squeezed) * x` (model.py:93) in place on the swizzled A k-blocks: item = 16-byte chunk of a row
      for (int idx = tid; idx < nmine * 1024; idx += PK_WORKERS) {
        const int li = idx >> 10, item = idx & 1023;
        const int row = item >> 3, pj = item & 7;
        const int kc = (pj ^ (row & 7)) << 3;   // channel (within the k-block) of this physical chunk
        const int im = min(min(m0 + row, sp.M - 1) / sp.rows_per_img - img0, 2);
        uint4* ptr = reinterpret_cast<uint4*>(sA + li * 16384 + item * 16);
        uint4 raw = *ptr;
        __half2* h = reinterpret_cast<__half2*>(&raw);
        const float* gp = sGate + (im * PK_MAX_MINE + li) * 64 + kc;
        const float4 g0 = lds128f(gp), g1 = lds128f(gp + 4);
        float2 f;
        f = __half22float2(h[0]); h[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
        f = __half22float2(h[1]); h[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
        f = __half22float2(h[2]); h[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
        f = __half22float2(h[3]); h[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
        *ptr = raw;
      }
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_g);
      // ---- partial tile: TMEM -> per-warp transpose -> L2 scratch, full 128-byte lines ----
      mbar_wait(&bar_d, 0, 0x6011);
      tc_fence_after();
      pk_workers_sync();   // the staging below reuses operand memory other warps wrote in the gate pass (ordered through
                           // bar_g -> MMA -> bar_d already; the block barrier makes that order explicit)
      const int q = warp & 3, g = warp >> 2;
      uint8_t* stg = sStage + warp * MB_STAGE_WARP_BYTES;
      float* part = sp.part + ((size_t)(mt * S + rank) * 128) * N;
      const int nchunks = sp.d_pitch >> 5;
      for (int cc = g; cc < nchunks; cc += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + cc * 32 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128(stg + (lane * 36 + j4 * 4) * 4, make_uint4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
        __syncwarp();
        const int c = cc * 32 + (lane & 7) * 4;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = q * 32 + rr * 4 + (lane >> 3);
          const uint4 t = lds128(stg + ((rr * 4 + (lane >> 3)) * 36 + (lane & 7) * 4) * 4);
          if (m0 + r < sp.M && c < N) *reinterpret_cast<uint4*>(part + (size_t)r * N + c) = t;
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    cluster_arrive(); cluster_wait();   // release / acquire at cluster scope: every partial tile of this m tile is visible
    // ---- owner: sum the partials in rank order, + bias (+ skip) -> fp16 -> global ----
    {
      const int c4n = N >> 2;
      const int nsrc = min(S, sp.nkb);
      const int r0 = rank * sp.rows_own;
      const int my_rows = max(0, min(min(sp.rows_own, 128 - r0), sp.M - m0 - r0));
      const float* ptile = sp.part + (size_t)mt * S * 128 * N;
      for (int idx = tid; idx < my_rows * c4n; idx += PK_WORKERS) {
        const int rl = idx / c4n, col = (idx - rl * c4n) * 4;
        const int row = r0 + rl;
        float4 a = __ldg(reinterpret_cast<const float4*>(sp.bias + col));
        float4 v[6];
#pragma unroll
        for (int d = 0; d < 6; ++d)
          v[d] = d < nsrc ? __ldcg(reinterpret_cast<const float4*>(ptile + ((size_t)d * 128 + row) * N + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 6; ++d) { a.x += v[d].x; a.y += v[d].y; a.z += v[d].z; a.w += v[d].w; }
        const size_t o = (size_t)(m0 + row) * N + col;
        if (sp.residual) {
          const uint2 rv = *reinterpret_cast<const uint2*>(sp.residual + o);
          const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&rv.x));
          const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&rv.y));
          a.x += x0.x; a.y += x0.y; a.z += x1.x; a.w += x1.y;
        }
        uint2 ov;
        *reinterpret_cast<__half2*>(&ov.x) = __floats2half2_rn(a.x, a.y);
        *reinterpret_cast<__half2*>(&ov.y) = __floats2half2_rn(a.z, a.w);
        *reinterpret_cast<uint2*>(sp.out + o) = ov;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PK_WORKERS / 32) tmem_dealloc(tmem_base, (uint32_t)sp.tmem_cols);
}

}  // namespace hp
