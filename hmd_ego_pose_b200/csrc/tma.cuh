// mbarrier / TMA / proxy-fence PTX wrappers shared by the tcgen05 GEMMs, the fused sepconv and the depthwise kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hp {

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- watchdog --------------------------------------------------------------------------------
// Every mbarrier wait of the library goes through mbar_wait: a tight try_wait spin (the fast path costs nothing
// extra), then -- once the wait is clearly not a pipeline hand-off any more -- __nanosleep back-off and a WALL-CLOCK
// bound (%globaltimer).  A wait that exceeds the bound is a protocol bug or a cross-stream resource deadlock: the
// thread records who / where / how long in a host-mapped TrapInfo (readable by the host even after the context died,
// reported through hmdpose_last_error) and traps, instead of hanging the GPU or dying without a trace.
// The bound is generous (a kernel launched programmatically early legitimately waits for its whole predecessor
// chain, and tools such as compute-sanitizer slow everything down ~100x).
struct TrapInfo {
  unsigned int code;        // 0 = nothing recorded, 1 = mbarrier watchdog
  unsigned int tag;         // call-site tag (kernel << 8 | barrier role)
  unsigned int block, thread;
  unsigned int bar_addr;    // shared-memory address of the barrier
  unsigned int parity;
  unsigned long long waited_ns;
};
static __device__ TrapInfo* g_trap_info = nullptr;   // per translation unit; set by set_trap_info_ptr()
constexpr unsigned long long MBAR_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity, uint32_t tag) {
  const unsigned long long t0 = global_ns();
  uint32_t ns = 32;
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (mbar_try(addr, parity)) return;
      __nanosleep(ns);
    }
    if (ns < 1024) ns <<= 1;
    const unsigned long long waited = global_ns() - t0;
    if (waited > MBAR_TIMEOUT_NS) {
      TrapInfo* ti = g_trap_info;
      if (ti != nullptr && atomicCAS(&ti->code, 0u, 1u) == 0u) {
        ti->tag = tag;
        ti->block = blockIdx.x;
        ti->thread = threadIdx.x;
        ti->bar_addr = addr;
        ti->parity = parity;
        ti->waited_ns = waited;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (int it = 0; it < 4096; ++it)     // hand-off waits resolve here (a failed try_wait already suspends briefly)
    if (mbar_try(addr, parity)) return;
  mbar_wait_slow(addr, parity, tag);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 4-D tiled TMA load of an NHWC activation tile: coordinates (channel, x, y, image); out-of-bounds elements
// (negative or past the edge) are zero-filled, which IS the TF-SAME zero padding of the reference's convs
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// host: point THIS translation unit's copy of g_trap_info at the (host-mapped) record; every TU that contains
// kernels calling mbar_wait installs it once per device (engine.cu, engine_gemm.cu)
static inline cudaError_t trap_info_install_tu(TrapInfo* mapped) {
  return cudaMemcpyToSymbol(g_trap_info, &mapped, sizeof(mapped));
}

}  // namespace hp
