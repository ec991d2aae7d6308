// Programmatic dependent launch (PDL): every kernel of the launch plan is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, calls pdl_trigger() first thing (so that the next kernel of
// the stream / graph may be scheduled onto SM resources as they free up) and pdl_wait() before it touches
// anything a predecessor wrote.  Launch latency, CTA scheduling and constant-only prologues (mbarrier init, TMEM
// allocation, weight prefetch) of kernel N+1 then overlap the tail of kernel N instead of adding ~2-3 us per launch
// to a ~100-launch dependency chain.  Without the attribute both calls are no-ops.
#pragma once
#include <cuda_runtime.h>

namespace hp {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // engine.cu: false when HMDPOSE_NO_PDL is set

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace hp
