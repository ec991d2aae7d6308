/*
 * hmdpose_internal.h -- test and profiling hooks of libhmdpose.so.  NOT part of the drop-in surface (hmdpose.h):
 * used by tests/, bench.py (roofline section) and tools/.  They may change without an ABI version bump.
 */
#ifndef HMDPOSE_INTERNAL_H
#define HMDPOSE_INTERNAL_H

#include "hmdpose.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Copy a named intermediate activation of the LAST run to host as fp32, NHWC order.  Returns the
 * number of elements (or a negative error); with out == NULL only returns the count.  Handles created with
 * HMDPOSE_KEEP_ALL=1 in the environment keep every intermediate tensor (no buffer re-use). */
int64_t hmdpose_debug_read(hmdpose_t* h, const char* name, float* out, int64_t capacity);
/* Per-launch device times of one pass over `batch` frames (<= micro-batch), measured with CUDA events on
 * the handle's stream around every kernel launch (un-graphed) and averaged over `reps` repetitions after
 * one warm-up.  mode: 0 = network only, 1 = + detection post-processing, 2 = + C# best-pose selection;
 * | 0x100 = in-situ cost (graph of steps[0..k] minus graph of steps[0..k-1]).
 * names/kernels: capacity x 64 chars (step name / kernel function); bytes/flops: the algorithmic
 * (compulsory) HBM bytes and 2*MAC flops of each launch as fused.  Inputs are whatever the
 * last host-API call staged.  Returns the number of launches (with ms == NULL: only the count). */
int hmdpose_profile_steps(hmdpose_t* h, int batch, int mode, int reps, char* names, char* kernels, float* ms,
                          double* bytes, double* flops, int capacity);
/* Standalone pointwise-GEMM check: D[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual) in the given
 * precision mode, host fp32 in/out.  impl: 0 = FFMA cross-check kernel, 1 = tcgen05 kind::f16 (fast mode),
 * 2 = its first version, 3 = tcgen05 3xTF32 split precision (parity mode). */
int hmdpose_test_gemm(int device, int impl, int precision, int M, int N, int K, const float* A,
                      const float* W, const float* bias, const float* a_scale, int rows_per_img,
                      const float* residual, int act, float* D, float* gpu_ms);

#ifdef __cplusplus
}
#endif
#endif
