// Post-processing kernels: anchor decode, translation recovery, score threshold + compaction,
// per-class NMS, top-k, gather/pad, and the C# receiver's arg-max pose selection.
//
// This translation unit is compiled WITHOUT FMA contraction (-fmad=false) and without fast-math so that
// every fp32 expression rounds exactly like the reference's CPU arithmetic (hmdegopose/layers.py:142-249,
// TF non_max_suppression_op.cc IOU()): kept indices / labels are bit-exact on identical inputs.
#include "postprocess.h"

#include "pdl.cuh"

#include <math.h>

namespace hp {

// ---------------------------------------------------------------------------------------------
// bbox_transform_inv (layers.py:169-200) + ClipBoxes (layers.py:122-136).  regression = (ty,tx,th,tw).
// exp is evaluated in double and rounded once (CUDA's fp32 expf is 2 ulp; CPU libms are <= 1 ulp).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 decode_box_one(float4 an, float4 d, float wmax, float hmax) {
  const float cxa = (an.x + an.z) / 2.0f;
  const float cya = (an.y + an.w) / 2.0f;
  const float wa = an.z - an.x;
  const float ha = an.w - an.y;
  const float ty = d.x, tx = d.y, th = d.z, tw = d.w;
  const float w = (float)exp((double)tw) * wa;
  const float h = (float)exp((double)th) * ha;
  const float cy = ty * ha + cya;
  const float cx = tx * wa + cxa;
  const float ymin = cy - h / 2.0f;
  const float xmin = cx - w / 2.0f;
  const float ymax = cy + h / 2.0f;
  const float xmax = cx + w / 2.0f;
  float4 o;
  o.x = fminf(fmaxf(xmin, 0.0f), wmax);
  o.y = fminf(fmaxf(ymin, 0.0f), hmax);
  o.z = fminf(fmaxf(xmax, 0.0f), wmax);
  o.w = fminf(fmaxf(ymax, 0.0f), hmax);
  return o;
}

__global__ void decode_boxes_kernel(const float* __restrict__ anchors, const float* __restrict__ reg, int B, int N,
                                    float wmax, float hmax, float* __restrict__ boxes) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int a = (int)(i % N);
  reinterpret_cast<float4*>(boxes)[i] =
      decode_box_one(reinterpret_cast<const float4*>(anchors)[a], reinterpret_cast<const float4*>(reg)[i], wmax, hmax);
}

// translation_transform_inv (layers.py:142-166) + CalculateTxTy (layers.py:212-249), op order as written
__device__ __forceinline__ void decode_translation_one(const float* ta, const float* d, const float* cam, float* out) {
  const float stride = ta[2];
  float x = ta[0] + d[0] * stride;
  float y = ta[1] + d[1] * stride;
  const float fx = cam[0], fy = cam[1], px = cam[2], py = cam[3], tzs = cam[4], ims = cam[5];
  x = x / ims;
  y = y / ims;
  const float tz = d[2] * tzs;
  x = x - px;
  y = y - py;
  out[0] = (x * tz) / fx;
  out[1] = (y * tz) / fy;
  out[2] = tz;
}

__global__ void decode_translation_kernel(const float* __restrict__ tanchors, const float* __restrict__ raw,
                                          const float* __restrict__ cam, int B, int N, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int a = (int)(i % N);
  const int b = (int)(i / N);
  float o[3];
  decode_translation_one(tanchors + 3 * a, raw + 3 * i, cam + 6 * b, o);
  out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

// ---------------------------------------------------------------------------------------------
// TF IOU() (core/kernels/image/non_max_suppression_op.cc): corners normalised per axis, fp32.
// ---------------------------------------------------------------------------------------------
struct Box { float lo0, lo1, hi0, hi1, area; };
__device__ __forceinline__ Box make_box(float4 b) {
  Box r;
  r.lo0 = fminf(b.x, b.z); r.hi0 = fmaxf(b.x, b.z);
  r.lo1 = fminf(b.y, b.w); r.hi1 = fmaxf(b.y, b.w);
  r.area = (r.hi0 - r.lo0) * (r.hi1 - r.lo1);
  return r;
}
__device__ __forceinline__ bool iou_gt(const Box& a, const Box& b, float thr) {
  if (a.area <= 0.0f || b.area <= 0.0f) return false;
  const float d0 = fmaxf(fminf(a.hi0, b.hi0) - fmaxf(a.lo0, b.lo0), 0.0f);
  const float d1 = fmaxf(fminf(a.hi1, b.hi1) - fmaxf(a.lo1, b.lo1), 0.0f);
  const float inter = d0 * d1;
  const float denom = a.area + b.area - inter;   // >= max(area) > 0
  // The IEEE division decides; it is skipped when the outcome is certain.  inter > thr*denom*(1+1e-6) implies
  // fl(inter/denom) > thr and inter < thr*denom*(1-1e-6) implies fl(inter/denom) <= thr (each rounding is 2^-24
  // relative), so the result is bit-identical to TF's `inter / denom > thr` on every input.
  if (thr >= 0.0f) {
    const float t = thr * denom;
    if (inter > t * 1.000001f) return true;
    if (inter < t * 0.999999f) return false;
  }
  const float iou = inter / denom;
  return iou > thr;
}

__device__ __forceinline__ uint32_t float_sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ascending bitonic sort of n (power of two) 64-bit keys by one block
__device__ void bitonic_sort(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// filter_detections, per (image, class): `tf.where(score > thr)` -> tf.image.non_max_suppression
// (layers.py:305-337).  One block per (image, class).
//   keys scratch: [B*C][cap] (cap = power of two >= N); kept_*: [B*C][max_det]; kept_count: [B*C]
// ---------------------------------------------------------------------------------------------
constexpr int FILTER_THREADS = 512;
constexpr int SORT_SMEM = 2048;   // keys sorted in shared memory (rank sort up to 1024, bitonic up to 2048)
constexpr int BOX_CACHE = 1024;   // candidate boxes cached in shared memory after the sort
constexpr int NMS_CHUNK = 64;     // candidates resolved per round (8 threads per candidate)

__device__ __forceinline__ float key_score(unsigned long long key) {
  const uint32_t s = ~(uint32_t)(key >> 32);                       // float_sortable(score)
  return __uint_as_float((s & 0x80000000u) ? (s ^ 0x80000000u) : ~s);
}

__global__ void __launch_bounds__(FILTER_THREADS) filter_nms_kernel(
    const float* __restrict__ boxes, const float* __restrict__ scores, int N, int C, int cap, float score_thr,
    float iou_thr, int max_det, unsigned long long* __restrict__ keys_g, int* __restrict__ kept_idx,
    float* __restrict__ kept_score, int* __restrict__ kept_count) {
  __shared__ unsigned long long skeys[SORT_SMEM];
  __shared__ float4 box_cache[BOX_CACHE];
  __shared__ int warp_tot[2][FILTER_THREADS / 32];
  __shared__ float4 sel_box[MAX_DET_CAP];
  __shared__ int sel_idx[MAX_DET_CAP];
  __shared__ float sel_score[MAX_DET_CAP];
  __shared__ float4 c_box[NMS_CHUNK];
  __shared__ unsigned long long c_mask[NMS_CHUNK];   // bit j set: candidate j (< i) of the round suppresses i
  __shared__ unsigned int c_alive[2];                // bit i set: no earlier-round selection suppresses i
  __shared__ int s_nsel;
  pdl_trigger();
  pdl_wait();

  const int bc = blockIdx.x;
  const int b = bc / C, c = bc - b * C;
  const float* sc = scores + (long long)b * N * C + c;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (long long)b * N;
  unsigned long long* keys = keys_g + (long long)bc * cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = FILTER_THREADS / 32;

  // 1. `tf.where(score > thr)`: ordered compaction, one barrier per sweep (double-buffered warp totals; every
  //    thread keeps the running total itself)
  int n = 0;
  for (int base = 0, sweep = 0; base < N; base += FILTER_THREADS, ++sweep) {
    const int i = base + tid;
    float s = 0.f;
    bool pass = false;
    if (i < N) { s = sc[(long long)i * C]; pass = s > score_thr; }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    int* wt = warp_tot[sweep & 1];
    if (lane == 0) wt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      const int v = wt[w];
      before += (w < warp) ? v : 0;
      total += v;
    }
    if (pass) {
      const int pos = n + before + __popc(bal & ((1u << lane) - 1u));
      keys[pos] = ((unsigned long long)(~float_sortable(s)) << 32) | (unsigned)i;
    }
    n += total;
  }
  __syncthreads();

  // 2. sort by (score desc, anchor index asc).  Keys are unique, so for n <= 1024 every thread ranks its keys by
  //    counting (no barriers); larger sets fall back to a bitonic network (shared, then global memory).
  unsigned long long* sorted = keys;
  if (n > 1 && n <= SORT_SMEM / 2) {
    unsigned long long* src = skeys;
    unsigned long long* dst = skeys + SORT_SMEM / 2;
    for (int i = tid; i < n; i += FILTER_THREADS) src[i] = keys[i];
    __syncthreads();
    for (int i = tid; i < n; i += FILTER_THREADS) {
      const unsigned long long k = src[i];
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < n; ++j) rank += (src[j] < k) ? 1 : 0;
      dst[rank] = k;
    }
    sorted = dst;
  } else if (n > 1) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 <= SORT_SMEM) {
      for (int i = tid; i < np2; i += FILTER_THREADS) skeys[i] = i < n ? keys[i] : ~0ull;
      __syncthreads();
      bitonic_sort(skeys, np2);
      sorted = skeys;
    } else {
      for (int i = n + tid; i < np2; i += FILTER_THREADS) keys[i] = ~0ull;
      __syncthreads();
      bitonic_sort(keys, np2);
    }
  } else if (n == 1) {
    if (tid == 0) skeys[0] = keys[0];
    sorted = skeys;
  }
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  const bool cached = n <= BOX_CACHE;
  if (cached) {
    for (int i = tid; i < n; i += FILTER_THREADS) box_cache[i] = bx[(int)(sorted[i] & 0xffffffffu)];
    __syncthreads();
  }

  // 3. greedy NMS (tf.image.non_max_suppression) in rounds of NMS_CHUNK sorted candidates.  Candidate i of a
  //    round is kept iff no box selected in earlier rounds suppresses it (alive) and no KEPT candidate j < i of
  //    the same round does (pairwise bit matrix, resolved by one thread) -- identical to the sequential scan.
  const int ci = tid >> 3, ct = tid & 7;   // candidate slot / helper thread
  for (int base = 0; base < n; base += NMS_CHUNK) {
    const int nsel0 = s_nsel;
    if (nsel0 >= max_det) break;
    const int cnt = min(NMS_CHUNK, n - base);
    if (tid < cnt) c_box[tid] = cached ? box_cache[base + tid] : bx[(int)(sorted[base + tid] & 0xffffffffu)];
    if (tid < 2) c_alive[tid] = 0u;
    __syncthreads();
    bool dead = false;
    unsigned long long m = 0ull;
    if (ci < cnt) {
      const Box me = make_box(c_box[ci]);
      for (int j = nsel0 - 1 - ct; j >= 0; j -= 8)
        if (iou_gt(me, make_box(sel_box[j]), iou_thr)) { dead = true; break; }
      for (int j = ct; j < ci; j += 8)
        if (iou_gt(me, make_box(c_box[j]), iou_thr)) m |= 1ull << j;
    }
    // combine the 8 helper threads of a candidate (consecutive lanes); executed by every lane of the warp
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      dead |= (__shfl_xor_sync(0xffffffffu, (int)dead, o) != 0);
      m |= __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (ci < cnt && ct == 0) {
      c_mask[ci] = m;
      if (!dead) atomicOr(&c_alive[ci >> 5], 1u << (ci & 31));
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long rem = ((unsigned long long)c_alive[1] << 32) | c_alive[0];
      unsigned long long kept = 0ull;
      int nsel = nsel0;
      while (rem != 0ull && nsel < max_det) {
        const int i = __ffsll((long long)rem) - 1;
        rem &= rem - 1ull;
        if ((c_mask[i] & kept) == 0ull) {
          kept |= 1ull << i;
          const unsigned long long key = sorted[base + i];
          sel_box[nsel] = c_box[i];
          sel_idx[nsel] = (int)(key & 0xffffffffu);
          sel_score[nsel] = key_score(key);
          ++nsel;
        }
      }
      s_nsel = nsel;
    }
    __syncthreads();
  }
  const int nsel = s_nsel;
  for (int i = tid; i < nsel; i += FILTER_THREADS) {
    kept_idx[(long long)bc * max_det + i] = sel_idx[i];
    kept_score[(long long)bc * max_det + i] = sel_score[i];
  }
  if (tid == 0) kept_count[bc] = nsel;
}

// ---------------------------------------------------------------------------------------------
// Single-class fused path (HMD-EgoPose has num_classes = 1): the whole of train.py:72-85 after the network in
// one kernel, one block per image.
//   compaction : every thread scans a contiguous slice of the scores with all loads in flight, one block-wide
//                scan gives the ordered write offsets (tf.where order)
//   sort       : rank by counting for n <= 1024 (keys are unique), bitonic otherwise
//   boxes      : decoded only for the candidates (bbox_transform_inv + ClipBoxes), cached in shared memory
//   NMS        : rounds of 64 candidates; the in-round dependency "kept_i = alive_i and no kept j<i suppresses i"
//                is solved by a warp as a fixed-point iteration on a 64-bit mask (final after the first pass for
//                element 0, after t passes for the first t elements: exactly the sequential greedy result)
//   gather     : boxes / score / label / rotation / translation (decoded for the kept rows only) / hand / anchor
//                index, padded with -1
// ---------------------------------------------------------------------------------------------
constexpr int FUSED_SLICE = 32;   // scores per thread per pass in the compaction

__global__ void __launch_bounds__(FILTER_THREADS) filter_fused_kernel(FilterArgs a) {
  __shared__ unsigned long long skeys[SORT_SMEM];
  __shared__ float4 box_cache[BOX_CACHE];
  __shared__ int warp_tot[FILTER_THREADS / 32];
  __shared__ float4 sel_box[MAX_DET_CAP];
  __shared__ int sel_idx[MAX_DET_CAP];
  __shared__ float sel_score[MAX_DET_CAP];
  __shared__ float4 c_box[NMS_CHUNK];
  __shared__ unsigned long long c_mask[NMS_CHUNK];
  __shared__ unsigned int c_alive[2];
  __shared__ int s_nsel;
  pdl_trigger();
  pdl_wait();

  const int b = blockIdx.x;
  const int N = a.N, max_det = a.max_det;
  const float* sc = a.scores + (long long)b * N;
  unsigned long long* keys = a.keys + (long long)b * a.cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = FILTER_THREADS / 32;
  auto box_of = [&](int idx) -> float4 {
    if (a.boxes) return reinterpret_cast<const float4*>(a.boxes)[(long long)b * N + idx];
    return decode_box_one(reinterpret_cast<const float4*>(a.anchors)[idx],
                          reinterpret_cast<const float4*>(a.reg)[(long long)b * N + idx], a.wmax, a.hmax);
  };

  // 1. ordered compaction
  int n = 0;
  for (int base = 0; base < N; base += FILTER_THREADS * FUSED_SLICE) {
    const int i0 = base + tid * FUSED_SLICE;
    float v[FUSED_SLICE];
#pragma unroll
    for (int j = 0; j < FUSED_SLICE; ++j) v[j] = (i0 + j < N) ? __ldg(sc + i0 + j) : -INFINITY;
    unsigned passmask = 0u;
#pragma unroll
    for (int j = 0; j < FUSED_SLICE; ++j) passmask |= (v[j] > a.score_thr) ? (1u << j) : 0u;
    const int mine = __popc(passmask);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();                      // previous pass has consumed warp_tot
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      const int t = warp_tot[w];
      before += (w < warp) ? t : 0;
      total += t;
    }
    int pos = n + before + incl - mine;
#pragma unroll
    for (int j = 0; j < FUSED_SLICE; ++j)
      if (passmask & (1u << j)) keys[pos++] = ((unsigned long long)(~float_sortable(v[j])) << 32) | (unsigned)(i0 + j);
    n += total;
  }
  __syncthreads();

  // 2. sort by (score desc, anchor index asc)
  unsigned long long* sorted = keys;
  if (n > 1 && n <= SORT_SMEM / 2) {
    unsigned long long* src = skeys;
    unsigned long long* dst = skeys + SORT_SMEM / 2;
    for (int i = tid; i < n; i += FILTER_THREADS) src[i] = keys[i];
    __syncthreads();
    for (int i = tid; i < n; i += FILTER_THREADS) {
      const unsigned long long k = src[i];
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < n; ++j) rank += (src[j] < k) ? 1 : 0;
      dst[rank] = k;
    }
    sorted = dst;
  } else if (n > 1) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 <= SORT_SMEM) {
      for (int i = tid; i < np2; i += FILTER_THREADS) skeys[i] = i < n ? keys[i] : ~0ull;
      __syncthreads();
      bitonic_sort(skeys, np2);
      sorted = skeys;
    } else {
      for (int i = n + tid; i < np2; i += FILTER_THREADS) keys[i] = ~0ull;
      __syncthreads();
      bitonic_sort(keys, np2);
    }
  } else if (n == 1) {
    if (tid == 0) skeys[0] = keys[0];
    sorted = skeys;
  }
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  const bool cached = n <= BOX_CACHE;
  if (cached) {
    for (int i = tid; i < n; i += FILTER_THREADS) box_cache[i] = box_of((int)(sorted[i] & 0xffffffffu));
    __syncthreads();
  }

  // 3. NMS rounds
  const int ci = tid >> 3, ct = tid & 7;
  for (int base = 0; base < n; base += NMS_CHUNK) {
    const int nsel0 = s_nsel;
    if (nsel0 >= max_det) break;
    const int cnt = min(NMS_CHUNK, n - base);
    if (tid < cnt) c_box[tid] = cached ? box_cache[base + tid] : box_of((int)(sorted[base + tid] & 0xffffffffu));
    if (tid < 2) c_alive[tid] = 0u;
    __syncthreads();
    bool dead = false;
    unsigned long long m = 0ull;
    if (ci < cnt) {
      const Box me = make_box(c_box[ci]);
      for (int j = nsel0 - 1 - ct; j >= 0; j -= 8)
        if (iou_gt(me, make_box(sel_box[j]), a.iou_thr)) { dead = true; break; }
      for (int j = ct; j < ci; j += 8)
        if (iou_gt(me, make_box(c_box[j]), a.iou_thr)) m |= 1ull << j;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      dead |= (__shfl_xor_sync(0xffffffffu, (int)dead, o) != 0);
      m |= __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (ci < cnt && ct == 0) {
      c_mask[ci] = m;
      if (!dead) atomicOr(&c_alive[ci >> 5], 1u << (ci & 31));
    }
    __syncthreads();
    if (warp == 0) {
      const unsigned long long alive = ((unsigned long long)c_alive[1] << 32) | c_alive[0];
      const unsigned long long m_lo = lane < cnt ? c_mask[lane] : 0ull;
      const unsigned long long m_hi = lane + 32 < cnt ? c_mask[lane + 32] : 0ull;
      unsigned long long kept = alive;
      for (int iter = 0; iter < NMS_CHUNK; ++iter) {
        const bool k_lo = ((alive >> lane) & 1ull) && (m_lo & kept) == 0ull;
        const bool k_hi = ((alive >> (lane + 32)) & 1ull) && (m_hi & kept) == 0ull;
        const unsigned long long nk = ((unsigned long long)__ballot_sync(0xffffffffu, k_hi) << 32) |
                                      __ballot_sync(0xffffffffu, k_lo);
        if (nk == kept) break;
        kept = nk;
      }
      // the detection cap keeps the first (max_det - nsel0) kept candidates, in order
      int room = max_det - nsel0;
      const int total = __popcll(kept);
      if (total > room) {
        unsigned long long t = kept;
        for (int q = 0; q < room; ++q) t &= t - 1ull;   // clear the `room` lowest set bits
        kept &= ~t;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = lane + 32 * half;
        if ((kept >> i) & 1ull) {
          const int pos = nsel0 + __popcll(kept & ((1ull << i) - 1ull));
          const unsigned long long key = sorted[base + i];
          sel_box[pos] = c_box[i];
          sel_idx[pos] = (int)(key & 0xffffffffu);
          sel_score[pos] = key_score(key);
        }
      }
      if (lane == 0) s_nsel = nsel0 + min(total, room);
    }
    __syncthreads();
  }

  // 4. gather + pad (layers.py:363-384)
  const int nsel = s_nsel;
  const int H = a.H;
  for (int r = tid; r < max_det; r += FILTER_THREADS) {
    const long long o = (long long)b * max_det + r;
    const bool ok = r < nsel;
    const int idx = ok ? sel_idx[r] : -1;
    if (a.o_scores) a.o_scores[o] = ok ? sel_score[r] : -1.0f;
    if (a.o_labels) a.o_labels[o] = ok ? 0 : -1;
    if (a.o_idx) a.o_idx[o] = idx;
    if (a.o_boxes) reinterpret_cast<float4*>(a.o_boxes)[o] = ok ? sel_box[r] : make_float4(-1.f, -1.f, -1.f, -1.f);
    // rotation / translation rows: gathered here from the dense head tensors, or (o_rot == o_trans == null) produced
    // by pose_gather_kernel, which evaluates the headers only at the kept anchors
    float rot[3] = {-1.f, -1.f, -1.f}, tr[3] = {-1.f, -1.f, -1.f};
    if (ok && (a.o_rot || a.o_trans)) {
      const long long row = (long long)b * N + idx;
      if (a.o_rot) { rot[0] = a.rotation[row * 3]; rot[1] = a.rotation[row * 3 + 1]; rot[2] = a.rotation[row * 3 + 2]; }
      if (a.o_trans) {
        if (a.translation) { tr[0] = a.translation[row * 3]; tr[1] = a.translation[row * 3 + 1]; tr[2] = a.translation[row * 3 + 2]; }
        else decode_translation_one(a.tanchors + 3 * idx, a.traw + 3 * row, a.cam + 6 * b, tr);
      }
    }
    if (a.o_rot) { a.o_rot[o * 3] = rot[0]; a.o_rot[o * 3 + 1] = rot[1]; a.o_rot[o * 3 + 2] = rot[2]; }
    if (a.o_trans) { a.o_trans[o * 3] = tr[0]; a.o_trans[o * 3 + 1] = tr[1]; a.o_trans[o * 3 + 2] = tr[2]; }
  }
  if (a.o_hand && a.hand) {
    for (int e = tid; e < max_det * H; e += FILTER_THREADS) {
      const int r = e / H, f = e - r * H;
      a.o_hand[((long long)b * max_det + r) * H + f] = r < nsel ? a.hand[((long long)b * N + sel_idx[r]) * H + f] : -1.0f;
    }
  }
}

void launch_filter_fused(const FilterArgs& a, int B, cudaStream_t st) {
  launch_k(filter_fused_kernel, dim3(B), dim3(FILTER_THREADS), 0, st, a);
}

// ---------------------------------------------------------------------------------------------
// EfficientDet-d0 detection variant.
//   d0_max_kernel : score = max over classes, arg-max class, `score > threshold` (utils/utils.py:93-94,104-108);
//                   passing anchors are appended (unordered) as unique sort keys
//   d0_nms_kernel : one block per image: sort (score desc, anchor asc), BBoxTransform + ClipBoxes for the candidates
//                   (efficientdet/utils.py:7-52), torchvision batched_nms = boxes offset by class*(max_coord+1) then
//                   plain NMS on the OFFSET boxes (IoU = inter/(a_i+a_j-inter), suppress iff > thr), output in keep order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) d0_max_kernel(D0Args a, int B) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * a.N) return;
  const int b = (int)(i / a.N), n = (int)(i - (long long)b * a.N);
  const float* c = a.cls + i * a.C;
  float best = c[0];
  int arg = 0;
  for (int k = 1; k < a.C; ++k) {
    const float v = c[k];
    if (v > best) { best = v; arg = k; }   // first maximum wins, like torch.max
  }
  if (best > a.threshold) {
    const int pos = atomicAdd(a.cand_count + b, 1);
    a.keys[(long long)b * a.cap + pos] = ((unsigned long long)(~float_sortable(best)) << 32) | (unsigned)n;
    a.cand_cls[i] = arg;
  }
}

__device__ __forceinline__ float4 d0_decode(float4 an /*y1,x1,y2,x2*/, float4 d, float wmax, float hmax) {
  const float yca = (an.x + an.z) / 2.0f;
  const float xca = (an.y + an.w) / 2.0f;
  const float ha = an.z - an.x;
  const float wa = an.w - an.y;
  const float w = (float)exp((double)d.w) * wa;
  const float h = (float)exp((double)d.z) * ha;
  const float yc = d.x * ha + yca;
  const float xc = d.y * wa + xca;
  float4 o;
  o.x = fmaxf(xc - w / 2.0f, 0.0f);
  o.y = fmaxf(yc - h / 2.0f, 0.0f);
  o.z = fminf(xc + w / 2.0f, wmax);
  o.w = fminf(yc + h / 2.0f, hmax);
  return o;
}
// torchvision nms_kernel.cpp: no corner normalisation, no empty-box guard (0/0 = NaN never suppresses)
__device__ __forceinline__ bool iou_gt_tv(float4 p, float4 q, float thr) {
  const float ap = (p.z - p.x) * (p.w - p.y), aq = (q.z - q.x) * (q.w - q.y);
  const float w = fmaxf(0.0f, fminf(p.z, q.z) - fmaxf(p.x, q.x));
  const float h = fmaxf(0.0f, fminf(p.w, q.w) - fmaxf(p.y, q.y));
  const float inter = w * h;
  const float ovr = inter / (ap + aq - inter);
  return ovr > thr;
}

__global__ void __launch_bounds__(FILTER_THREADS) d0_nms_kernel(D0Args a) {
  __shared__ unsigned long long skeys[SORT_SMEM];
  __shared__ float4 box_cache[BOX_CACHE];
  __shared__ float4 sel_box[D0_SEL_SMEM];
  __shared__ int s_trunc;
  __shared__ float4 c_box[NMS_CHUNK];
  __shared__ unsigned long long c_mask[NMS_CHUNK];
  __shared__ unsigned int c_alive[2];
  __shared__ float red[FILTER_THREADS / 32];
  __shared__ int s_nsel;
  __shared__ float s_max;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N, max_out = a.max_out;
  unsigned long long* keys = a.keys + (long long)b * a.cap;
  const int n = min(a.cand_count[b], N);
  const int* ccls = a.cand_cls + (long long)b * N;
  auto raw_box = [&](int idx) -> float4 {
    return d0_decode(reinterpret_cast<const float4*>(a.anchors_yxyx)[idx],
                     reinterpret_cast<const float4*>(a.reg)[(long long)b * N + idx], a.wmax, a.hmax);
  };
  // sort
  unsigned long long* sorted = keys;
  if (n > 1 && n <= SORT_SMEM / 2) {
    unsigned long long* src = skeys;
    unsigned long long* dst = skeys + SORT_SMEM / 2;
    for (int i = tid; i < n; i += FILTER_THREADS) src[i] = keys[i];
    __syncthreads();
    for (int i = tid; i < n; i += FILTER_THREADS) {
      const unsigned long long k = src[i];
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < n; ++j) rank += (src[j] < k) ? 1 : 0;
      dst[rank] = k;
    }
    sorted = dst;
  } else if (n > 1) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 <= SORT_SMEM) {
      for (int i = tid; i < np2; i += FILTER_THREADS) skeys[i] = i < n ? keys[i] : ~0ull;
      __syncthreads();
      bitonic_sort(skeys, np2);
      sorted = skeys;
    } else {
      for (int i = n + tid; i < np2; i += FILTER_THREADS) keys[i] = ~0ull;
      __syncthreads();
      bitonic_sort(keys, np2);
    }
  } else if (n == 1) {
    if (tid == 0) skeys[0] = keys[0];
    sorted = skeys;
  }
  if (tid == 0) { s_nsel = 0; s_trunc = 0; }
  float4* gsel = reinterpret_cast<float4*>(a.sel_scratch) + (long long)b * max_out;
  __syncthreads();
  // decode the candidates, max coordinate over all of them (batched_nms: boxes.max())
  const bool cached = n <= BOX_CACHE;
  float4* gbox = reinterpret_cast<float4*>(a.box_scratch) + (long long)b * N;
  float mx = -INFINITY;
  for (int i = tid; i < n; i += FILTER_THREADS) {
    const float4 bx = raw_box((int)(sorted[i] & 0xffffffffu));
    if (cached) box_cache[i] = bx; else gbox[i] = bx;
    mx = fmaxf(mx, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int w = 1; w < FILTER_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    s_max = m;
  }
  __syncthreads();
  const float offs = s_max + 1.0f;
  auto off_box = [&](int i) -> float4 {   // candidate i (sorted position) offset by its class
    const float4 bx = cached ? box_cache[i] : gbox[i];
    const float o = (float)ccls[(int)(sorted[i] & 0xffffffffu)] * offs;
    return make_float4(bx.x + o, bx.y + o, bx.z + o, bx.w + o);
  };
  // NMS rounds (same structure as filter_fused_kernel)
  const int ci = tid >> 3, ct = tid & 7;
  for (int base = 0; base < n; base += NMS_CHUNK) {
    const int nsel0 = s_nsel;
    if (nsel0 >= max_out) {   // capacity reached with candidates left: the reference would keep going
      if (tid == 0) s_trunc = 1;
      break;
    }
    const int cnt = min(NMS_CHUNK, n - base);
    if (tid < cnt) c_box[tid] = off_box(base + tid);
    if (tid < 2) c_alive[tid] = 0u;
    __syncthreads();
    bool dead = false;
    unsigned long long m = 0ull;
    if (ci < cnt) {
      const float4 me = c_box[ci];
      for (int j = nsel0 - 1 - ct; j >= 0; j -= 8)
        if (iou_gt_tv(j < D0_SEL_SMEM ? sel_box[j] : gsel[j], me, a.iou_thr)) { dead = true; break; }
      for (int j = ct; j < ci; j += 8)
        if (iou_gt_tv(c_box[j], me, a.iou_thr)) m |= 1ull << j;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      dead |= (__shfl_xor_sync(0xffffffffu, (int)dead, o) != 0);
      m |= __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (ci < cnt && ct == 0) {
      c_mask[ci] = m;
      if (!dead) atomicOr(&c_alive[ci >> 5], 1u << (ci & 31));
    }
    __syncthreads();
    if (warp == 0) {
      const unsigned long long alive = ((unsigned long long)c_alive[1] << 32) | c_alive[0];
      const unsigned long long m_lo = lane < cnt ? c_mask[lane] : 0ull;
      const unsigned long long m_hi = lane + 32 < cnt ? c_mask[lane + 32] : 0ull;
      unsigned long long kept = alive;
      for (int iter = 0; iter < NMS_CHUNK; ++iter) {
        const bool k_lo = ((alive >> lane) & 1ull) && (m_lo & kept) == 0ull;
        const bool k_hi = ((alive >> (lane + 32)) & 1ull) && (m_hi & kept) == 0ull;
        const unsigned long long nk = ((unsigned long long)__ballot_sync(0xffffffffu, k_hi) << 32) |
                                      __ballot_sync(0xffffffffu, k_lo);
        if (nk == kept) break;
        kept = nk;
      }
      const int room = max_out - nsel0;
      const int total = __popcll(kept);
      if (total > room) {
        if (lane == 0) s_trunc = 1;
        unsigned long long t = kept;
        for (int q = 0; q < room; ++q) t &= t - 1ull;
        kept &= ~t;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = lane + 32 * half;
        if ((kept >> i) & 1ull) {
          const int pos = nsel0 + __popcll(kept & ((1ull << i) - 1ull));
          const unsigned long long key = sorted[base + i];
          const int idx = (int)(key & 0xffffffffu);
          if (pos < D0_SEL_SMEM) sel_box[pos] = c_box[i]; else gsel[pos] = c_box[i];
          const long long o = (long long)b * max_out + pos;
          reinterpret_cast<float4*>(a.o_rois)[o] = cached ? box_cache[base + i] : gbox[base + i];
          a.o_cls[o] = ccls[idx];
          a.o_scores[o] = key_score(key);
          a.o_idx[o] = idx;
        }
      }
      if (lane == 0) s_nsel = nsel0 + min(total, room);
    }
    __syncthreads();
  }
  const int nsel = s_nsel;
  for (int r = nsel + tid; r < max_out; r += FILTER_THREADS) {
    const long long o = (long long)b * max_out + r;
    reinterpret_cast<float4*>(a.o_rois)[o] = make_float4(-1.f, -1.f, -1.f, -1.f);
    a.o_cls[o] = -1; a.o_scores[o] = -1.0f; a.o_idx[o] = -1;
  }
  if (tid == 0) a.o_count[b] = s_trunc ? -nsel : nsel;
}

// ---------------------------------------------------------------------------------------------
// Frame pre-processing: cv2.resize (INTER_LINEAR on uint8: OpenCV's 11-bit fixed-point formula) so that the long side
// is S, /255 in float32, (x - mean) and (/ std) evaluated in float64 and rounded to float32 as numpy does for the
// reference's in-place ops, zero padding bottom/right.  One thread per output pixel.  (-fmad=false TU.)
// ---------------------------------------------------------------------------------------------
// OpenCV resize.cpp coefficient tables.  Horizontal: a tap outside the image gets weight 0.  Vertical: the weights are
// kept and the ROW INDICES are clamped (on the border rows of an up-scaled image both taps read the same row).
__device__ __forceinline__ void resize_coeff(int d, int dn, int sn, int& s0, int& s1, int& a0, int& a1, bool vertical = false) {
  const double scale = (double)sn / (double)dn;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (!vertical) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  a1 = __float2int_rn(f * 2048.0f);            // cvRound: round half to even
  a0 = __float2int_rn((1.0f - f) * 2048.0f);
  s0 = min(max(s, 0), sn - 1);
  s1 = min(max(s + 1, 0), sn - 1);
}
__device__ __forceinline__ int resize_vert(int h0, int h1, int b0, int b1) {
  const int u = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
  return min(max(u, 0), 255);
}

__global__ void __launch_bounds__(256) preprocess_kernel(PreArgs a) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.B * a.S * a.S;
  if (i >= total) return;
  const int x = (int)(i % a.S), y = (int)((i / a.S) % a.S), b = (int)(i / ((long long)a.S * a.S));
  float* o = a.out + i * 3;
  if (x >= a.rw || y >= a.rh) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; return; }
  int sx0, sx1, ax0, ax1, sy0, sy1, ay0, ay1;
  resize_coeff(x, a.rw, a.w, sx0, sx1, ax0, ax1);
  resize_coeff(y, a.rh, a.h, sy0, sy1, ay0, ay1, true);
  const uint8_t* im = a.img + (long long)b * a.h * a.w * 3;
  const uint8_t* r0 = im + (long long)sy0 * a.w * 3;
  const uint8_t* r1 = im + (long long)sy1 * a.w * 3;
  const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)r0[sx0 * 3 + c] * ax0 + (int)r0[sx1 * 3 + c] * ax1;
    const int h1 = (int)r1[sx0 * 3 + c] * ax0 + (int)r1[sx1 * 3 + c] * ax1;
    int u = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
    u = min(max(u, 0), 255);
    float v = __fdiv_rn((float)u, 255.0f);
    v = (float)((double)v - mean[c]);
    v = (float)((double)v / stdv[c]);
    o[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// The C# receiver's frame path in ONE kernel (WebRTCNetCoreSandbox/Program.cs:137-200, 381-445), bit-exact against the
// OpenCV calls it makes: cvtColor(YUV2BGR_YV12) on the I420 buffer (chroma planes read swapped; BT.601 20-bit fixed
// point, imgproc/color_yuv.simd.hpp), centre crop, resize to mid x mid, resize so that the long side is S (both
// INTER_LINEAR on 8-bit data: every stage rounds to uint8 like OpenCV does), then ConvertTo(CV_32F), Divide(255),
// Subtract(mean), Divide(std) in FLOAT32 (OpenCV converts the scalars to the Mat's depth), zero pad.
// One thread per output pixel: 4 taps of the mid image, each 4 taps of the crop, each one YUV -> BGR conversion.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void yv12_bgr(const uint8_t* fr, int h, int w, int yy, int xx, int* bgr) {
  const int n = h * w;
  const int y = max(0, (int)fr[yy * w + xx] - 16) * 1220542;
  const int ci = (yy >> 1) * (w >> 1) + (xx >> 1);
  const int vv = (int)fr[n + ci] - 128;              // first chroma plane, read as V (it holds the I420 U plane)
  const int uu = (int)fr[n + (n >> 2) + ci] - 128;
  const int half = 1 << 19;
  bgr[0] = min(max((y + half + 2116026 * uu) >> 20, 0), 255);
  bgr[1] = min(max((y + half - 852492 * vv - 409993 * uu) >> 20, 0), 255);
  bgr[2] = min(max((y + half + 1673527 * vv) >> 20, 0), 255);
}

__global__ void __launch_bounds__(256) preprocess_i420_kernel(I420Args a) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.B * a.S * a.S;
  if (i >= total) return;
  const int x = (int)(i % a.S), y = (int)((i / a.S) % a.S), b = (int)(i / ((long long)a.S * a.S));
  float* o = a.out + i * 3;
  if (x >= a.rw || y >= a.rh) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; return; }
  const uint8_t* fr = a.img + (long long)b * (a.h * a.w * 3 / 2);
  const int off_w = (a.w - a.crop) / 2, off_h = (a.h - a.crop) / 2;
  int sx[2], sy[2], ax[2], ay[2];
  resize_coeff(x, a.rw, a.mid, sx[0], sx[1], ax[0], ax[1]);
  resize_coeff(y, a.rh, a.mid, sy[0], sy[1], ay[0], ay[1], true);
  int hsum[2][3];   // horizontal pass of the second resize on the two mid rows
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int cy[2], by[2];
    resize_coeff(sy[r], a.mid, a.crop, cy[0], cy[1], by[0], by[1], true);
    int m[2][3];    // the two mid pixels (sy[r], sx[0..1])
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int cx[2], bx[2];
      resize_coeff(sx[q], a.mid, a.crop, cx[0], cx[1], bx[0], bx[1]);
      int hh[2][3];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        int p0[3], p1[3];
        yv12_bgr(fr, a.h, a.w, off_h + cy[rr], off_w + cx[0], p0);
        yv12_bgr(fr, a.h, a.w, off_h + cy[rr], off_w + cx[1], p1);
#pragma unroll
        for (int c = 0; c < 3; ++c) hh[rr][c] = p0[c] * bx[0] + p1[c] * bx[1];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) m[q][c] = resize_vert(hh[0][c], hh[1][c], by[0], by[1]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) hsum[r][c] = m[0][c] * ax[0] + m[1][c] * ax[1];
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int u = resize_vert(hsum[0][c], hsum[1][c], ay[0], ay[1]);
    float v = __fdiv_rn((float)u, 255.0f);
    v = __fsub_rn(v, mean[c]);
    o[c] = __fdiv_rn(v, stdv[c]);
  }
}

void launch_preprocess_i420(const I420Args& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.S * a.S;
  launch_k(preprocess_i420_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a);
}

void launch_preprocess(const PreArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.S * a.S;
  launch_k(preprocess_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a);
}

void launch_d0(const D0Args& a, int B, cudaStream_t st) {
  cudaMemsetAsync(a.cand_count, 0, sizeof(int) * B, st);
  const long long total = (long long)B * a.N;
  launch_k(d0_max_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a, B);
  launch_k(d0_nms_kernel, dim3(B), dim3(FILTER_THREADS), 0, st, a);
}

void launch_d0_nms(const D0Args& a, int B, cudaStream_t st) {
  launch_k(d0_nms_kernel, dim3(B), dim3(FILTER_THREADS), 0, st, a);
}

// ---------------------------------------------------------------------------------------------
// Concatenate the per-class keeps (class-major, layers.py:349-358), tf.nn.top_k (descending, ties ->
// lower position), gather, pad with -1, labels -> int32 (layers.py:363-384).  One block per image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_gather_kernel(
    const float* __restrict__ boxes, const float* __restrict__ rotation, const float* __restrict__ translation,
    const float* __restrict__ hand, int N, int C, int H, int max_det, const int* __restrict__ kept_idx,
    const float* __restrict__ kept_score, const int* __restrict__ kept_count, float* __restrict__ o_boxes,
    float* __restrict__ o_scores, int* __restrict__ o_labels, float* __restrict__ o_rot, float* __restrict__ o_trans,
    float* __restrict__ o_hand, int* __restrict__ o_idx) {
  __shared__ int slot_anchor[MAX_DET_CAP];
  __shared__ int slot_label[MAX_DET_CAP];
  __shared__ float slot_score[MAX_DET_CAP];
  __shared__ int s_total;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int t = 0;
    for (int c = 0; c < C; ++c) t += kept_count[b * C + c];
    s_total = t;
  }
  for (int i = tid; i < max_det; i += blockDim.x) { slot_anchor[i] = -1; slot_label[i] = -1; slot_score[i] = -1.0f; }
  __syncthreads();
  const int total = s_total;
  const int k = min(max_det, total);
  // single class: the NMS selection order already is (score desc, position asc) -> top_k is the identity
  if (C == 1) {
    for (int j = tid; j < k; j += blockDim.x) {
      slot_anchor[j] = kept_idx[(long long)b * max_det + j];
      slot_label[j] = 0;
      slot_score[j] = kept_score[(long long)b * max_det + j];
    }
  }
  // rank by counting over the class-major concatenation
  for (int e = tid; C > 1 && e < C * max_det; e += blockDim.x) {
    const int c = e / max_det, j = e - c * max_det;
    if (j >= kept_count[b * C + c]) continue;
    const float s = kept_score[(long long)(b * C + c) * max_det + j];
    int rank = 0;
    for (int c2 = 0; c2 < C; ++c2) {
      const int n2 = kept_count[b * C + c2];
      const float* s2 = kept_score + (long long)(b * C + c2) * max_det;
      for (int j2 = 0; j2 < n2; ++j2) {
        const float o = s2[j2];
        const bool before = (c2 < c) || (c2 == c && j2 < j);
        if (o > s || (o == s && before)) ++rank;
      }
    }
    if (rank < k) {
      slot_anchor[rank] = kept_idx[(long long)(b * C + c) * max_det + j];
      slot_label[rank] = c;
      slot_score[rank] = s;
    }
  }
  __syncthreads();
  for (int r = tid; r < max_det; r += blockDim.x) {
    const long long o = (long long)b * max_det + r;
    if (o_scores) o_scores[o] = slot_score[r];
    if (o_labels) o_labels[o] = slot_label[r];
    if (o_idx) o_idx[o] = slot_anchor[r];
  }
  const int per = 4 + 3 + 3 + H;
  for (int e = tid; e < max_det * per; e += blockDim.x) {
    const int r = e / per, f = e - r * per;
    const int a = slot_anchor[r];
    const long long row = (long long)b * N + a;
    const long long orow = (long long)b * max_det + r;
    if (f < 4) {
      if (o_boxes) o_boxes[orow * 4 + f] = a >= 0 ? boxes[row * 4 + f] : -1.0f;
    } else if (f < 7) {
      if (o_rot) o_rot[orow * 3 + (f - 4)] = a >= 0 ? rotation[row * 3 + (f - 4)] : -1.0f;
    } else if (f < 10) {
      if (o_trans) o_trans[orow * 3 + (f - 7)] = a >= 0 ? translation[row * 3 + (f - 7)] : -1.0f;
    } else {
      if (o_hand) o_hand[orow * H + (f - 10)] = a >= 0 ? hand[row * H + (f - 10)] : -1.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// C# receiver (Program.cs:786-960): the pose of the arg-max-score anchor if its score > thr
// (ties -> lowest anchor index), boxes decoded with the C# twin's column convention
// (Program.cs:654-784), Rect fields truncated to int, rotation * pi, translation / 1000.
// One block per image; out11 per image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) best_kernel(const float* __restrict__ anchors, const float* __restrict__ tanchors,
                                                   const float* __restrict__ reg, const float* __restrict__ scores,
                                                   const float* __restrict__ rot, const float* __restrict__ traw,
                                                   const float* __restrict__ cam, int N, int C, float score_thr,
                                                   float wmax, float hmax, float* __restrict__ out11) {
  __shared__ unsigned long long red[256];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  unsigned long long best = 0ull;
  for (int i = tid; i < N; i += blockDim.x) {
    const float s = scores[((long long)b * N + i) * C];
    if (s > score_thr) {
      const unsigned long long key = ((unsigned long long)float_sortable(s) << 32) | (unsigned)(0xffffffffu - (unsigned)i);
      best = key > best ? key : best;
    }
  }
  red[tid] = best;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] = red[tid + o] > red[tid] ? red[tid + o] : red[tid];
    __syncthreads();
  }
  if (tid != 0) return;
  float* out = out11 + 11 * b;
  for (int i = 0; i < 11; ++i) out[i] = 0.0f;
  if (red[0] == 0ull) return;
  const int a = (int)(0xffffffffu - (unsigned)(red[0] & 0xffffffffu));
  const long long row = (long long)b * N + a;
  const float* an = anchors + 4 * a;
  const float* d = reg + 4 * row;
  const float tx = d[0], ty = d[1], th = d[2], tw = d[3];  // Program.cs:672-716 column naming
  const float cxa = (an[0] + an[2]) / 2.0f;
  const float cya = (an[1] + an[3]) / 2.0f;
  const float wa = an[2] - an[0];
  const float ha = an[3] - an[1];
  const float w = (float)exp((double)tw) * wa;
  const float h = (float)exp((double)th) * ha;
  const float cy = ty * ha + cya;
  const float cx = tx * wa + cxa;
  float ymin = cy - h / 2.0f, xmin = cx - w / 2.0f, ymax = cy + h / 2.0f, xmax = cx + w / 2.0f;
  xmin = fminf(fmaxf(xmin, 0.0f), wmax);
  ymin = fminf(fmaxf(ymin, 0.0f), hmax);
  xmax = fminf(fmaxf(xmax, 0.0f), wmax);
  ymax = fminf(fmaxf(ymax, 0.0f), wmax);  // Program.cs:773 clamps ymax with width
  out[0] = scores[row * C];
  out[1] = truncf(xmin); out[2] = truncf(ymin); out[3] = truncf(xmax); out[4] = truncf(ymax);
  const float pi = 3.14159274101257324f;  // (float)Math.PI
  out[5] = rot[row * 3] * pi; out[6] = rot[row * 3 + 1] * pi; out[7] = rot[row * 3 + 2] * pi;
  float t[3];
  decode_translation_one(tanchors + 3 * a, traw + 3 * row, cam + 6 * b, t);
  const float mm = 1.0f / 1000.0f;
  out[8] = t[0] * mm; out[9] = t[1] * mm; out[10] = t[2] * mm;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
void launch_decode_boxes(const float* anchors, const float* reg, int B, int N, int width, int height, float* boxes,
                         cudaStream_t st) {
  const long long total = (long long)B * N;
  launch_k(decode_boxes_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, anchors, reg, B, N, (float)(width - 1),
                                                                         (float)(height - 1), boxes);
}
void launch_decode_translation(const float* tanchors, const float* raw, const float* cam, int B, int N, float* out,
                               cudaStream_t st) {
  const long long total = (long long)B * N;
  launch_k(decode_translation_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, tanchors, raw, cam, B, N, out);
}
void launch_filter_nms(const PostBuffers& pb, const float* boxes, const float* scores, int B, int N, int C,
                       float score_thr, float iou_thr, int max_det, cudaStream_t st) {
  launch_k(filter_nms_kernel, dim3(B * C), dim3(FILTER_THREADS), 0, st, boxes, scores, N, C, pb.cap, score_thr, iou_thr, max_det,
                                                      pb.keys, pb.kept_idx, pb.kept_score, pb.kept_count);
}
void launch_topk_gather(const PostBuffers& pb, const float* boxes, const float* rotation, const float* translation,
                        const float* hand, int B, int N, int C, int H, int max_det, float* o_boxes, float* o_scores,
                        int* o_labels, float* o_rot, float* o_trans, float* o_hand, int* o_idx, cudaStream_t st) {
  launch_k(topk_gather_kernel, dim3(B), dim3(256), 0, st, boxes, rotation, translation, hand, N, C, H, max_det, pb.kept_idx,
                                        pb.kept_score, pb.kept_count, o_boxes, o_scores, o_labels, o_rot, o_trans,
                                        o_hand, o_idx);
}
void launch_best(const float* anchors, const float* tanchors, const float* reg, const float* scores, const float* rot,
                 const float* traw, const float* cam, int B, int N, int C, float score_thr, int width, int height,
                 float* out11, cudaStream_t st) {
  launch_k(best_kernel, dim3(B), dim3(256), 0, st, anchors, tanchors, reg, scores, rot, traw, cam, N, C, score_thr, (float)(width - 1),
                                 (float)(height - 1), out11);
}

}  // namespace hp
