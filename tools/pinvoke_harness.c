/*
 * P/Invoke stand-in (dotnet/mono are not in this image): a plain-C caller that dlopen()s libhmdpose.so and
 * calls the exact cdecl symbols the C# [DllImport] declarations in INTEGRATION.md bind, with the managed
 * call pattern of unity-sandbox/WebRTCNetCoreSandbox/Program.cs:208-276: one float[196608] frame in,
 * one pose out, one call per frame.  Reports p50 / p99 wall-clock latency of hmdpose_run_best.
 *
 *   pinvoke_harness <libhmdpose.so> <weights.blob> [image_size=256] [frames=5000] [warmup=200] [precision=1] [u8=0]
 * u8 = 1: the frame is a 504 x 896 uint8 RGB image (a HoloLens 2 video frame) and the call is hmdpose_run_best_u8, i.e.
 * the receiver's ResizeAndNormalizeMat (Program.cs:397-445) runs on the device as well.
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hmdpose.h"

typedef void (*fn_default_config)(hmdpose_config_t*);
typedef int (*fn_create_ex)(const hmdpose_config_t*, const char*, hmdpose_t**);
typedef int (*fn_run_best)(hmdpose_t*, const float*, const float*, float*);
typedef int (*fn_run_best_u8)(hmdpose_t*, const uint8_t*, int, int, const float*, float*, float*);
typedef void (*fn_destroy)(hmdpose_t*);
typedef const char* (*fn_last_error)(const hmdpose_t*);
typedef float (*fn_last_gpu_ms)(const hmdpose_t*);
typedef int (*fn_launches)(const hmdpose_t*);

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s libhmdpose.so weights.blob [size] [frames] [warmup] [precision]\n", argv[0]);
    return 2;
  }
  const int size = argc > 3 ? atoi(argv[3]) : 256;
  const int frames = argc > 4 ? atoi(argv[4]) : 5000;
  const int warmup = argc > 5 ? atoi(argv[5]) : 200;
  const int precision = argc > 6 ? atoi(argv[6]) : HMDPOSE_PRECISION_FAST;
  const int u8 = argc > 7 ? atoi(argv[7]) : 0;
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  fn_default_config default_config = (fn_default_config)dlsym(lib, "hmdpose_default_config");
  fn_create_ex create_ex = (fn_create_ex)dlsym(lib, "hmdpose_create_ex");
  fn_run_best run_best = (fn_run_best)dlsym(lib, "hmdpose_run_best");
  fn_run_best_u8 run_best_u8 = (fn_run_best_u8)dlsym(lib, "hmdpose_run_best_u8");
  fn_destroy destroy = (fn_destroy)dlsym(lib, "hmdpose_destroy");
  fn_last_error last_error = (fn_last_error)dlsym(lib, "hmdpose_last_error");
  fn_last_gpu_ms last_gpu_ms = (fn_last_gpu_ms)dlsym(lib, "hmdpose_last_gpu_ms");
  fn_launches launches = (fn_launches)dlsym(lib, "hmdpose_last_launch_count");
  if (!default_config || !create_ex || !run_best || !run_best_u8 || !destroy || !last_error || !last_gpu_ms || !launches) {
    fprintf(stderr, "missing symbol\n");
    return 1;
  }
  hmdpose_config_t cfg;
  default_config(&cfg);
  cfg.image_size = size; cfg.max_batch = 1; cfg.precision = precision;
  hmdpose_t* h = NULL;
  int rc = create_ex(&cfg, argv[2], &h);
  if (rc != 0) { fprintf(stderr, "hmdpose_create_ex failed (%d): %s\n", rc, last_error(NULL)); return 1; }
  const size_t n = (size_t)3 * size * size;
  float* frame = (float*)malloc(n * sizeof(float));  /* pageable, like a pinned-by-GC managed array */
  unsigned s = 12345u;
  /* standard-normal pixels (Box-Muller over an LCG): the statistics of a normalised camera frame, so the synthetic
   * network fires (score > 0.5) and the best-pose decode branch is inside the timed call */
  for (size_t i = 0; i < n; i += 2) {
    s = s * 1664525u + 1013904223u; const double u1 = ((s >> 8) + 1.0) / 16777217.0;
    s = s * 1664525u + 1013904223u; const double u2 = (s >> 8) / 16777216.0;
    const double r = sqrt(-2.0 * log(u1));
    frame[i] = (float)(r * cos(6.283185307179586 * u2));
    if (i + 1 < n) frame[i + 1] = (float)(r * sin(6.283185307179586 * u2));
  }
  const int fh = 504, fw = 896;
  uint8_t* frame8 = (uint8_t*)malloc((size_t)fh * fw * 3);
  for (size_t i = 0; i < (size_t)fh * fw * 3; ++i) { s = s * 1664525u + 1013904223u; frame8[i] = (uint8_t)(s >> 24); }
  const float cam[6] = {480.f, 480.f, 128.f, 128.f, 1000.f, 1.f};
  float out[HMDPOSE_BEST_LEN];
  double* lat = (double*)malloc(sizeof(double) * (size_t)frames);
  double gpu_sum = 0.0;
  for (int i = 0; i < warmup + frames; ++i) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    rc = u8 ? run_best_u8(h, frame8, fh, fw, cam, out, NULL) : run_best(h, frame, cam, out);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (rc != 0) { fprintf(stderr, "hmdpose_run_best failed (%d): %s\n", rc, last_error(h)); return 1; }
    if (i >= warmup) {
      lat[i - warmup] = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
      gpu_sum += last_gpu_ms(h);
    }
  }
  qsort(lat, (size_t)frames, sizeof(double), cmp_double);
  printf("{\"harness\": \"pinvoke_stand_in\", \"api\": \"%s\", \"image_size\": %d, \"precision\": %d, "
         "\"frames\": %d, \"warmup\": %d, \"p50_ms\": %.4f, \"p90_ms\": %.4f, \"p99_ms\": %.4f, \"min_ms\": %.4f, "
         "\"gpu_ms_mean\": %.4f, \"launches_per_frame\": %d, \"score\": %.5f}\n",
         u8 ? "hmdpose_run_best_u8 (504x896 uint8 frame, pre-processing on the device)" : "hmdpose_run_best", size, precision,
         frames, warmup, lat[frames / 2], lat[(int)(frames * 0.9)], lat[(int)(frames * 0.99)], lat[0],
         gpu_sum / frames, launches(h), out[0]);
  destroy(h);
  free(frame); free(frame8); free(lat);
  dlclose(lib);
  return 0;
}
