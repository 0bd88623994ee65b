/* c_abi_smoke.c -- the hot path driven from C ALONE through include/fav_b200.h: K1 (self-sufficient
 * fav_corrupt_normalize: no host-language tables) -> K2 (fav_forward_mc) -> K3+K4 (fav_epilogue_accumulate).
 * This is what a non-Python binder of the reference seam (vision_simulator.py:25-36 knobs -> corruption / severity,
 * main.py:160 per-frame call) would write.  The only input is a weight blob (format FAVW1, fav_load_weights).
 * Prints one line per sweep cell with the integer arena header + a checksum of the whole arena row;
 * tests/test_gpu_parity.py compares them with the Python host path (bit-identical).
 *
 * build: see failure-aware-vision_b200/csrc/build.sh (gcc, links libfav_b200.so + libcudart). */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "fav_b200.h"

#define CHECK(rc, what)                                                      \
  do {                                                                       \
    if ((rc) != 0) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", what, (int)(rc), fav_last_error()); \
      return 2;                                                              \
    }                                                                        \
  } while (0)
#define CU(expr)                                                             \
  do {                                                                       \
    cudaError_t e_ = (expr);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(e_));            \
      return 3;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: c_abi_smoke weights.favw [n_images] [T]\n"); return 1; }
  const int N = argc > 2 ? atoi(argv[2]) : 96, T = argc > 3 ? atoi(argv[3]) : 4, C = 10, H = 32, W = 32;
  const uint64_t seed = 6;
  const float tau = 0.5f, p_drop = 0.2f;
  const float mean[3] = {0.4914f, 0.4822f, 0.4465f}, std[3] = {0.2470f, 0.2435f, 0.2616f};
  const int cells[][2] = {{FAV_CLEAN, 0}, {FAV_GAUSSIAN_NOISE, 3}, {FAV_SHOT_NOISE, 2}, {FAV_DEFOCUS_BLUR, 4}, {FAV_MOTION_BLUR, 5},
                          {FAV_ZOOM_BLUR, 1}, {FAV_FOG, 2}, {FAV_CONTRAST, 3}, {FAV_PIXELATE, 5}, {FAV_JPEG, 1},
                          {FAV_GLASS_BLUR, 4}, {FAV_SNOW, 2}, {FAV_ELASTIC, 2}, {FAV_FROST, 3}, {FAV_IMPULSE_NOISE, 5},
                          {FAV_BRIGHTNESS, 4}};
  const int n_cells = (int)(sizeof(cells) / sizeof(cells[0]));

  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  fseek(f, 0, SEEK_END);
  const long nbytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* blob = malloc((size_t)nbytes);
  if (fread(blob, 1, (size_t)nbytes, f) != (size_t)nbytes) { fprintf(stderr, "short read\n"); return 1; }
  fclose(f);

  fav_handle h = NULL;
  CHECK(fav_abi_version() != FAV_ABI_VERSION, "fav_abi_version");
  CHECK(fav_init(0, &h), "fav_init");
  CHECK(fav_load_weights(h, blob, (size_t)nbytes, FAV_RESNET18, C, H, W), "fav_load_weights");
  CHECK(fav_reserve(h, N, T), "fav_reserve");

  const size_t words = fav_hist_words(C, 15, 4096);
  uint8_t* d_img; int32_t* d_lab; void* d_x; float* d_logits; int64_t* d_arena;
  CU(cudaMalloc((void**)&d_img, (size_t)N * H * W * 3));
  CU(cudaMalloc((void**)&d_lab, (size_t)N * 4));
  CU(cudaMalloc(&d_x, (size_t)N * H * W * 3 * 2));
  CU(cudaMalloc((void**)&d_logits, (size_t)N * T * C * 4));
  CU(cudaMalloc((void**)&d_arena, (size_t)n_cells * words * 8));
  CU(cudaMemset(d_arena, 0, (size_t)n_cells * words * 8));
  CHECK(fav_synth_images(h, d_img, N, H, W, seed, 0, NULL), "fav_synth_images");
  CHECK(fav_synth_labels(h, d_lab, N, C, seed, 0, NULL), "fav_synth_labels");
  for (int i = 0; i < n_cells; ++i) {
    CHECK(fav_corrupt_normalize(h, d_img, d_x, N, H, W, cells[i][0], cells[i][1], seed, 0, mean, std, 0, NULL), "fav_corrupt_normalize");
    CHECK(fav_forward_mc(h, d_x, d_logits, N, T, p_drop, seed, 0, NULL), "fav_forward_mc");
    CHECK(fav_epilogue_accumulate(h, d_logits, d_lab, N, T, C, tau, 15, 4096, d_arena + (size_t)i * words, NULL, NULL, NULL, NULL,
                                  NULL, NULL), "fav_epilogue_accumulate");
  }
  CU(cudaDeviceSynchronize());
  int64_t* arena = (int64_t*)malloc((size_t)n_cells * words * 8);
  CU(cudaMemcpy(arena, d_arena, (size_t)n_cells * words * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n_cells; ++i) {
    const int64_t* a = arena + (size_t)i * words;
    uint64_t sum = 1469598103934665603ull;                       /* FNV-1a over the row */
    for (size_t k = 0; k < words; ++k) { sum ^= (uint64_t)a[k]; sum *= 1099511628211ull; }
    printf("cell %d %d n %lld correct %lld flags %lld sum_conf %lld sum_h %lld sum_mi %lld fnv %llu\n", cells[i][0], cells[i][1],
           (long long)a[FAV_HIST_N], (long long)a[FAV_HIST_NCORRECT], (long long)a[FAV_HIST_NFLAG], (long long)a[FAV_HIST_SUM_CONF],
           (long long)a[FAV_HIST_SUM_H], (long long)a[FAV_HIST_SUM_MI], (unsigned long long)sum);
  }
  printf("launches %llu\n", (unsigned long long)fav_launch_count(h));
  CHECK(fav_destroy(h), "fav_destroy");
  return 0;
}
