// tables.h -- host-side K1 parameter tables (tables.cu) shared with corrupt.cu.
#pragma once
#include <stdint.h>
#include <vector>
#include <cuda_runtime.h>

namespace fav {
struct Ctx;

struct K1Params {
  std::vector<float> fp;            // per-corruption float constants
  std::vector<int32_t> ip;          // per-corruption integer constants / table geometry
  std::vector<uint8_t> table;       // device table image (may be empty)
};
struct K1Entry {
  K1Params p;
  void* d_table = nullptr;
};

int profile_of(unsigned flags, int height, int width);                      // 0 = CIFAR-10-C constants, 1 = ImageNet-C
int corrupt_params(int corruption, int severity, int h, int w, int profile, K1Params& out);
int k1_lookup(Ctx* ctx, int corruption, int severity, int h, int w, int profile, cudaStream_t st, const K1Entry** out);
int k1_scratch(Ctx* ctx, size_t bytes, void** out);
void k1_cache_destroy(Ctx* ctx);
}  // namespace fav
