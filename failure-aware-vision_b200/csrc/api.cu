// api.cu -- handle lifetime, error plumbing.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace fav {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void plan_destroy(Plan* p);   // forward.cu
}  // namespace fav

namespace fav { void comm_destroy(Ctx* ctx); void k1_cache_destroy(Ctx* ctx); int k1_scratch_release(Ctx* ctx); }
using namespace fav;

extern "C" int fav_abi_version(void) { return FAV_ABI_VERSION; }
extern "C" const char* fav_last_error(void) { return g_err; }

extern "C" int fav_init(int device, fav_handle* out) {
  FAV_REQUIRE(out, "fav_init: out is null");
  *out = nullptr;
  int count = 0;
  FAV_CUDA_OK(cudaGetDeviceCount(&count));
  FAV_REQUIRE(device >= 0 && device < count, "fav_init: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  FAV_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("fav_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback",
              device, prop.major, prop.minor);
    return FAV_E_DEVICE;
  }
  FAV_CUDA_OK(cudaSetDevice(device));
  fav_ctx* c = new fav_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  *out = c;
  return FAV_OK;
}

extern "C" int fav_reset(fav_handle h) {
  FAV_REQUIRE(h, "fav_reset: null handle");
  if (h->ws) { FAV_CUDA_OK(cudaFree(h->ws)); h->ws = nullptr; h->ws_bytes = 0; }
  return k1_scratch_release(h);          // the K1 scratch goes too; weights and the per-cell K1 tables stay
}

extern "C" int fav_destroy(fav_handle h) {
  if (!h) return FAV_OK;
  if (h->ws) cudaFree(h->ws);
  if (h->plan) plan_destroy(h->plan);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->stats_buf) cudaFree(h->stats_buf);
  if (h->splitk_buf) cudaFree(h->splitk_buf);
  comm_destroy(h);
  k1_cache_destroy(h);
  delete h;
  return FAV_OK;
}

extern "C" int fav_set_option(fav_handle h, const char* name, int value) {
  FAV_REQUIRE(h && name, "fav_set_option: null argument");
  if (strcmp(name, "splitk") == 0) { h->allow_splitk = value != 0; return FAV_OK; }
  if (strcmp(name, "k1_legacy") == 0) { h->k1_legacy = value != 0; return FAV_OK; }
  if (strcmp(name, "k1_list_stencil") == 0) { h->k1_list_stencil = value != 0; return FAV_OK; }
  set_error("fav_set_option: unknown option '%s'", name);
  return FAV_E_ARG;
}

extern "C" uint64_t fav_launch_count(fav_handle h) { return h ? h->launches : 0; }

extern "C" int fav_conv_timing_enable(fav_handle h, int on) {
  FAV_REQUIRE(h, "null handle");
  h->timing = on != 0;
  h->ev_used = 0;
  h->ev_gflop.clear();
  h->ev_gbyte.clear();
  return FAV_OK;
}

extern "C" int fav_conv_timing_read(fav_handle h, float* total_ms, int* n_launches) {
  FAV_REQUIRE(h && total_ms && n_launches, "fav_conv_timing_read: null pointer");
  float tot = 0.f;
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    FAV_CUDA_OK(cudaEventSynchronize(h->ev_pool[i + 1]));
    float ms = 0.f;
    FAV_CUDA_OK(cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *n_launches = int(h->ev_used / 2);
  h->ev_used = 0;
  h->ev_gflop.clear();
  h->ev_gbyte.clear();
  return FAV_OK;
}

extern "C" int fav_conv_timing_read_all(fav_handle h, float* ms, float* gflop, int cap, int* n_launches) {
  FAV_REQUIRE(h && ms && gflop && n_launches, "fav_conv_timing_read_all: null pointer");
  int n = 0;
  for (size_t i = 0; i + 1 < h->ev_used && n < cap; i += 2, ++n) {
    FAV_CUDA_OK(cudaEventSynchronize(h->ev_pool[i + 1]));
    FAV_CUDA_OK(cudaEventElapsedTime(&ms[n], h->ev_pool[i], h->ev_pool[i + 1]));
    gflop[n] = n < int(h->ev_gflop.size()) ? h->ev_gflop[n] : 0.f;
  }
  *n_launches = n;
  h->ev_used = 0;
  h->ev_gflop.clear();
  h->ev_gbyte.clear();
  return FAV_OK;
}

// algorithmic GB of each recorded launch; like fav_conv_stats_read it must be called BEFORE fav_conv_timing_read*
extern "C" int fav_conv_timing_read_bytes(fav_handle h, float* gbyte, int cap, int* n_launches) {
  FAV_REQUIRE(h && gbyte && n_launches, "fav_conv_timing_read_bytes: null pointer");
  int n = int(h->ev_used / 2);
  if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) gbyte[i] = i < int(h->ev_gbyte.size()) ? h->ev_gbyte[i] : 0.f;
  *n_launches = n;
  return FAV_OK;
}

// per-launch role counters (cycles summed over CTAs): {tma wait-empty, tma total, mma wait-full, mma wait-tmem-empty,
// mma total, epilogue wait-tmem-full, epilogue total, #CTAs}; valid for launches made while timing is enabled, must be
// read BEFORE fav_conv_timing_read* (which resets the launch index).
extern "C" int fav_conv_stats_read(fav_handle h, uint64_t* out, int cap_launches, int* n_launches) {
  FAV_REQUIRE(h && out && n_launches, "fav_conv_stats_read: null pointer");
  int n = int(h->ev_used / 2);
  if (n > cap_launches) n = cap_launches;
  if (n > 512) n = 512;
  *n_launches = n;
  if (n == 0 || !h->stats_buf) return FAV_OK;
  FAV_CUDA_OK(cudaDeviceSynchronize());
  FAV_CUDA_OK(cudaMemcpy(out, h->stats_buf, (size_t)n * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return FAV_OK;
}
