// common.cuh -- shared device helpers (Philox, bf16 packing) and the host-side context.
// sm_100a only.  RNG contract: SURVEY.md Appendix A.1; oracle twin: oracle/philox.py.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/fav_b200.h"

namespace fav {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
#define FAV_CUDA_OK(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fav::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FAV_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)
#define FAV_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      fav::set_error(__VA_ARGS__);      \
      return FAV_E_ARG;                 \
    }                                   \
  } while (0)

// entry points run on the handle's device whatever the caller's current device is
#define FAV_DEVICE(h)                                                              \
  do {                                                                             \
    int _dev = -1;                                                                 \
    if (cudaGetDevice(&_dev) != cudaSuccess || _dev != (h)->device) FAV_CUDA_OK(cudaSetDevice((h)->device)); \
  } while (0)

// ---------------------------------------------------------------- Philox4x32-10
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

constexpr int KIND_IMAGES = 1, KIND_LABELS = 2, KIND_CORRUPT = 3, KIND_DROPOUT = 4, KIND_AUX = 5;
__host__ __device__ constexpr uint32_t stream_id(int kind, int a = 0, int b = 0) {
  return (uint32_t(kind) << 16) | (uint32_t(a) << 8) | uint32_t(b);
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// (0,1] uniform: fl(fl(x>>8) * 2^-24 + 2^-25), two separately rounded fp32 ops (no FMA).
__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return __fadd_rn(__fmul_rn(float(x >> 8), 5.9604644775390625e-08f), 2.98023223876953125e-08f);
}

// Box-Muller: two u32 -> (r cos th, r sin th)
__device__ __forceinline__ float2 box_muller(uint32_t xa, uint32_t xb) {
  // MUFU-based fast math (abs error ~1e-6, far inside the 1e-3 parity bar): r = sqrt(-2 ln u1) and the angle is
  // evaluated on (-pi, pi] where __sinf / __cosf are most accurate, using cos(t + pi) = -cos t, sin(t + pi) = -sin t
  const float u1 = u32_to_uniform(xa), u2 = u32_to_uniform(xb);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  r = -r;
  const float t = fmaf(6.283185307179586f, u2, -3.141592653589793f);
  return make_float2(r * __cosf(t), r * __sinf(t));
}

// MC-dropout contract (oracle twin: oracle/model.py dropout_mask): one Philox call covers 16 consecutive NHWC channels, one
// BYTE per channel in natural order (channel c of the chunk = byte c of the 16 output bytes, x0's low byte first); a channel
// is dropped iff its byte < thr8 = round(p * 256) (p is quantised to 1/256), kept values are scaled by the exact inverse of
// the realised keep probability, fl32(256 / (256 - thr8)).
inline uint32_t dropout_thr8(float p) {
  const double t = double(p) * 256.0 + 0.5;
  return t < 0.0 ? 0u : t >= 256.0 ? 255u : uint32_t(t);
}
inline float dropout_scale8(uint32_t thr8) { return 256.0f / float(256u - thr8); }
// Operands of the four-bytes-at-a-time compare below: byte >= K is  msb | (low7 >= K)  for K <= 128 and
// msb & (low7 >= K - 128)  for K > 128; "low7 >= k" is the carry of low7 + (0x80 - k) into bit 7 (no carry leaves the byte).
inline uint32_t dropout_add4(uint32_t thr8) { return (thr8 <= 128u ? 0x80u - thr8 : 0x100u - thr8) * 0x01010101u; }
inline uint32_t dropout_hi4(uint32_t thr8) { return thr8 > 128u ? 0xFFFFFFFFu : 0u; }

// keep flags of four channels in the sign bits of the four bytes of the result (three instructions: AND, ADD, one LOP3)
__device__ __forceinline__ uint32_t dropout_keep4(uint32_t r, uint32_t add4, uint32_t hi4) {
  const uint32_t t = (r & 0x7F7F7F7Fu) + add4;
  return (t & r) | (~hi4 & (t | r));
}
// sign bits of bytes (0, 1) / (2, 3) -> 0xFFFF-per-kept-channel masks for a packed bf16 pair (PRMT sign replication)
// (inline PTX: __byte_perm only honours the low three bits of each selector nibble, bit 3 = "replicate the sign" needs prmt.b32)
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t dropout_pair_lo(uint32_t keep4) { return prmt_b32(keep4, 0u, 0x9988u); }
__device__ __forceinline__ uint32_t dropout_pair_hi(uint32_t keep4) { return prmt_b32(keep4, 0u, 0xBBAAu); }

// Philox4x32-10 with the ten round keys precomputed on the host (kernel parameters: the key schedule costs no instructions)
struct PhiloxKeys { uint32_t k0[10], k1[10]; };
inline PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
  PhiloxKeys k;
  for (int r = 0; r < 10; ++r) { k.k0[r] = k0 + uint32_t(r) * 0x9E3779B9u; k.k1[r] = k1 + uint32_t(r) * 0xBB67AE85u; }
  return k;
}
__device__ __forceinline__ uint4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k.k0[r], n2 = hi0 ^ c3 ^ k.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// ReLU on a packed bf16 pair (one HMNMX2 instead of two FMNMX before the pack; rounding to bf16 is monotone and keeps 0, so
// max(round(x), 0) == round(max(x, 0)), also after a multiplication by a positive dropout scale)
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0u));
  return d;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ---------------------------------------------------------------- host context
struct ConvLayer;   // conv.cu
struct Plan;        // forward.cu

struct Ctx {
  int device = 0;
  int num_sms = 148;
  uint64_t launches = 0;
  // weights + plan
  Plan* plan = nullptr;
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // optional per-launch timing of the conv kernel (bench roofline)
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<float> ev_gflop;
  std::vector<float> ev_gbyte;    // algorithmic bytes per launch (operands read once, result written once), in GB
  void* stats_buf = nullptr;      // 512 launches x 8 role counters (cycles)
  // fp32 accumulation scratch of the split-K convolutions (few output tiles, many k-blocks: the batch-1 streaming gate)
  void* splitk_buf = nullptr;     // [tickets int per output tile | fp32 partial tiles]
  size_t splitk_bytes = 0;
  bool allow_splitk = false;      // fav_set_option(h, "splitk", 1)
  // multi-GPU: one process per GPU, communicator for the histogram all-reduce (comm.cu)
  void* nccl_comm = nullptr;
  // K1 host tables (tables.cu): per-(corruption, severity, h, w, profile) constants + device tables, library-owned scratch
  void* k1_cache = nullptr;
  // kernel attributes (max dynamic shared memory) are per device: set once per handle, not once per process
  bool attr_conv = false, attr_flat = false, attr_pair = false, attr_shot = false, attr_plasma = false;
  bool k1_legacy = false;         // fav_set_option(h, "k1_legacy", 1): the round-1 K1 kernels (A/B measurements, fallback)
  bool k1_list_stencil = false;   // fav_set_option(h, "k1_list_stencil", 1): defocus_blur through the tap-list loop (A/B of the dense loop)
  int world = 1, rank = 0;
};

}  // namespace fav

struct fav_ctx : fav::Ctx {};
