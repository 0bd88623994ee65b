// corrupt.cu -- K1: image-corruption generators fused with normalize (uint8 NHWC -> bf16 NHWC).
//
// Replaces (reference): the two noise/brightness sliders of platform/backend/vision_simulator.py:25-36
// and the display-only JS effects of platform/frontend/js/app.js:789-799,834-851 -- the reference has
// no server-side generator.  Definitions follow oracle/corruptions.py.
//
// HBM-bound design: one thread owns a group of 16 whole pixels (48 source bytes = three 128-bit
// loads, 96 output bytes = six 128-bit stores); noise comes from counter-based Philox keyed by
// (seed, global image index, element chunk) so nothing but the source and the result touches HBM.
// Blur stencils stage the source tile plus halo in shared memory as packed RGBX words.
#include "common.cuh"
#include "tables.h"

namespace fav {

enum { PW_CLEAN = 0, PW_GAUSS, PW_SHOT, PW_IMPULSE, PW_BRIGHT, PW_CONTRAST, PW_FOG, PW_FROST, PW_SNOW };

struct PointwiseArgs {
  const uint8_t* src;
  void* dst;
  int n, per;                 // images, elements per image (h*w*3)
  int groups_per_image;       // ceil(per/48)
  uint32_t gpi_magic;         // floor(2^32 / groups_per_image) + 1 when total groups * groups_per_image < 2^32, else 0 (-> plain division)
  unsigned long long cpi_m64; // k1_chunk16: ceil(2^64 / chunks per image): umul64hi(c, m) == c / cpi for every c < 2^32 (0 -> plain division)
  uint32_t k0, k1;            // Philox key
  uint32_t first_image;
  uint32_t stream;            // Philox c3
  float f0, f1;               // corruption constants
  float tint[3];              // frost texture colour
  uint32_t u0, u1;            // integer constants (thresholds / table width)
  const void* table;          // shot: int32 kmin[256], uint32 thr[256][width], uint16 jump[256][256]
  const void* scratch;        // contrast: uint64 sums[n][3]; fog: float stats[n][4] + maps
  int hw, width, mapsize;     // fog geometry
  size_t map_offset;          // bytes from scratch to the plasma maps
  float mean[3], inv_std[3];
  unsigned flags;
};

// b / 255 for an integer b in [0,255], correctly rounded (== __fdiv_rn(b, 255)) in three FMA-pipe instructions:
// q = b*r, residual e = fma(-q, 255, b), q' = fma(e, r, q).  Verified exhaustively for all 256 inputs (tests/test_host.py).
__device__ __forceinline__ float div255(float b) {
  const float r = 0.003921568859368563f;      // RN(1/255)
  const float q = __fmul_rn(b, r);
  return __fmaf_rn(__fmaf_rn(-q, 255.0f, b), r, q);
}
__device__ __forceinline__ float u8f(uint32_t b) { return div255(float(b)); }

template <int MODE>
__global__ void __launch_bounds__(256) k1_pointwise(const PointwiseArgs a) {
  const long long total = (long long)a.n * a.groups_per_image;
  const bool vec = (a.per % 48) == 0 && ((reinterpret_cast<uintptr_t>(a.src) & 15) == 0);
  const bool bgr = a.flags & FAV_SRC_BGR;
  const bool out_f32 = a.flags & FAV_OUT_F32;
  const bool no_norm = a.flags & FAV_NO_NORMALIZE;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    // image index = g / groups_per_image: a multiply-high when the host could prove it exact (the 64-bit division it replaces
    // costs more instructions than the whole clean path of a group)
    const int img = a.gpi_magic ? (a.groups_per_image == 1 ? int(g) : int(__umulhi(uint32_t(g), a.gpi_magic))) : int(g / a.groups_per_image);
    const int gi = int(g - (long long)img * a.groups_per_image);
    const int e0 = gi * 48;
    const int cnt = min(48, a.per - e0);
    const size_t base = (size_t)img * a.per + e0;
    uint32_t w[12];
    if (vec) {
      const uint4* p = reinterpret_cast<const uint4*>(a.src + base);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint4 v = __ldg(p + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (4 * i + b < cnt) v |= uint32_t(a.src[base + 4 * i + b]) << (8 * b);
        w[i] = v;
      }
    }
    float x[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) {
      const int sb = i - (i % 3) + 2 - (i % 3);                // logical RGB element i <- BGR source sb
      const uint32_t va = (w[i >> 2] >> (8 * (i & 3))) & 0xFF, vb = (w[sb >> 2] >> (8 * (sb & 3))) & 0xFF;
      x[i] = float(bgr ? vb : va);
    }
    const uint32_t gimg = a.first_image + uint32_t(img);

    if (MODE == PW_GAUSS) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const uint4 r = philox4x32_10(uint32_t(e0 / 4 + j), gimg, 0u, a.stream, a.k0, a.k1);
        const float2 za = box_muller(r.x, r.y), zb = box_muller(r.z, r.w);
        x[4 * j + 0] = __fmaf_rn(a.f0, za.x, div255(x[4 * j + 0]));
        x[4 * j + 1] = __fmaf_rn(a.f0, za.y, div255(x[4 * j + 1]));
        x[4 * j + 2] = __fmaf_rn(a.f0, zb.x, div255(x[4 * j + 2]));
        x[4 * j + 3] = __fmaf_rn(a.f0, zb.y, div255(x[4 * j + 3]));
      }
    } else if (MODE == PW_SHOT) {
      const int* kmin = reinterpret_cast<const int*>(a.table);
      const uint32_t* thr = reinterpret_cast<const uint32_t*>(a.table) + 256;
      const int width = int(a.u0);
      const unsigned short* jump = reinterpret_cast<const unsigned short*>(thr + 256 * (size_t)width);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const uint4 r = philox4x32_10(uint32_t(e0 / 4 + j), gimg, 0u, a.stream, a.k0, a.k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int v = int(x[4 * j + q]);
          const uint32_t* row = thr + (size_t)v * width;
          // first index with row[idx] > u: start from the per-(value, top byte of u) lower bound, then probe linearly
          int lo;
          if (width <= 160) {
            lo = int(__ldg(jump + v * 256 + (rr[q] >> 24)));
            while (lo < width && __ldg(row + lo) <= rr[q]) ++lo;
          } else {                                       // large lambda: the tables spill L1, plain binary search does fewer probes
            lo = 0;
            int hi = width;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (__ldg(row + mid) <= rr[q]) lo = mid + 1; else hi = mid;
            }
          }
          x[4 * j + q] = __fdiv_rn(float(__ldg(kmin + v) + lo), a.f0);
        }
      }
    } else if (MODE == PW_IMPULSE) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const uint4 r = philox4x32_10(uint32_t(e0 / 4 + j), gimg, 0u, a.stream, a.k0, a.k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = div255(x[4 * j + q]);
          if (rr[q] < a.u1) v = 1.0f;
          if (rr[q] < a.u0) v = 0.0f;
          x[4 * j + q] = v;
        }
      }
    } else if (MODE == PW_BRIGHT) {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const float r = div255(x[3 * p]), gg = div255(x[3 * p + 1]), b = div255(x[3 * p + 2]);
        const float v = fmaxf(r, fmaxf(gg, b));
        const float v2 = fminf(__fadd_rn(v, a.f0), 1.0f);
        if (v > 0.0f) {
          const float s = __fdiv_rn(v2, v);
          x[3 * p] = r * s; x[3 * p + 1] = gg * s; x[3 * p + 2] = b * s;
        } else {
          x[3 * p] = v2; x[3 * p + 1] = v2; x[3 * p + 2] = v2;
        }
      }
    } else if (MODE == PW_CONTRAST) {
      const unsigned long long* sums = reinterpret_cast<const unsigned long long*>(a.scratch) + 3 * (size_t)img;
      float mu[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) mu[c] = __fdiv_rn(__ull2float_rn(sums[bgr ? 2 - c : c]), 255.0f * float(a.hw));
#pragma unroll
      for (int i = 0; i < 48; ++i)
        x[i] = __fadd_rn(__fmul_rn(__fsub_rn(div255(x[i]), mu[i % 3]), a.f0), mu[i % 3]);
    } else if (MODE == PW_FOG) {
      const float* st = reinterpret_cast<const float*>(a.scratch) + 4 * (size_t)img;
      const float* map = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.scratch) + a.map_offset) +
                         (size_t)img * a.mapsize * a.mapsize;
      const float mn = st[0], pmx = st[1], xmx = st[2];
      const float gain = __fdiv_rn(xmx, __fadd_rn(xmx, a.f0));
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int pix = e0 / 3 + p;
        float pl = 0.f;
        if (pix < a.hw) {
          const int yy = pix / a.width, xx = pix - yy * a.width;
          pl = __fdiv_rn(__fsub_rn(map[yy * a.mapsize + xx], mn), pmx);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
          x[3 * p + c] = __fmul_rn(__fadd_rn(div255(x[3 * p + c]), __fmul_rn(a.f0, pl)), gain);
      }
    } else if (MODE == PW_SNOW) {
      // x' = blend x + (1 - blend) max(x, 1.5 gray + 0.5);  out = x' + B[y,x] + B[H-1-y, W-1-x]   (B = blurred snow layer)
      const float* B = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.scratch) + a.map_offset) + (size_t)img * a.hw;
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int pix = e0 / 3 + p;
        const float r = div255(x[3 * p]), gg = div255(x[3 * p + 1]), b = div255(x[3 * p + 2]);
        const float gray = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, gg)), __fmul_rn(0.114f, b));
        const float lift = __fadd_rn(__fmul_rn(gray, 1.5f), 0.5f);
        float s0 = 0.f, s1 = 0.f;
        if (pix < a.hw) { s0 = B[pix]; s1 = B[a.hw - 1 - pix]; }
        const float c3[3] = {r, gg, b};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float xb = __fadd_rn(__fmul_rn(a.f0, c3[c]), __fmul_rn(a.f1, fmaxf(c3[c], lift)));
          x[3 * p + c] = __fadd_rn(__fadd_rn(xb, s0), s1);
        }
      }
    } else if (MODE == PW_FROST) {
      // procedural frost (the ImageNet-C frost photographs are not available): texture = plasma through a contrast curve,
      // tinted; out = c0 * x + c1 * texture
      const float* st = reinterpret_cast<const float*>(a.scratch) + 4 * (size_t)img;
      const float* map = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.scratch) + a.map_offset) +
                         (size_t)img * a.mapsize * a.mapsize;
      const float mn = st[0], pmx = st[1];
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int pix = e0 / 3 + p;
        float f = 0.f;
        if (pix < a.hw) {
          const int yy = pix / a.width, xx = pix - yy * a.width;
          const float pl = __fdiv_rn(__fsub_rn(map[yy * a.mapsize + xx], mn), pmx);
          f = fminf(fmaxf(__fsub_rn(__fmul_rn(1.35f, pl), 0.1f), 0.f), 1.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
          x[3 * p + c] = __fadd_rn(__fmul_rn(a.f0, div255(x[3 * p + c])), __fmul_rn(a.f1, __fmul_rn(f, a.tint[c])));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 48; ++i) x[i] = div255(x[i]);
    }

    // clip + normalize + store
#pragma unroll
    for (int i = 0; i < 48; ++i) {
      float v = fminf(fmaxf(x[i], 0.0f), 1.0f);
      if (!no_norm) v = __fmul_rn(__fsub_rn(v, a.mean[i % 3]), a.inv_std[i % 3]);
      x[i] = v;
    }
    if (out_f32) {
      float* o = reinterpret_cast<float*>(a.dst) + base;
      if (cnt == 48 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 12; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 48; ++i) if (i < cnt) o[i] = x[i];
      }
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.dst) + base;
      if (cnt == 48 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
          reinterpret_cast<uint4*>(o)[i] =
              make_uint4(pack_bf16x2(x[8 * i], x[8 * i + 1]), pack_bf16x2(x[8 * i + 2], x[8 * i + 3]),
                         pack_bf16x2(x[8 * i + 4], x[8 * i + 5]), pack_bf16x2(x[8 * i + 6], x[8 * i + 7]));
      } else {
#pragma unroll
        for (int i = 0; i < 48; ++i) if (i < cnt) o[i] = __float2bfloat16_rn(x[i]);
      }
    }
  }
}

// ---------------------------------------------------------------- per-image channel sums (contrast)
__global__ void __launch_bounds__(256) k1_channel_sums(const uint8_t* __restrict__ src, int per,
                                                       unsigned long long* __restrict__ sums) {
  const int img = blockIdx.x;
  const uint8_t* p = src + (size_t)img * per;
  unsigned int s[3] = {0, 0, 0};
  const bool vec = (per % 48) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  if (vec) {   // 16 whole pixels = three 128-bit loads per thread; byte j of the group belongs to channel j % 3
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g * 48 < per; g += gridDim.y * blockDim.x) {
      const uint4* q = reinterpret_cast<const uint4*>(p + (size_t)g * 48);
      uint32_t w[12];
#pragma unroll
      for (int i = 0; i < 3; ++i) { const uint4 v = __ldg(q + i); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#pragma unroll
      for (int j = 0; j < 48; ++j) s[j % 3] += (w[j >> 2] >> (8 * (j & 3))) & 0xFF;
    }
  } else {
    for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix * 3 < per; pix += gridDim.y * blockDim.x) {
      s[0] += p[3 * pix]; s[1] += p[3 * pix + 1]; s[2] += p[3 * pix + 2];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    unsigned int v = s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sums[3 * (size_t)img + c], (unsigned long long)v);
  }
}

// small images (<= 48 * 32 * 8 bytes, e.g. 32x32): one WARP per image, 8 images per CTA, no atomics and no pre-zeroing
__global__ void __launch_bounds__(256) k1_channel_sums_small(const uint8_t* __restrict__ src, int per, int n,
                                                             unsigned long long* __restrict__ sums) {
  const int img = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (img >= n) return;
  const uint8_t* p = src + (size_t)img * per;
  unsigned int s[3] = {0, 0, 0};
  for (int g = lane; g * 48 < per; g += 32) {
    const uint4* q = reinterpret_cast<const uint4*>(p + (size_t)g * 48);
    uint32_t w[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const uint4 v = __ldg(q + i); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#pragma unroll
    for (int j = 0; j < 48; ++j) s[j % 3] += (w[j >> 2] >> (8 * (j & 3))) & 0xFF;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    unsigned int v = s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sums[3 * (size_t)img + c] = (unsigned long long)v;
  }
}

// contrast on small images in ONE pass (per <= 48 * 64 bytes, e.g. 32x32x3; RGB in, normalised bf16 out): a warp owns an
// image, each lane keeps its (up to two) 16-pixel groups in registers, the channel sums are a warp reduction, then the
// same arithmetic as k1_pointwise<PW_CONTRAST> is applied to the registers -- the image is read once.
__global__ void __launch_bounds__(256) k1_contrast_small(const PointwiseArgs a) {
  const int img = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (img >= a.n) return;
  const uint8_t* p = a.src + (size_t)img * a.per;
  const int ngroups = a.per / 48;
  uint32_t w[2][12];
  unsigned int s[3] = {0, 0, 0};
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int g = lane + 32 * it;
#pragma unroll
    for (int i = 0; i < 12; ++i) w[it][i] = 0;
    if (g < ngroups) {
      const uint4* q = reinterpret_cast<const uint4*>(p + (size_t)g * 48);
#pragma unroll
      for (int i = 0; i < 3; ++i) { const uint4 v = __ldg(q + i); w[it][4 * i] = v.x; w[it][4 * i + 1] = v.y; w[it][4 * i + 2] = v.z; w[it][4 * i + 3] = v.w; }
#pragma unroll
      for (int j = 0; j < 48; ++j) s[j % 3] += (w[it][j >> 2] >> (8 * (j & 3))) & 0xFF;
    }
  }
  float mu[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    unsigned int v = s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    mu[c] = __fdiv_rn(__ull2float_rn((unsigned long long)v), 255.0f * float(a.hw));
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int g = lane + 32 * it;
    if (g >= ngroups) continue;
    float x[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) {
      const float b = float((w[it][i >> 2] >> (8 * (i & 3))) & 0xFF);
      float v = __fadd_rn(__fmul_rn(__fsub_rn(div255(b), mu[i % 3]), a.f0), mu[i % 3]);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      x[i] = __fmul_rn(__fsub_rn(v, a.mean[i % 3]), a.inv_std[i % 3]);
    }
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + (size_t)img * a.per + (size_t)g * 48);
#pragma unroll
    for (int i = 0; i < 6; ++i)
      o[i] = make_uint4(pack_bf16x2(x[8 * i], x[8 * i + 1]), pack_bf16x2(x[8 * i + 2], x[8 * i + 3]),
                        pack_bf16x2(x[8 * i + 4], x[8 * i + 5]), pack_bf16x2(x[8 * i + 6], x[8 * i + 7]));
  }
}

// ---------------------------------------------------------------- fog: diamond-square plasma, one CTA per image
__global__ void __launch_bounds__(1024) k1_plasma(const uint8_t* __restrict__ src, int per, int mapsize,
                                                  float decay, uint32_t k0, uint32_t k1, uint32_t first_image,
                                                  uint32_t stream, float* __restrict__ stats,
                                                  float* __restrict__ maps) {
  const int img = blockIdx.x;
  float* M = maps + (size_t)img * mapsize * mapsize;
  const uint32_t gimg = first_image + uint32_t(img);
  const int mask = mapsize - 1;
  auto noise = [&](int y, int x) {
    const uint4 r = philox4x32_10(uint32_t(y * mapsize + x), gimg, 0u, stream, k0, k1);
    return __fsub_rn(__fmul_rn(2.0f, u32_to_uniform(r.x)), 1.0f);
  };
  if (threadIdx.x == 0) M[0] = 0.0f;
  __syncthreads();
  float wib = 100.0f;
  for (int step = mapsize; step >= 2; step >>= 1) {
    const int hf = step >> 1, cells = mapsize / step;
    const float w2 = __fmul_rn(wib, wib);
    // squares
    for (int i = threadIdx.x; i < cells * cells; i += blockDim.x) {
      const int y = (i / cells) * step, x = (i % cells) * step;
      const float a = M[y * mapsize + x], b = M[((y + step) & mask) * mapsize + x];
      const float c = M[y * mapsize + ((x + step) & mask)], d = M[((y + step) & mask) * mapsize + ((x + step) & mask)];
      const float sq = __fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d));
      M[(y + hf) * mapsize + x + hf] = __fadd_rn(__fmul_rn(sq, 0.25f), __fmul_rn(w2, noise(y + hf, x + hf)));
    }
    __syncthreads();
    // diamonds
    for (int i = threadIdx.x; i < cells * cells; i += blockDim.x) {
      const int y = (i / cells) * step, x = (i % cells) * step;
      const float ul = M[y * mapsize + x];
      const float dr = M[(y + hf) * mapsize + x + hf];
      {  // M[y, x+hf] = (dr + dr_up) + (ul + ul_right)
        const float dru = M[((y - hf) & mask) * mapsize + x + hf];
        const float ulr = M[y * mapsize + ((x + step) & mask)];
        const float lt = __fadd_rn(__fadd_rn(dr, dru), __fadd_rn(ul, ulr));
        M[y * mapsize + x + hf] = __fadd_rn(__fmul_rn(lt, 0.25f), __fmul_rn(w2, noise(y, x + hf)));
      }
      {  // M[y+hf, x] = (dr + dr_left) + (ul + ul_down)
        const float drl = M[(y + hf) * mapsize + ((x - hf) & mask)];
        const float uld = M[((y + step) & mask) * mapsize + x];
        const float tt = __fadd_rn(__fadd_rn(dr, drl), __fadd_rn(ul, uld));
        M[(y + hf) * mapsize + x] = __fadd_rn(__fmul_rn(tt, 0.25f), __fmul_rn(w2, noise(y + hf, x)));
      }
    }
    __syncthreads();
    wib = __fdiv_rn(wib, decay);
  }
  // min / max of the map, max of the image
  __shared__ float s_mn[32], s_mx[32];
  __shared__ unsigned s_xm[32];
  float mn = 3.4e38f, mx = -3.4e38f;
  for (int i = threadIdx.x; i < mapsize * mapsize; i += blockDim.x) {
    const float v = M[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  unsigned xm = 0;
  const uint8_t* p = src + (size_t)img * per;
  for (int i = threadIdx.x; i < per; i += blockDim.x) xm = max(xm, (unsigned)p[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    xm = max(xm, __shfl_xor_sync(0xffffffffu, xm, o));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; s_xm[threadIdx.x >> 5] = xm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < int(blockDim.x >> 5); ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); xm = max(xm, s_xm[i]); }
    stats[4 * img + 0] = mn;
    stats[4 * img + 1] = __fsub_rn(mx, mn);
    stats[4 * img + 2] = __fdiv_rn(float(xm), 255.0f);
    stats[4 * img + 3] = 0.f;
  }
}

// ---------------------------------------------------------------- element emitters of the fast pointwise kernels
// clip to [0,1], normalize (unless FAV_NO_NORMALIZE) and store CNT consecutive elements that start at element `base`
// (channel of element i = (c0 + i) % 3).  CNT = 16: base is a multiple of 16; CNT = 12: base is a multiple of 12.
template <int CNT>
__device__ __forceinline__ void emit_elems(void* dst, size_t base, int c0, float* v, const float* mean, const float* inv_std,
                                           unsigned flags) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    float t = fminf(fmaxf(v[i], 0.0f), 1.0f);
    if (!(flags & FAV_NO_NORMALIZE)) t = __fmul_rn(__fsub_rn(t, mean[(c0 + i) % 3]), inv_std[(c0 + i) % 3]);
    v[i] = t;
  }
  if (flags & FAV_OUT_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + base);     // 16-byte aligned for both CNTs
#pragma unroll
    for (int i = 0; i < CNT / 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if (CNT == 16) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + base);
#pragma unroll
    for (int i = 0; i < 2; ++i)
      o[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                        pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  } else {
    uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + base);   // 24-byte chunks: 8-byte aligned
#pragma unroll
    for (int i = 0; i < CNT / 4; ++i)
      o[i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
  }
}

// ---------------------------------------------------------------- element-wise modes, one 16-element chunk per thread
// clean / gaussian_noise / impulse_noise do not couple the three channels of a pixel, so a thread can own ONE 16-byte chunk
// (one 128-bit load, two 128-bit bf16 stores): consecutive threads touch consecutive chunks (fully coalesced, instead of
// 16-byte accesses at a 48- / 96-byte stride) and the kernel needs ~40 registers instead of 68-74 (twice the resident
// warps).  Same arithmetic and the same Philox counters as k1_pointwise (chunk c of an image = Philox calls 4c .. 4c+3):
// bit-identical results.
template <int MODE>
__global__ void __launch_bounds__(256) k1_chunk16(const PointwiseArgs a) {
  const int cpi = a.per >> 4;                                    // chunks per image (per % 16 == 0, host-checked)
  const long long total = (long long)a.n * cpi;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
    // image index = c / cpi: a multiply-high when the host could prove it exact (a 64-bit division costs as much as the
    // whole clean path of a chunk)
    const int img = a.cpi_m64 ? int(__umul64hi((unsigned long long)c, a.cpi_m64)) : int(c / cpi);
    const int cc = int(c - (long long)img * cpi);
    const size_t base = (size_t)img * a.per + (size_t)cc * 16;
    const uint32_t gimg = a.first_image + uint32_t(img);
    const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(a.src + base));
    const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
    float x[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = div255(float((w[j] >> (8 * q)) & 0xFFu));
      if (MODE == PW_GAUSS) {
        const uint4 r = philox4x32_10(uint32_t(4 * cc + j), gimg, 0u, a.stream, a.k0, a.k1);
        const float2 za = box_muller(r.x, r.y), zb = box_muller(r.z, r.w);
        b[0] = __fmaf_rn(a.f0, za.x, b[0]); b[1] = __fmaf_rn(a.f0, za.y, b[1]);
        b[2] = __fmaf_rn(a.f0, zb.x, b[2]); b[3] = __fmaf_rn(a.f0, zb.y, b[3]);
      } else if (MODE == PW_IMPULSE) {
        const uint4 r = philox4x32_10(uint32_t(4 * cc + j), gimg, 0u, a.stream, a.k0, a.k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (rr[q] < a.u1) b[q] = 1.0f;
          if (rr[q] < a.u0) b[q] = 0.0f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) x[4 * j + q] = b[q];
    }
    // channel of element i = (16 cc + i) % 3 = (cc + i) % 3: rotate the per-channel constants once instead of indexing them
    const int c0 = cc % 3;
    if (MODE == PW_CONTRAST) {                                   // (x - mu_c) * c + mu_c with the per-image channel means of the pre-pass
      const unsigned long long* sums = reinterpret_cast<const unsigned long long*>(a.scratch) + 3 * (size_t)img;
      float mu[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) mu[c] = __fdiv_rn(__ull2float_rn(__ldg(sums + c)), 255.0f * float(a.hw));
      const float mr[3] = {c0 == 0 ? mu[0] : (c0 == 1 ? mu[1] : mu[2]), c0 == 0 ? mu[1] : (c0 == 1 ? mu[2] : mu[0]),
                           c0 == 0 ? mu[2] : (c0 == 1 ? mu[0] : mu[1])};
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __fadd_rn(__fmul_rn(__fsub_rn(x[i], mr[i % 3]), a.f0), mr[i % 3]);
    }
    const float m[3] = {c0 == 0 ? a.mean[0] : (c0 == 1 ? a.mean[1] : a.mean[2]), c0 == 0 ? a.mean[1] : (c0 == 1 ? a.mean[2] : a.mean[0]),
                        c0 == 0 ? a.mean[2] : (c0 == 1 ? a.mean[0] : a.mean[1])};
    const float is[3] = {c0 == 0 ? a.inv_std[0] : (c0 == 1 ? a.inv_std[1] : a.inv_std[2]),
                         c0 == 0 ? a.inv_std[1] : (c0 == 1 ? a.inv_std[2] : a.inv_std[0]),
                         c0 == 0 ? a.inv_std[2] : (c0 == 1 ? a.inv_std[0] : a.inv_std[1])};
    emit_elems<16>(a.dst, base, 0, x, m, is, a.flags);
  }
}

// ---------------------------------------------------------------- shot noise with the guide table in shared memory
// Same definition as k1_pointwise<PW_SHOT> (k = kmin[v] + #{j : thr[v][j] <= u}, u = one Philox word per element), different
// search: guide[v][u >> 24] (uint16: k for the draw (b << 24) in the low 10 bits, number of thresholds that fall inside the
// top-byte cell in the high 6, saturated) answers ~80 % of the draws with ONE shared-memory read; only draws whose cell
// contains a threshold probe the 32-bit thresholds (global, L2-resident).  k / c is the exact division by the 3-FMA
// residual correction (verified for every k < 1024 and every c of the severity tables, tests/test_host.py).
struct ShotArgs {
  const uint8_t* src;
  void* dst;
  int n, per, groups_per_image;
  uint32_t k0, k1, first_image, stream;
  float c, rc;
  int width;
  const int* kmin;
  const uint32_t* thr;
  const uint16_t* guide;
  float mean[3], inv_std[3];
  unsigned flags;
};
__global__ void __launch_bounds__(1024) k1_shot_smem(const ShotArgs a) {
  extern __shared__ __align__(16) uint16_t s_guide[];          // [256][256], then int kmin[256]
  int* s_kmin = reinterpret_cast<int*>(s_guide + 65536);
  {
    const uint4* g4 = reinterpret_cast<const uint4*>(a.guide);
    uint4* s4 = reinterpret_cast<uint4*>(s_guide);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s4[i] = __ldg(g4 + i);
    if (threadIdx.x < 256) s_kmin[threadIdx.x] = __ldg(a.kmin + threadIdx.x);
  }
  __syncthreads();
  const long long total = (long long)a.n * a.groups_per_image;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int img = int(g / a.groups_per_image);
    const int gi = int(g - (long long)img * a.groups_per_image);
    const int e0 = gi * 48;                                      // per % 48 == 0 (host-checked): whole groups only
    const size_t base = (size_t)img * a.per + e0;
    const uint32_t gimg = a.first_image + uint32_t(img);
    const uint4* p = reinterpret_cast<const uint4*>(a.src + base);
#pragma unroll
    for (int q3 = 0; q3 < 3; ++q3) {
      const uint4 v4 = __ldg(p + q3);
      const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
      float x[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 r = philox4x32_10(uint32_t(e0 / 4 + 4 * q3 + j), gimg, 0u, a.stream, a.k0, a.k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        int k4[4], idx4[4];
        uint32_t c4[4], t[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t b = (w[j] >> (8 * q)) & 0xFFu;
          const uint32_t gg = s_guide[b * 256u + (rr[q] >> 24)];
          k4[q] = int(gg & 1023u);
          c4[q] = gg >> 10;                                      // thresholds inside this draw's top-byte cell (saturated at 63)
          idx4[q] = int(b) * a.width + k4[q] - s_kmin[b];
        }
        // up to three thresholds of the cell as PREDICATED, mutually independent loads (no divergent loop, twelve loads in
        // flight per thread): ~80 % of the draws need none.  Cells b >= 1 count UP from the cell's start (probe idx + i),
        // cell 0 counts DOWN from its end (probe idx - 1 - i): see the guide table in tables.cu.
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool down = (rr[q] >> 24) == 0u;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            t[q][i] = c4[q] > uint32_t(i) ? __ldg(a.thr + (down ? idx4[q] - 1 - i : idx4[q] + i)) : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t u = rr[q];
          const bool down = (u >> 24) == 0u;
          int step = 0;                                          // sorted thresholds: the passing probes form a prefix
#pragma unroll
          for (int i = 0; i < 3; ++i) step += (c4[q] > uint32_t(i) && (down ? t[q][i] > u : t[q][i] <= u)) ? 1 : 0;
          int k = down ? k4[q] - step : k4[q] + step;
          if (c4[q] > 3u && step == 3) {                         // rare: more than three thresholds of the cell on the draw's side
            const int row0 = int((w[j] >> (8 * q)) & 0xFFu) * a.width;
            if (down) {
              int pp = idx4[q] - 4;
              while (pp >= row0 && __ldg(a.thr + pp) > u) { --pp; --k; }
            } else {
              int pp = idx4[q] + 3;
              while (pp < row0 + a.width && __ldg(a.thr + pp) <= u) { ++pp; ++k; }
            }
          }
          const float kf = float(k), qf = __fmul_rn(kf, a.rc);
          x[4 * j + q] = __fmaf_rn(__fmaf_rn(-qf, a.c, kf), a.rc, qf);         // == __fdiv_rn(kf, c)
        }
      }
      emit_elems<16>(a.dst, base + 16 * q3, q3 % 3, x, a.mean, a.inv_std, a.flags);
    }
  }
}

// ---------------------------------------------------------------- fog / frost on small frames: plasma in shared memory
// One WARP per image: the diamond-square map (mapsize <= 64) lives in shared memory, the levels are separated by
// __syncwarp, min / max / image max are warp reductions, and the same warp then applies the map to its image four pixels
// at a time -- one kernel, the image is read twice (max, apply) and written once, nothing else touches global memory.
// Arithmetic and operation order are those of k1_plasma + k1_pointwise<PW_FOG / PW_FROST> (bit-identical results).
struct PlasmaFusedArgs {
  const uint8_t* src;
  void* dst;
  int n, h, w, mapsize, frost;
  float f0, f1, decay;
  float tint[3];
  uint32_t k0, k1, first_image, stream;
  float mean[3], inv_std[3];
  unsigned flags;
};
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k1_plasma_fused(const PlasmaFusedArgs a) {
  extern __shared__ float s_maps[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x * WARPS + warp;
  if (img >= a.n) return;                                       // whole warps leave: no block-wide barrier below
  const int ms = a.mapsize, mask = ms - 1, hw = a.h * a.w, per = hw * 3;
  float* M = s_maps + (size_t)warp * ms * ms;
  const uint32_t gimg = a.first_image + uint32_t(img);
  auto noise = [&](int y, int x) {
    const uint4 r = philox4x32_10(uint32_t(y * ms + x), gimg, 0u, a.stream, a.k0, a.k1);
    return __fsub_rn(__fmul_rn(2.0f, u32_to_uniform(r.x)), 1.0f);
  };
  if (lane == 0) M[0] = 0.0f;
  __syncwarp();
  float wib = 100.0f;
  for (int step = ms; step >= 2; step >>= 1) {
    const int hf = step >> 1, cells = ms / step;
    const float w2 = __fmul_rn(wib, wib);
    for (int i = lane; i < cells * cells; i += 32) {            // squares
      const int y = (i / cells) * step, x = (i % cells) * step;
      const float c00 = M[y * ms + x], c10 = M[((y + step) & mask) * ms + x];
      const float c01 = M[y * ms + ((x + step) & mask)], c11 = M[((y + step) & mask) * ms + ((x + step) & mask)];
      const float sq = __fadd_rn(__fadd_rn(c00, c10), __fadd_rn(c01, c11));
      M[(y + hf) * ms + x + hf] = __fadd_rn(__fmul_rn(sq, 0.25f), __fmul_rn(w2, noise(y + hf, x + hf)));
    }
    __syncwarp();
    for (int i = lane; i < cells * cells; i += 32) {            // diamonds
      const int y = (i / cells) * step, x = (i % cells) * step;
      const float ul = M[y * ms + x], dr = M[(y + hf) * ms + x + hf];
      const float dru = M[((y - hf) & mask) * ms + x + hf], ulr = M[y * ms + ((x + step) & mask)];
      const float lt = __fadd_rn(__fadd_rn(dr, dru), __fadd_rn(ul, ulr));
      const float drl = M[(y + hf) * ms + ((x - hf) & mask)], uld = M[((y + step) & mask) * ms + x];
      const float tt = __fadd_rn(__fadd_rn(dr, drl), __fadd_rn(ul, uld));
      M[y * ms + x + hf] = __fadd_rn(__fmul_rn(lt, 0.25f), __fmul_rn(w2, noise(y, x + hf)));
      M[(y + hf) * ms + x] = __fadd_rn(__fmul_rn(tt, 0.25f), __fmul_rn(w2, noise(y + hf, x)));
    }
    __syncwarp();
    wib = __fdiv_rn(wib, a.decay);
  }
  float mn = 3.4e38f, mx = -3.4e38f;
  for (int i = lane; i < ms * ms; i += 32) { const float v = M[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  unsigned xm = 0;
  const uint8_t* p = a.src + (size_t)img * per;
  if ((per & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) {
    const uint32_t* p4 = reinterpret_cast<const uint32_t*>(p);
    for (int i = lane; i < (per >> 2); i += 32) xm = __vmaxu4(xm, __ldg(p4 + i));
    xm = max(max(xm & 0xFFu, (xm >> 8) & 0xFFu), max((xm >> 16) & 0xFFu, xm >> 24));
  } else {
    for (int i = lane; i < per; i += 32) xm = max(xm, (unsigned)p[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    xm = max(xm, __shfl_xor_sync(0xffffffffu, xm, o));
  }
  const float pmx = __fsub_rn(mx, mn), xmx = __fdiv_rn(float(xm), 255.0f);
  const float gain = __fdiv_rn(xmx, __fadd_rn(xmx, a.f0));
  const bool quad_ok = (reinterpret_cast<uintptr_t>(p) & 3) == 0 && !(a.flags & FAV_SRC_BGR);
  const int nq = (hw + 3) >> 2;
  for (int q = lane; q < nq; q += 32) {
    const int pix0 = 4 * q, npx = min(4, hw - pix0);
    uint32_t w3[3] = {0, 0, 0};
    if (quad_ok && npx == 4) {
      const uint32_t* p4 = reinterpret_cast<const uint32_t*>(p + 3 * (size_t)pix0);
      w3[0] = __ldg(p4); w3[1] = __ldg(p4 + 1); w3[2] = __ldg(p4 + 2);
    } else {
#pragma unroll
      for (int e = 0; e < 12; ++e) {                             // compile-time indices: w3 / v stay in registers
        const int se = (a.flags & FAV_SRC_BGR) ? e - (e % 3) + 2 - (e % 3) : e;
        if (e < 3 * npx) w3[e >> 2] |= uint32_t(p[3 * (size_t)pix0 + se]) << (8 * (e & 3));
      }
    }
    float v[12];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      const int pix = min(pix0 + pp, hw - 1);
      const int yy = pix / a.w, xx = pix - yy * a.w;
      const float pl = __fdiv_rn(__fsub_rn(M[yy * ms + xx], mn), pmx);
      const float f = fminf(fmaxf(__fsub_rn(__fmul_rn(1.35f, pl), 0.1f), 0.f), 1.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e = 3 * pp + c;
        const float xb = div255(float((w3[e >> 2] >> (8 * (e & 3))) & 0xFFu));
        v[e] = a.frost ? __fadd_rn(__fmul_rn(a.f0, xb), __fmul_rn(a.f1, __fmul_rn(f, a.tint[c])))
                       : __fmul_rn(__fadd_rn(xb, __fmul_rn(a.f0, pl)), gain);
      }
    }
    const size_t base = (size_t)img * per + 3 * (size_t)pix0;
    const bool aligned = (a.flags & FAV_OUT_F32) ? ((reinterpret_cast<uintptr_t>(a.dst) & 15) == 0 && (per & 3) == 0)
                                                 : ((reinterpret_cast<uintptr_t>(a.dst) & 7) == 0 && (per & 3) == 0);
    if (npx == 4 && aligned) {
      emit_elems<12>(a.dst, base, 0, v, a.mean, a.inv_std, a.flags);
    } else {
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        if (e >= 3 * npx) continue;
        float t = fminf(fmaxf(v[e], 0.0f), 1.0f);
        if (!(a.flags & FAV_NO_NORMALIZE)) t = __fmul_rn(__fsub_rn(t, a.mean[e % 3]), a.inv_std[e % 3]);
        if (a.flags & FAV_OUT_F32) reinterpret_cast<float*>(a.dst)[base + e] = t;
        else reinterpret_cast<__nv_bfloat16*>(a.dst)[base + e] = __float2bfloat16_rn(t);
      }
    }
  }
}

// ---------------------------------------------------------------- shared store helper for the gather kernels
struct OutArgs {
  void* dst;
  float mean[3], inv_std[3];
  unsigned flags;
};
__device__ __forceinline__ void store_pixel(const OutArgs& o, size_t pix, float r, float g, float b) {
  float v[3] = {r, g, b};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float t = fminf(fmaxf(v[c], 0.0f), 1.0f);
    if (!(o.flags & FAV_NO_NORMALIZE)) t = __fmul_rn(__fsub_rn(t, o.mean[c]), o.inv_std[c]);
    v[c] = t;
  }
  if (o.flags & FAV_OUT_F32) {
    float* d = reinterpret_cast<float*>(o.dst) + 3 * pix;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
  } else {
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(o.dst) + 3 * pix;
    d[0] = __float2bfloat16_rn(v[0]); d[1] = __float2bfloat16_rn(v[1]); d[2] = __float2bfloat16_rn(v[2]);
  }
}

// ---------------------------------------------------------------- tap-list stencil (defocus, motion)
// table: n_entries x [ int32 ntaps, pad[3], then max_taps x {int16 dy, int16 dx, float w} ]
struct TapArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w;
  int n_entries, max_taps, border;       // border 0 = reflect101, 1 = clamp
  int dy_min, dy_max, dx_min, dx_max;
  const uint8_t* table;
  uint32_t k0, k1, first_image, stream;  // entry = Philox(AUX).x % n_entries when n_entries > 1
  int tiles_x, tiles_y;
  unsigned src_bgr;
  int raw_stage;                         // 1: one tile per image, raw image staged in shared memory with 16-byte loads
  int out_staged;                        // 1: bf16 rows are 16-byte aligned -> coalesced stores through shared memory
  int dense;                             // 1: one entry whose taps fill most of their box -> register-tiled dense loop
};
constexpr int TAP_TILE = 32;

__device__ __forceinline__ int border_idx(int i, int n, int mode) {
  if (mode == 1) return min(max(i, 0), n - 1);
  i = abs(i);
  if (i >= n) i = 2 * (n - 1) - i;
  i = abs(i);
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256) k1_taps(const TapArgs a) {
  // the tile + halo is staged as one float4 (r, g, b, -) per pixel: the byte -> float conversion happens once per staged
  // pixel instead of once per (tap, output pixel), and a tap costs one 16-byte shared load + three FMAs per output pixel
  extern __shared__ float4 s_tile[];
  const int SW = TAP_TILE + a.dx_max - a.dx_min, SH = TAP_TILE + a.dy_max - a.dy_min;
  // dense mode: the weights of the tap box, replicated per output row of a thread: s_w4[dyy * BW + dx] = {w[dyy][dx],
  // w[dyy-1][dx], w[dyy-2][dx], w[dyy-3][dx]} (zero outside the box), so that one staged pixel feeds four vertically adjacent
  // outputs with one 16-byte broadcast load of weights
  const int BW = a.dx_max - a.dx_min + 1, BH = a.dy_max - a.dy_min + 1;
  float4* s_w4 = s_tile + SW * SH;
  const int n_w4 = a.dense ? (BH + 3) * BW : 0;
  uint2* s_taps = reinterpret_cast<uint2*>(s_w4 + n_w4);
  const int img = blockIdx.x;
  const int ty = blockIdx.y / a.tiles_x, tx = blockIdx.y - ty * a.tiles_x;
  const int y0 = ty * TAP_TILE, x0 = tx * TAP_TILE;
  int entry = 0;
  if (a.n_entries > 1) {
    const uint4 r = philox4x32_10(0u, a.first_image + uint32_t(img), 0u, a.stream, a.k0, a.k1);
    entry = int(r.x % uint32_t(a.n_entries));
  }
  const uint8_t* ent = a.table + (size_t)entry * (16 + 8 * (size_t)a.max_taps);
  const int ntaps = *reinterpret_cast<const int*>(ent);
  for (int i = threadIdx.x; i < n_w4; i += blockDim.x) s_w4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int disorder = 0;                        // the dense loop reproduces the list order only if the list is row-major
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) {
    const uint2 t = reinterpret_cast<const uint2*>(ent + 16)[i];
    const int dy = int(short(t.x & 0xFFFF)), dx = int(short(t.x >> 16));
    s_taps[i] = make_uint2(uint32_t((dy - a.dy_min) * SW + (dx - a.dx_min)), t.y);
    if (a.dense && i > 0) {
      const uint32_t q = reinterpret_cast<const uint2*>(ent + 16)[i - 1].x;
      const int py = int(short(q & 0xFFFF)), px = int(short(q >> 16));
      disorder |= (py > dy) || (py == dy && px >= dx);
    }
  }
  const bool dense = a.dense && !__syncthreads_or(disorder);   // also orders the zero fill before the scatter below
  if (dense) {
    float* w4 = reinterpret_cast<float*>(s_w4);
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) {
      const uint2 t = reinterpret_cast<const uint2*>(ent + 16)[i];
      const int ddy = int(short(t.x & 0xFFFF)) - a.dy_min, ddx = int(short(t.x >> 16)) - a.dx_min;
#pragma unroll
      for (int j = 0; j < 4; ++j) w4[((ddy + j) * BW + ddx) * 4 + j] = __uint_as_float(t.y);
    }
  }
  const uint8_t* p = a.src + (size_t)img * a.h * a.w * 3;
  // source row / column of every halo row / column (border rule resolved once per row and column, not once per staged pixel)
  int* s_ymap = reinterpret_cast<int*>(s_taps + a.max_taps);
  int* s_xmap = s_ymap + SH;
  for (int i = threadIdx.x; i < SH + SW; i += blockDim.x) {
    if (i < SH) s_ymap[i] = border_idx(y0 + i + a.dy_min, a.h, a.border);
    else s_xmap[i - SH] = border_idx(x0 + (i - SH) + a.dx_min, a.w, a.border);
  }
  if (a.raw_stage) {
    // small frames (one tile per image): the raw image comes in with coalesced 16-byte loads and the halo is expanded from
    // shared memory (the byte gathers of the generic path cost more than the stencil itself at 32x32)
    uint4* s_raw = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(s_xmap + SW) + 15) & ~uintptr_t(15));
    const int nv = (a.h * a.w * 3) >> 4;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) s_raw[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
    p = reinterpret_cast<const uint8_t*>(s_raw);
  }
  __syncthreads();
  const float inv_sw = 1.0f / float(SW);
  for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) {
    int sy = __float2int_rd(__fmul_rn(float(i) + 0.5f, inv_sw));          // i / SW without an integer division (exact: i < 2^20)
    int sx = i - sy * SW;
    if (sx < 0) { sx += SW; --sy; } else if (sx >= SW) { sx -= SW; ++sy; }
    const uint8_t* q = p + ((size_t)s_ymap[sy] * a.w + s_xmap[sx]) * 3;
    uint32_t c0 = q[0], c1 = q[1], c2 = q[2];
    if (a.src_bgr) { const uint32_t t = c0; c0 = c2; c2 = t; }
    s_tile[i] = make_float4(u8f(c0), u8f(c1), u8f(c2), 0.f);              // x / 255 once per staged pixel: no division per output
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly0 = (threadIdx.x >> 5) * 4;   // a warp owns four adjacent rows: 4 pixels per thread
  float acc[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
  if (dense) {
    // output j of the thread takes tap row ddy from halo row ly0 + j + ddy: walking the halo rows dyy = j + ddy in order
    // visits every output's taps in list (row-major) order; the zero weights of the padding add exactly nothing
    for (int dyy = 0; dyy < BH + 3; ++dyy) {
      const float4* trow = s_tile + (ly0 + dyy) * SW + lx;
      const float4* wrow = s_w4 + dyy * BW;
      for (int dx = 0; dx < BW; ++dx) {
        const float4 v = trow[dx];
        const float4 wq = wrow[dx];
        acc[0][0] = fmaf(wq.x, v.x, acc[0][0]); acc[0][1] = fmaf(wq.x, v.y, acc[0][1]); acc[0][2] = fmaf(wq.x, v.z, acc[0][2]);
        acc[1][0] = fmaf(wq.y, v.x, acc[1][0]); acc[1][1] = fmaf(wq.y, v.y, acc[1][1]); acc[1][2] = fmaf(wq.y, v.z, acc[1][2]);
        acc[2][0] = fmaf(wq.z, v.x, acc[2][0]); acc[2][1] = fmaf(wq.z, v.y, acc[2][1]); acc[2][2] = fmaf(wq.z, v.z, acc[2][2]);
        acc[3][0] = fmaf(wq.w, v.x, acc[3][0]); acc[3][1] = fmaf(wq.w, v.y, acc[3][1]); acc[3][2] = fmaf(wq.w, v.z, acc[3][2]);
      }
    }
  } else {
    for (int t = 0; t < ntaps; ++t) {
      const uint2 tp = s_taps[t];
      const float wgt = __uint_as_float(tp.y);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = s_tile[(ly0 + j) * SW + lx + tp.x];
        acc[j][0] = fmaf(wgt, v.x, acc[j][0]);
        acc[j][1] = fmaf(wgt, v.y, acc[j][1]);
        acc[j][2] = fmaf(wgt, v.z, acc[j][2]);
      }
    }
  }
  // bf16 output of a full-width tile whose rows are 16-byte aligned: stage the 32 x 32 x 3 results in shared memory (over
  // the halo, which is dead now) and write each 192-byte row with 16-byte stores instead of 2-byte stores at a 6-byte stride
  const bool staged = a.out_staged && x0 + TAP_TILE <= a.w;
  if (staged) {
    __syncthreads();                                           // every thread is done reading the halo
    __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(s_tile);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = fminf(fmaxf(acc[j][c], 0.0f), 1.0f);
        if (!(a.out.flags & FAV_NO_NORMALIZE)) t = __fmul_rn(__fsub_rn(t, a.out.mean[c]), a.out.inv_std[c]);
        s_out[((ly0 + j) * TAP_TILE + lx) * 3 + c] = __float2bfloat16_rn(t);
      }
    }
    __syncthreads();
    const int rows = min(TAP_TILE, a.h - y0);
    for (int i = threadIdx.x; i < rows * 12; i += blockDim.x) {            // 12 x 16 bytes per tile row
      const int r = i / 12, k = i - r * 12;
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out.dst) + (((size_t)img * a.h + y0 + r) * a.w + x0) * 3) + k;
      *o = reinterpret_cast<const uint4*>(s_out + (size_t)r * TAP_TILE * 3)[k];
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = y0 + ly0 + j, x = x0 + lx;
    if (y < a.h && x < a.w)
      store_pixel(a.out, ((size_t)img * a.h + y) * a.w + x, acc[j][0], acc[j][1], acc[j][2]);
  }
}

// ---------------------------------------------------------------- zoom blur (bilinear gathers)
// table: nz x [ (h + w) x {int16 i0, int16 i1, float frac} ]  rows first then columns
struct ZoomArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w, nz;
  const uint2* table;
  unsigned src_bgr;
};
__global__ void __launch_bounds__(256) k1_zoom(const ZoomArgs a) {
  // bilinear interpolation is linear: the zoom layers are interpolated on the raw byte values (exact small integers in
  // fp32) and the sum is divided by 255 (nz + 1) once per output -- 36 fewer FMAs per (layer, pixel) than converting every
  // gathered byte to x / 255 first; the result differs from the oracle's order of operations by a few fp32 ulp
  const long long total = (long long)a.n * a.h * a.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.h * a.w));
    const int rem = int(i - (long long)img * a.h * a.w);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint8_t* p = a.src + (size_t)img * a.h * a.w * 3;
    float acc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = float(p[((size_t)y * a.w + x) * 3 + c]);
    for (int z = 0; z < a.nz; ++z) {
      const uint2 ry = __ldg(a.table + (size_t)z * (a.h + a.w) + y);
      const uint2 rx = __ldg(a.table + (size_t)z * (a.h + a.w) + a.h + x);
      const int y0 = int(ry.x & 0xFFFF), y1 = int(ry.x >> 16), x0 = int(rx.x & 0xFFFF), x1 = int(rx.x >> 16);
      const float fy = __uint_as_float(ry.y), fx = __uint_as_float(rx.y);
      const float gx = 1.0f - fx, gy = 1.0f - fy;
      const uint8_t* q00 = p + ((size_t)y0 * a.w + x0) * 3;
      const uint8_t* q01 = p + ((size_t)y0 * a.w + x1) * 3;
      const uint8_t* q10 = p + ((size_t)y1 * a.w + x0) * 3;
      const uint8_t* q11 = p + ((size_t)y1 * a.w + x1) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float top = fmaf(float(q01[c]), fx, float(q00[c]) * gx);
        const float bot = fmaf(float(q11[c]), fx, float(q10[c]) * gx);
        acc[c] += fmaf(bot, fy, top * gy);
      }
    }
    const float den = 255.0f * float(a.nz + 1);
    float r = __fdiv_rn(acc[0], den), g = __fdiv_rn(acc[1], den), b = __fdiv_rn(acc[2], den);
    if (a.src_bgr) { const float t = r; r = b; b = t; }
    store_pixel(a.out, (size_t)i, r, g, b);
  }
}

// ---------------------------------------------------------------- pixelate (Pillow BOX down-sample, nearest up-sample)
// PIL: x.resize((int(w c), int(h c)), BOX).resize((w, h), BOX).  Pillow resamples in two passes (horizontal first) with
// 22-bit fixed-point coefficients and a uint8 intermediate (src/libImaging/Resample.c: precompute_coeffs,
// normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc); both passes are restated exactly, so the bytes
// equal PIL's.  BOX up-sampling has a single tap of weight 1: a gather.
// table (int32): sw x {xmin, cnt, k[kh]} | sh x {ymin, cnt, k[kv]} | upx[w] | upy[h]
struct PixArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w, sw, sh, kh, kv;
  const int* table;
  uint8_t* small;                // scratch: [n][sh][sw][3]
  unsigned src_bgr;
};
__global__ void __launch_bounds__(256) k1_pixelate_down(const PixArgs a) {
  const long long total = (long long)a.n * a.sh * a.sw;
  const int* hx = a.table;
  const int* vy = a.table + (size_t)a.sw * (2 + a.kh);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.sh * a.sw));
    const int rem = int(i - (long long)img * a.sh * a.sw);
    const int j = rem / a.sw, ii = rem - j * a.sw;
    const int* rx = hx + (size_t)ii * (2 + a.kh);
    const int* ry = vy + (size_t)j * (2 + a.kv);
    const int xmin = __ldg(rx), xcnt = __ldg(rx + 1), ymin = __ldg(ry), ycnt = __ldg(ry + 1);
    const uint8_t* p = a.src + (size_t)img * a.h * a.w * 3;
    int acc[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int r = 0; r < ycnt; ++r) {
      const uint8_t* row = p + ((size_t)(ymin + r) * a.w + xmin) * 3;
      int hs[3] = {1 << 21, 1 << 21, 1 << 21};
      for (int t = 0; t < xcnt; ++t) {
        const int k = __ldg(rx + 2 + t);
        hs[0] += k * int(row[3 * t]); hs[1] += k * int(row[3 * t + 1]); hs[2] += k * int(row[3 * t + 2]);
      }
      const int kvv = __ldg(ry + 2 + r);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += kvv * min(max(hs[c] >> 22, 0), 255);          // clip8 of the horizontal pass
    }
    uint8_t* o = a.small + (size_t)i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = uint8_t(min(max(acc[c] >> 22, 0), 255));
  }
}
__global__ void __launch_bounds__(256) k1_pixelate_up(const PixArgs a) {
  const long long total = (long long)a.n * a.h * a.w;
  const int* upx = a.table + (size_t)a.sw * (2 + a.kh) + (size_t)a.sh * (2 + a.kv);
  const int* upy = upx + a.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.h * a.w));
    const int rem = int(i - (long long)img * a.h * a.w);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint8_t* q = a.small + (((size_t)img * a.sh + __ldg(upy + y)) * a.sw + __ldg(upx + x)) * 3;
    float v[3] = {u8f(q[0]), u8f(q[1]), u8f(q[2])};
    if (a.src_bgr) { const float t = v[0]; v[0] = v[2]; v[2] = t; }
    store_pixel(a.out, (size_t)i, v[0], v[1], v[2]);
  }
}

// ---------------------------------------------------------------- snow: layer generation and its motion blur
// Layer: 8-term Irwin-Hall field (one Philox call per pixel, integer sum -> bit-exact) -> clipped bilinear zoom ->
// hard threshold -> clip.  Blur: per-image tap list (integer angle -135..-45), clamp border.
struct SnowArgs {
  int n, h, w;
  float loc, scale, thresh, ih_scale;
  const uint2* ztab;                       // (h + w) x {i0 | i1 << 16, frac}
  const uint8_t* taps;                     // tap table (k1_taps format)
  int n_entries, max_taps;
  uint32_t k0, k1, first_image, stream, aux_stream;
  float* layer;                            // [n, h, w]
  float* blurred;                          // [n, h, w]
};
__device__ __forceinline__ float snow_field_at(const SnowArgs& a, uint32_t gimg, int yy, int xx) {
  const uint4 r = philox4x32_10(uint32_t(yy * a.w + xx), gimg, 0u, a.stream, a.k0, a.k1);
  const int ssum = int((r.x & 0xFFFFu) + (r.x >> 16) + (r.y & 0xFFFFu) + (r.y >> 16) + (r.z & 0xFFFFu) + (r.z >> 16) +
                       (r.w & 0xFFFFu) + (r.w >> 16)) - 4 * 65535;
  return __fadd_rn(a.loc, __fmul_rn(a.scale, __fmul_rn(float(ssum), a.ih_scale)));
}
__global__ void __launch_bounds__(256) k1_snow_layer(const SnowArgs a) {
  const long long total = (long long)a.n * a.h * a.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.h * a.w));
    const int rem = int(i - (long long)img * a.h * a.w);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint32_t gimg = a.first_image + uint32_t(img);
    const uint2 ry = __ldg(a.ztab + y), rx = __ldg(a.ztab + a.h + x);
    const int y0 = int(ry.x & 0xFFFF), y1 = int(ry.x >> 16), x0 = int(rx.x & 0xFFFF), x1 = int(rx.x >> 16);
    const float fy = __uint_as_float(ry.y), fx = __uint_as_float(rx.y);
    const float v00 = snow_field_at(a, gimg, y0, x0), v01 = snow_field_at(a, gimg, y0, x1);
    const float v10 = snow_field_at(a, gimg, y1, x0), v11 = snow_field_at(a, gimg, y1, x1);
    const float top = __fadd_rn(__fmul_rn(v00, __fsub_rn(1.0f, fx)), __fmul_rn(v01, fx));
    const float bot = __fadd_rn(__fmul_rn(v10, __fsub_rn(1.0f, fx)), __fmul_rn(v11, fx));
    float v = __fadd_rn(__fmul_rn(top, __fsub_rn(1.0f, fy)), __fmul_rn(bot, fy));
    if (v < a.thresh) v = 0.f;
    a.layer[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}
__global__ void __launch_bounds__(256) k1_snow_blur(const SnowArgs a) {
  const long long total = (long long)a.n * a.h * a.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.h * a.w));
    const int rem = int(i - (long long)img * a.h * a.w);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint4 r = philox4x32_10(0u, a.first_image + uint32_t(img), 0u, a.aux_stream, a.k0, a.k1);
    const uint8_t* ent = a.taps + (size_t)(r.x % uint32_t(a.n_entries)) * (16 + 8 * (size_t)a.max_taps);
    const int ntaps = *reinterpret_cast<const int*>(ent);
    const uint2* tp = reinterpret_cast<const uint2*>(ent + 16);
    const float* L = a.layer + (size_t)img * a.h * a.w;
    float acc = 0.f;
    for (int t = 0; t < ntaps; ++t) {
      const uint2 q = __ldg(tp + t);
      const int dy = int(short(q.x & 0xFFFF)), dx = int(short(q.x >> 16));
      acc = fmaf(__uint_as_float(q.y), L[min(max(y + dy, 0), a.h - 1) * a.w + min(max(x + dx, 0), a.w - 1)], acc);
    }
    a.blurred[i] = acc;
  }
}

// ---------------------------------------------------------------- elastic_transform
// (1) random affine warp (bilinear, reflect-101) + raw U(-1,1) displacement fields, (2) Gaussian x-pass, (3) Gaussian y-pass,
// scale by alpha, bilinear sample of the warped image with symmetric reflection, normalize.  scratch planes (fp32, n*h*w
// each): warped R,G,B | U0,U1 | P0,P1.
struct ElasticArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w, radius;
  int folded;                 // 1: taps holds the folded smoothing matrices [w][w] (transposed: [src x][dst x]) then [h][h] ([dst y][src y])
  float alpha, mag, c0, c1, sq;
  const float* taps;
  uint32_t k0, k1, first_image, stream, aux_stream;
  float* scratch;
  unsigned src_bgr;
};
__device__ __forceinline__ int reflect_sym(int i, int n) {
  i %= 2 * n;
  if (i < 0) i += 2 * n;
  return i >= n ? 2 * n - 1 - i : i;
}
__device__ __forceinline__ int reflect101_clamped(int i, int n) {
  i = min(max(i, -(n - 1)), 2 * (n - 1));
  i = abs(i);
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}
__global__ void __launch_bounds__(256) k1_elastic_prep(const ElasticArgs a) {
  const long long hw = (long long)a.h * a.w, total = a.n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / hw);
    const int rem = int(i - img * hw);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint32_t gimg = a.first_image + uint32_t(img);
    // inverse affine map dst -> src from the three jittered control points
    const uint4 ja = philox4x32_10(0u, gimg, 0u, a.aux_stream, a.k0, a.k1);
    const uint4 jb = philox4x32_10(1u, gimg, 0u, a.aux_stream, a.k0, a.k1);
    const uint32_t jw[6] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y};
    float jit[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) jit[k] = (2.0f * u32_to_uniform(jw[k]) - 1.0f) * a.mag;
    const float p[3][2] = {{a.c0 + a.sq, a.c1 + a.sq}, {a.c0 + a.sq, a.c1 - a.sq}, {a.c0 - a.sq, a.c1 - a.sq}};
    float q[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) { q[k][0] = p[k][0] + jit[2 * k]; q[k][1] = p[k][1] + jit[2 * k + 1]; }
    const float e1x = q[1][0] - q[0][0], e1y = q[1][1] - q[0][1], e2x = q[2][0] - q[0][0], e2y = q[2][1] - q[0][1];
    const float d1x = p[1][0] - p[0][0], d1y = p[1][1] - p[0][1], d2x = p[2][0] - p[0][0], d2y = p[2][1] - p[0][1];
    const float inv = 1.0f / (e1x * e2y - e2x * e1y);
    const float qi00 = e2y * inv, qi01 = -e2x * inv, qi10 = -e1y * inv, qi11 = e1x * inv;
    const float A00 = d1x * qi00 + d2x * qi10, A01 = d1x * qi01 + d2x * qi11;
    const float A10 = d1y * qi00 + d2y * qi10, A11 = d1y * qi01 + d2y * qi11;
    const float dx = float(x) - q[0][0], dy = float(y) - q[0][1];
    const float sx = p[0][0] + (A00 * dx + A01 * dy), sy = p[0][1] + (A10 * dx + A11 * dy);
    const float fx0 = floorf(sx), fy0 = floorf(sy), fx = sx - fx0, fy = sy - fy0;
    const int xa = reflect101_clamped(int(fx0), a.w), xb = reflect101_clamped(int(fx0) + 1, a.w);
    const int ya = reflect101_clamped(int(fy0), a.h), yb = reflect101_clamped(int(fy0) + 1, a.h);
    const uint8_t* im = a.src + (size_t)img * hw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = a.src_bgr ? 2 - c : c;
      const float v00 = u8f(im[((size_t)ya * a.w + xa) * 3 + cs]), v01 = u8f(im[((size_t)ya * a.w + xb) * 3 + cs]);
      const float v10 = u8f(im[((size_t)yb * a.w + xa) * 3 + cs]), v11 = u8f(im[((size_t)yb * a.w + xb) * 3 + cs]);
      const float top = v00 * (1.0f - fx) + v01 * fx, bot = v10 * (1.0f - fx) + v11 * fx;
      a.scratch[(size_t)c * total + i] = top * (1.0f - fy) + bot * fy;
    }
    const uint4 r = philox4x32_10(uint32_t(rem), gimg, 0u, a.stream, a.k0, a.k1);
    a.scratch[3 * (size_t)total + i] = 2.0f * u32_to_uniform(r.x) - 1.0f;
    a.scratch[4 * (size_t)total + i] = 2.0f * u32_to_uniform(r.y) - 1.0f;
  }
}
__global__ void __launch_bounds__(256) k1_elastic_xpass(const ElasticArgs a) {
  const long long hw = (long long)a.h * a.w, total = a.n * hw;
  if (a.folded && (a.w & 3) == 0) {
    // folded matrices, four destination pixels per thread: one 16-byte load of weights and two broadcast loads of the source
    // row feed eight FMAs (same per-output operation order as the scalar loop below)
    const int w4 = a.w >> 2;
    for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < (total >> 2); i4 += (long long)gridDim.x * blockDim.x) {
      const long long rowi = i4 / w4;                                  // image row index (img * h + y)
      const int x = int(i4 - rowi * w4) << 2;
      const float* u0 = a.scratch + 3 * (size_t)total + rowi * a.w;
      const float* u1 = a.scratch + 4 * (size_t)total + rowi * a.w;
      const float4* mt = reinterpret_cast<const float4*>(a.taps + x);
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int xx = 0; xx < a.w; ++xx) {
        const float4 k = __ldg(mt + (size_t)xx * w4);
        const float b0 = u0[xx], b1 = u1[xx];
        s0[0] = fmaf(k.x, b0, s0[0]); s0[1] = fmaf(k.y, b0, s0[1]); s0[2] = fmaf(k.z, b0, s0[2]); s0[3] = fmaf(k.w, b0, s0[3]);
        s1[0] = fmaf(k.x, b1, s1[0]); s1[1] = fmaf(k.y, b1, s1[1]); s1[2] = fmaf(k.z, b1, s1[2]); s1[3] = fmaf(k.w, b1, s1[3]);
      }
      *reinterpret_cast<float4*>(a.scratch + 5 * (size_t)total + rowi * a.w + x) = make_float4(s0[0], s0[1], s0[2], s0[3]);
      *reinterpret_cast<float4*>(a.scratch + 6 * (size_t)total + rowi * a.w + x) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int rem = int(i % hw), x = rem % a.w;
    const float* u0 = a.scratch + 3 * (size_t)total + (i - x);
    const float* u1 = a.scratch + 4 * (size_t)total + (i - x);
    float s0 = 0.f, s1 = 0.f;
    if (a.folded) {
      // a Gaussian much longer than the row wraps around the reflected row many times (224-pixel rows, 1025 taps at severity 1):
      // the host folds it into one weight per (destination, source) pixel, so the pass is a w x w mat-vec without index arithmetic
      const float* mt = a.taps + x;
      for (int xx = 0; xx < a.w; ++xx) {
        const float k = __ldg(mt + (size_t)xx * a.w);
        s0 = fmaf(k, u0[xx], s0); s1 = fmaf(k, u1[xx], s1);
      }
    } else {
      int pos = (x - a.radius) % (2 * a.w);                       // position in the reflected period, advanced without a modulo per tap
      if (pos < 0) pos += 2 * a.w;
      for (int t = 0; t <= 2 * a.radius; ++t) {
        const int xx = pos < a.w ? pos : 2 * a.w - 1 - pos;
        const float k = __ldg(a.taps + t);
        s0 = fmaf(k, u0[xx], s0); s1 = fmaf(k, u1[xx], s1);
        if (++pos == 2 * a.w) pos = 0;
      }
    }
    a.scratch[5 * (size_t)total + i] = s0;
    a.scratch[6 * (size_t)total + i] = s1;
  }
}
__global__ void __launch_bounds__(256) k1_elastic_final(const ElasticArgs a) {
  const long long hw = (long long)a.h * a.w, total = a.n * hw;
  // displaced bilinear sample of the warped image at pixel (y, x) of image img with displacement (d0, d1) * alpha
  auto sample_store = [&](int img, int y, int x, float d0, float d1) {
    const float sx = float(x) + d0 * a.alpha, sy = float(y) + d1 * a.alpha;
    const float fx0 = floorf(sx), fy0 = floorf(sy), fx = sx - fx0, fy = sy - fy0;
    const int xa = reflect_sym(int(fx0), a.w), xb = reflect_sym(int(fx0) + 1, a.w);
    const int ya = reflect_sym(int(fy0), a.h), yb = reflect_sym(int(fy0) + 1, a.h);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* im = a.scratch + (size_t)c * total + (size_t)img * hw;
      const float top = im[ya * a.w + xa] * (1.0f - fx) + im[ya * a.w + xb] * fx;
      const float bot = im[yb * a.w + xa] * (1.0f - fx) + im[yb * a.w + xb] * fx;
      o[c] = top * (1.0f - fy) + bot * fy;
    }
    store_pixel(a.out, (size_t)img * hw + (size_t)y * a.w + x, o[0], o[1], o[2]);
  };
  if (a.folded && (a.w & 3) == 0) {
    // folded y-pass, four neighbouring columns per thread: one broadcast weight and two 16-byte loads feed eight FMAs
    const int w4 = a.w >> 2;
    for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < (total >> 2); i4 += (long long)gridDim.x * blockDim.x) {
      const long long rowi = i4 / w4;
      const int x = int(i4 - rowi * w4) << 2;
      const int img = int(rowi / a.h), y = int(rowi - (long long)img * a.h);
      const float4* p0 = reinterpret_cast<const float4*>(a.scratch + 5 * (size_t)total + (size_t)img * hw + x);
      const float4* p1 = reinterpret_cast<const float4*>(a.scratch + 6 * (size_t)total + (size_t)img * hw + x);
      const float* mh = a.taps + (size_t)a.w * a.w + (size_t)y * a.h;
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int yy = 0; yy < a.h; ++yy) {
        const float k = __ldg(mh + yy);
        const float4 v0 = p0[(size_t)yy * w4], v1 = p1[(size_t)yy * w4];
        d0[0] = fmaf(k, v0.x, d0[0]); d0[1] = fmaf(k, v0.y, d0[1]); d0[2] = fmaf(k, v0.z, d0[2]); d0[3] = fmaf(k, v0.w, d0[3]);
        d1[0] = fmaf(k, v1.x, d1[0]); d1[1] = fmaf(k, v1.y, d1[1]); d1[2] = fmaf(k, v1.z, d1[2]); d1[3] = fmaf(k, v1.w, d1[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sample_store(img, y, x + j, d0[j], d1[j]);
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / hw);
    const int rem = int(i - img * hw);
    const int y = rem / a.w, x = rem - y * a.w;
    const float* p0 = a.scratch + 5 * (size_t)total + (size_t)img * hw + x;
    const float* p1 = a.scratch + 6 * (size_t)total + (size_t)img * hw + x;
    float d0 = 0.f, d1 = 0.f;
    if (a.folded) {
      const float* mh = a.taps + (size_t)a.w * a.w + (size_t)y * a.h;
      for (int yy = 0; yy < a.h; ++yy) {
        const float k = __ldg(mh + yy);
        d0 = fmaf(k, p0[(size_t)yy * a.w], d0); d1 = fmaf(k, p1[(size_t)yy * a.w], d1);
      }
    } else {
      int pos = (y - a.radius) % (2 * a.h);
      if (pos < 0) pos += 2 * a.h;
      for (int t = 0; t <= 2 * a.radius; ++t) {
        const int yy = pos < a.h ? pos : 2 * a.h - 1 - pos;
        const float k = __ldg(a.taps + t);
        d0 = fmaf(k, p0[(size_t)yy * a.w], d0); d1 = fmaf(k, p1[(size_t)yy * a.w], d1);
        if (++pos == 2 * a.h) pos = 0;
      }
    }
    sample_store(img, y, x, d0, d1);
  }
}

// ---------------------------------------------------------------- jpeg_compression: libjpeg baseline round trip
// What PIL's save(quality=c) + reload computes (entropy coding is lossless and omitted), stage by stage after libjpeg:
// jccolor.c rgb_ycc_convert -> jcsample.c h2v2_downsample (bias 1,2,1,2) -> jfdctint.c jpeg_fdct_islow -> jcdctmgr.c
// quantize (divisor 8 q) -> jidctint.c jpeg_idct_islow + range limit -> jdsample.c h2v2_fancy_upsample -> jdcolor.c
// ycc_rgb_convert.  All integer: BYTE-EXACT against Pillow (tests) and against oracle/jpeg.py.
// Two kernels: (1) one CTA of 64 threads per 16x16 MCU: colour transform, chroma subsampling, six 8x8 blocks through
// FDCT / quantise / IDCT, reconstructed Y / Cb / Cr planes to scratch; (2) per pixel: fancy chroma upsampling (needs the
// neighbouring MCUs' chroma), colour transform, normalize.  table: int32 lum[64], chr[64] (natural order).
struct JpegArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w, mcu_x, mcu_y;
  const int* table;
  uint8_t* planes;               // scratch per image: Y [H16][W16] | Cb [H16/2][W16/2] | Cr [H16/2][W16/2]
  unsigned src_bgr;
};
#define JDESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
// jfdctint.c, one 1-D pass (CONST_BITS 13, PASS1_BITS 2); first: row pass (results scaled up by 4)
__device__ __forceinline__ void jpeg_fdct_1d(int* d, bool first) {
  const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
  const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  const int sh = first ? 11 : 15;
  d[0] = first ? (tmp10 + tmp11) << 2 : JDESCALE(tmp10 + tmp11, 2);
  d[4] = first ? (tmp10 - tmp11) << 2 : JDESCALE(tmp10 - tmp11, 2);
  int z1 = (tmp12 + tmp13) * 4433;
  d[2] = (z1 + tmp13 * 6270 + (1 << (sh - 1))) >> sh;
  d[6] = (z1 - tmp12 * 15137 + (1 << (sh - 1))) >> sh;
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * 9633;
  const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * -16069 + z5; z4 = z4 * -3196 + z5;
  d[7] = (t4 + z1 + z3 + (1 << (sh - 1))) >> sh;
  d[5] = (t5 + z2 + z4 + (1 << (sh - 1))) >> sh;
  d[3] = (t6 + z2 + z3 + (1 << (sh - 1))) >> sh;
  d[1] = (t7 + z1 + z4 + (1 << (sh - 1))) >> sh;
}
// jidctint.c, one 1-D pass; first: column pass into the workspace (descale 11), else row pass (descale 18)
__device__ __forceinline__ void jpeg_idct_1d(int* c, bool first) {
  int z2 = c[2], z3 = c[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 - z3 * 15137, tmp3 = z1 + z2 * 6270;
  int tmp0 = (c[0] + c[4]) << 13, tmp1 = (c[0] - c[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = c[7]; tmp1 = c[5]; tmp2 = c[3]; tmp3 = c[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * -16069 + z5; z4 = z4 * -3196 + z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int sh = first ? 11 : 18;
  c[0] = (tmp10 + tmp3 + (1 << (sh - 1))) >> sh; c[7] = (tmp10 - tmp3 + (1 << (sh - 1))) >> sh;
  c[1] = (tmp11 + tmp2 + (1 << (sh - 1))) >> sh; c[6] = (tmp11 - tmp2 + (1 << (sh - 1))) >> sh;
  c[2] = (tmp12 + tmp1 + (1 << (sh - 1))) >> sh; c[5] = (tmp12 - tmp1 + (1 << (sh - 1))) >> sh;
  c[3] = (tmp13 + tmp0 + (1 << (sh - 1))) >> sh; c[4] = (tmp13 - tmp0 + (1 << (sh - 1))) >> sh;
}
// jdmaster.c prepare_range_limit_table as the IDCT indexes it: (v & 1023) into a table centred on 128
__device__ __forceinline__ uint32_t jpeg_range_limit(int v) {
  v &= 1023;
  return uint32_t(v < 128 ? v + 128 : (v < 512 ? 255 : (v < 896 ? 0 : v - 896)));
}
__device__ __forceinline__ void jpeg_ycc(const uint8_t* q, bool bgr, int& Y, int& Cb, int& Cr) {
  int R = q[0], G = q[1], B = q[2];
  if (bgr) { const int t = R; R = B; B = t; }
  Y = (19595 * R + 38470 * G + 7471 * B + 32768) >> 16;
  Cb = (-11059 * R - 21709 * G + 32768 * B + 8388608 + 32767) >> 16;
  Cr = (32768 * R - 27439 * G - 5329 * B + 8388608 + 32767) >> 16;
}
__global__ void __launch_bounds__(64) k1_jpeg_codec(const JpegArgs a) {
  __shared__ int blk[6][8][9];      // Y00, Y01, Y10, Y11, Cb, Cr; rows padded to 9 words
  __shared__ int sQ[128];
  const int img = blockIdx.x, my = blockIdx.y / a.mcu_x, mx = blockIdx.y - my * a.mcu_x;
  const int t = threadIdx.x;
  sQ[t] = a.table[t]; sQ[t + 64] = a.table[t + 64];
  const uint8_t* p = a.src + (size_t)img * a.h * a.w * 3;
  const bool bgr = a.src_bgr != 0;
  // luma: right / bottom edges replicate (jcsample.c expand_right_edge, jcprepct.c expand_bottom_edge)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = t + 64 * j, ty = i >> 4, tx = i & 15;
    const int y = min(my * 16 + ty, a.h - 1), x = min(mx * 16 + tx, a.w - 1);
    int Y, Cb, Cr;
    jpeg_ycc(p + ((size_t)y * a.w + x) * 3, bgr, Y, Cb, Cr);
    blk[(ty >> 3) * 2 + (tx >> 3)][ty & 7][tx & 7] = Y - 128;
  }
  // chroma: an odd last row is doubled before the 2x2 mean, then the DOWNSAMPLED rows replicate (jcprepct.c pre_process_data)
  {
    const int cy = t >> 3, cx = t & 7, ch = (a.h + 1) >> 1;
    const int gr = min(my * 8 + cy, ch - 1), gc = mx * 8 + cx;
    const int y0 = 2 * gr, y1 = min(2 * gr + 1, a.h - 1), x0 = min(2 * gc, a.w - 1), x1 = min(2 * gc + 1, a.w - 1);
    int sb = 0, sr = 0, Y, Cb, Cr;
    jpeg_ycc(p + ((size_t)y0 * a.w + x0) * 3, bgr, Y, Cb, Cr); sb += Cb; sr += Cr;
    jpeg_ycc(p + ((size_t)y0 * a.w + x1) * 3, bgr, Y, Cb, Cr); sb += Cb; sr += Cr;
    jpeg_ycc(p + ((size_t)y1 * a.w + x0) * 3, bgr, Y, Cb, Cr); sb += Cb; sr += Cr;
    jpeg_ycc(p + ((size_t)y1 * a.w + x1) * 3, bgr, Y, Cb, Cr); sb += Cb; sr += Cr;
    const int bias = 1 + (cx & 1);                       // h2v2_downsample: 1, 2, 1, 2, ... along the output row
    blk[4][cy][cx] = ((sb + bias) >> 2) - 128;
    blk[5][cy][cx] = ((sr + bias) >> 2) - 128;
  }
  __syncthreads();
  const int b = t >> 3, r = t & 7;
  int d[8];
  if (t < 48) {                                          // FDCT pass 1: rows
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = blk[b][r][k];
    jpeg_fdct_1d(d, true);
#pragma unroll
    for (int k = 0; k < 8; ++k) blk[b][r][k] = d[k];
  }
  __syncthreads();
  if (t < 48) {                                          // FDCT pass 2 (columns) -> quantise / dequantise -> IDCT pass 1 (columns)
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = blk[b][k][r];
    jpeg_fdct_1d(d, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int Q = sQ[(b < 4 ? 0 : 64) + k * 8 + r], dv = Q * 8;
      const int qa = (abs(d[k]) + (dv >> 1)) / dv;       // jcdctmgr.c quantize: round half away from zero
      d[k] = (d[k] < 0 ? -qa : qa) * Q;
    }
    jpeg_idct_1d(d, true);
#pragma unroll
    for (int k = 0; k < 8; ++k) blk[b][k][r] = d[k];
  }
  __syncthreads();
  if (t < 48) {                                          // IDCT pass 2: rows -> range limit -> 8 bytes to the planes
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = blk[b][r][k];
    jpeg_idct_1d(d, false);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo |= jpeg_range_limit(d[k]) << (8 * k); hi |= jpeg_range_limit(d[k + 4]) << (8 * k); }
    const int W16 = a.mcu_x * 16, H16 = a.mcu_y * 16;
    uint8_t* base = a.planes + (size_t)img * ((size_t)H16 * W16 * 3 / 2);
    uint8_t* o;
    if (b < 4) o = base + (size_t)(my * 16 + (b >> 1) * 8 + r) * W16 + mx * 16 + (b & 1) * 8;
    else o = base + (size_t)H16 * W16 + (size_t)(b - 4) * (H16 / 2) * (W16 / 2) + (size_t)(my * 8 + r) * (W16 / 2) + mx * 8;
    *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
  }
}
__global__ void __launch_bounds__(256) k1_jpeg_finish(const JpegArgs a) {
  const long long total = (long long)a.n * a.h * a.w;
  const int W16 = a.mcu_x * 16, H16 = a.mcu_y * 16, CW = W16 / 2;
  const int ch = (a.h + 1) >> 1, cw = (a.w + 1) >> 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / (a.h * a.w));
    const int rem = int(i - (long long)img * a.h * a.w);
    const int y = rem / a.w, x = rem - y * a.w;
    const uint8_t* pY = a.planes + (size_t)img * ((size_t)H16 * W16 * 3 / 2);
    const uint8_t* pC = pY + (size_t)H16 * W16;
    const int Y = pY[(size_t)y * W16 + x];
    // h2v2_fancy_upsample: the nearer chroma row / column weighs 3, the farther 1; context beyond the image replicates
    const int r = y >> 1, far = (y & 1) ? min(r + 1, ch - 1) : max(r - 1, 0);
    const int c = x >> 1, nb = (x & 1) ? min(c + 1, cw - 1) : max(c - 1, 0);
    const int bias = (x & 1) ? 7 : 8;
    int cc[2];
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      const uint8_t* P = pC + (size_t)pl * (H16 / 2) * CW;
      const int cs = 3 * int(P[r * CW + c]) + int(P[far * CW + c]);
      const int cn = 3 * int(P[r * CW + nb]) + int(P[far * CW + nb]);
      cc[pl] = ((3 * cs + cn + bias) >> 4) - 128;
    }
    const int R = min(max(Y + ((91881 * cc[1] + 32768) >> 16), 0), 255);
    const int G = min(max(Y + ((-22554 * cc[0] - 46802 * cc[1] + 32768) >> 16), 0), 255);
    const int B = min(max(Y + ((116130 * cc[0] + 32768) >> 16), 0), 255);
    store_pixel(a.out, (size_t)i, div255(float(R)), div255(float(G)), div255(float(B)));
  }
}

// ---------------------------------------------------------------- glass_blur: blur -> sequential local swaps -> blur
// One CTA per image, the whole image resident in shared memory as packed RGBX words.  Stage 1: exact fixed-point Gaussian
// (bit-exact bytes).  Stage 2: the scan-order swap chain.  A swap at (hh, ww) touches two pixels within rows hh-d .. hh+d-1
// and columns ww-d .. ww+d-1, so swaps of different scan rows commute unless their columns are closer than 2d: the chain
// runs as a WAVEFRONT -- lane r of warp 0 walks scan row r, 2d columns behind lane r-1 (every pair of swaps that shares
// a pixel keeps its sequential order, so the result is bit-identical to the scalar loop of oracle/corruptions.py), while
// all threads pre-compute the Philox offsets of a band of rows.  Stage 3: fp32 Gaussian + normalize.
// table: int32 q16[2r+1], float k[2r+1]
struct GlassArgs {
  const uint8_t* src;
  OutArgs out;
  int n, h, w, delta, iters, radius;
  const int* table;
  uint32_t k0, k1, first_image, stream;
  unsigned src_bgr;
};
constexpr int GLASS_CHUNK = 16384;      // swap offsets (one byte each) computed ahead, in whole scan rows
constexpr int GLASS_BAND = 16;          // output rows per thread in the separable blurs
template <bool BANDED>   // BANDED: separable band-scheme blurs (large images); else one thread per pixel, direct 2-D sums (few registers:
                         // small images run many CTAs per SM)
__global__ void __launch_bounds__(BANDED ? 512 : 256) k1_glass(const GlassArgs a) {
  extern __shared__ uint32_t g_img[];                       // h*w RGBX words, then GLASS_CHUNK packed offsets, then taps
  const int hw = a.h * a.w, nt = 2 * a.radius + 1;
  uint8_t* s_off = reinterpret_cast<uint8_t*>(g_img + hw);
  int* s_q = reinterpret_cast<int*>(s_off + GLASS_CHUNK);
  float* s_k = reinterpret_cast<float*>(s_q + nt);
  const int img = blockIdx.x;
  const uint8_t* p = a.src + (size_t)img * hw * 3;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) { s_q[i] = a.table[i]; s_k[i] = reinterpret_cast<const float*>(a.table)[nt + i]; }
  __syncthreads();
  // stage 1: exact integer blur, (sum_y q[y] * ((sum_x q[x] * u8 + 128) >> 8) + 2^23) >> 24.  Separable without a full
  // intermediate image: a thread owns column x of a band of GLASS_BAND output rows, computes the x-pass of each source row the
  // band needs ONCE and adds it into the (<= 2r+1) output rows it contributes to (accumulators in registers).
  const int bands = (a.h + GLASS_BAND - 1) / GLASS_BAND;
  if (BANDED) {
    for (int item = threadIdx.x; item < bands * a.w; item += blockDim.x) {
      const int bnd = item / a.w, x = item - bnd * a.w, y0 = bnd * GLASS_BAND;
      unsigned acc[GLASS_BAND][3];
#pragma unroll
      for (int j = 0; j < GLASS_BAND; ++j) { acc[j][0] = 0; acc[j][1] = 0; acc[j][2] = 0; }
      for (int yy = y0 - a.radius; yy < y0 + GLASS_BAND + a.radius; ++yy) {
        const uint8_t* row = p + (size_t)min(max(yy, 0), a.h - 1) * a.w * 3;
        unsigned h3[3] = {0, 0, 0};
        for (int dx = -a.radius; dx <= a.radius; ++dx) {
          const uint8_t* q = row + min(max(x + dx, 0), a.w - 1) * 3;
          const unsigned wq = unsigned(s_q[dx + a.radius]);
          h3[0] += wq * q[0]; h3[1] += wq * q[1]; h3[2] += wq * q[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) h3[c] = (h3[c] + 128u) >> 8;
#pragma unroll
        for (int j = 0; j < GLASS_BAND; ++j) {
          const int t = yy - (y0 + j) + a.radius;
          if (t >= 0 && t <= 2 * a.radius) {
            const unsigned wy = unsigned(s_q[t]);
            acc[j][0] += wy * h3[0]; acc[j][1] += wy * h3[1]; acc[j][2] += wy * h3[2];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < GLASS_BAND; ++j) {
        if (y0 + j >= a.h) continue;
        unsigned bb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) bb[c] = min((acc[j][c] + (1u << 23)) >> 24, 255u);
        if (a.src_bgr) { const unsigned q = bb[0]; bb[0] = bb[2]; bb[2] = q; }
        g_img[(y0 + j) * a.w + x] = bb[0] | (bb[1] << 8) | (bb[2] << 16);
      }
    }
  } else {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int y = i / a.w, x = i - y * a.w;
      unsigned acc[3] = {0, 0, 0};
      for (int dy = -a.radius; dy <= a.radius; ++dy) {
        const uint8_t* row = p + (size_t)min(max(y + dy, 0), a.h - 1) * a.w * 3;
        unsigned h3[3] = {0, 0, 0};
        for (int dx = -a.radius; dx <= a.radius; ++dx) {
          const uint8_t* q = row + min(max(x + dx, 0), a.w - 1) * 3;
          const unsigned wq = unsigned(s_q[dx + a.radius]);
          h3[0] += wq * q[0]; h3[1] += wq * q[1]; h3[2] += wq * q[2];
        }
        const unsigned wy = unsigned(s_q[dy + a.radius]);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += wy * ((h3[c] + 128u) >> 8);
      }
      unsigned bb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) bb[c] = min((acc[c] + (1u << 23)) >> 24, 255u);
      if (a.src_bgr) { const unsigned q = bb[0]; bb[0] = bb[2]; bb[2] = q; }
      g_img[i] = bb[0] | (bb[1] << 8) | (bb[2] << 16);
    }
  }
  __syncthreads();
  // stage 2: swaps.  step j = R * sw + k visits scan row R (iteration R / sh, hh = h - delta - R % sh) at ww = w - delta - k;
  // offsets d in [-delta, delta-1] from Philox(j)
  const int sh = a.h - 2 * a.delta, sw = a.w - 2 * a.delta;
  const int rows_total = a.iters * sh;
  const uint32_t gimg = a.first_image + uint32_t(img), m = uint32_t(2 * a.delta);
  int band = GLASS_CHUNK / sw;                                  // scan rows whose offsets fit the buffer
  if (band > rows_total) band = rows_total;
  // lag between consecutive rows: 2 delta columns keep conflicting swaps in order; a lane must also be done with its row
  // before its next one (32 rows later) starts
  const int lag = max(2 * a.delta, (sw + 31) / 32);
  const int lane = threadIdx.x & 31;
  for (int R0 = 0; R0 < rows_total; R0 += band) {
    const int nb = min(band, rows_total - R0);
    for (int i = threadIdx.x; i < nb * sw; i += blockDim.x) {
      const uint4 r = philox4x32_10(uint32_t(R0 * sw + i), gimg, 0u, a.stream, a.k0, a.k1);
      s_off[i] = uint8_t((r.y % m) | ((r.x % m) << 4));           // (dy + delta) | (dx + delta) << 4
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      int r_cur = lane, k = -lane * lag;                        // this lane's current band row and its column index at step s
      int hh = a.h - a.delta - (R0 + r_cur) % sh;
      const int steps = (nb - 1) * lag + sw;
      for (int s = 0; s < steps; ++s) {
        if (k >= sw) {                                          // row finished: on to the row 32 further (32 * lag >= sw: k <= 0 again)
          r_cur += 32; k -= 32 * lag;
          hh = a.h - a.delta - (R0 + r_cur) % sh;
        }
        if (k >= 0 && k < sw && r_cur < nb) {
          const int ww = a.w - a.delta - k;
          const uint32_t o = s_off[r_cur * sw + k];
          const int i0 = hh * a.w + ww, i1 = (hh + int(o & 15u) - a.delta) * a.w + (ww + int(o >> 4) - a.delta);
          const uint32_t t0 = g_img[i0];
          g_img[i0] = g_img[i1];
          g_img[i1] = t0;
        }
        ++k;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  // stage 3: fp32 Gaussian (x then y, accumulated in tap order like the oracle), same band scheme: the contributions to an
  // output row arrive in ascending source-row order, i.e. exactly the tap order of the direct double loop
  if (BANDED) {
    for (int item = threadIdx.x; item < bands * a.w; item += blockDim.x) {
      const int bnd = item / a.w, x = item - bnd * a.w, y0 = bnd * GLASS_BAND;
      float acc[GLASS_BAND][3];
#pragma unroll
      for (int j = 0; j < GLASS_BAND; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; }
      for (int yy = y0 - a.radius; yy < y0 + GLASS_BAND + a.radius; ++yy) {
        const uint32_t* row = g_img + min(max(yy, 0), a.h - 1) * a.w;
        float h3[3] = {0.f, 0.f, 0.f};
        for (int dx = -a.radius; dx <= a.radius; ++dx) {
          const uint32_t v = row[min(max(x + dx, 0), a.w - 1)];
          const float wk = s_k[dx + a.radius];
          h3[0] = fmaf(wk, div255(float(v & 0xFF)), h3[0]);
          h3[1] = fmaf(wk, div255(float((v >> 8) & 0xFF)), h3[1]);
          h3[2] = fmaf(wk, div255(float((v >> 16) & 0xFF)), h3[2]);
        }
#pragma unroll
        for (int j = 0; j < GLASS_BAND; ++j) {
          const int t = yy - (y0 + j) + a.radius;
          if (t >= 0 && t <= 2 * a.radius) {
            const float wy = s_k[t];
            acc[j][0] = fmaf(wy, h3[0], acc[j][0]); acc[j][1] = fmaf(wy, h3[1], acc[j][1]); acc[j][2] = fmaf(wy, h3[2], acc[j][2]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < GLASS_BAND; ++j)
        if (y0 + j < a.h) store_pixel(a.out, (size_t)img * hw + (size_t)(y0 + j) * a.w + x, acc[j][0], acc[j][1], acc[j][2]);
    }
  } else {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int y = i / a.w, x = i - y * a.w;
      float acc[3] = {0.f, 0.f, 0.f};
      for (int dy = -a.radius; dy <= a.radius; ++dy) {
        const uint32_t* row = g_img + min(max(y + dy, 0), a.h - 1) * a.w;
        float h3[3] = {0.f, 0.f, 0.f};
        for (int dx = -a.radius; dx <= a.radius; ++dx) {
          const uint32_t v = row[min(max(x + dx, 0), a.w - 1)];
          const float wk = s_k[dx + a.radius];
          h3[0] = fmaf(wk, div255(float(v & 0xFF)), h3[0]);
          h3[1] = fmaf(wk, div255(float((v >> 8) & 0xFF)), h3[1]);
          h3[2] = fmaf(wk, div255(float((v >> 16) & 0xFF)), h3[2]);
        }
        const float wy = s_k[dy + a.radius];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = fmaf(wy, h3[c], acc[c]);
      }
      store_pixel(a.out, (size_t)img * hw + i, acc[0], acc[1], acc[2]);
    }
  }
}

// ---------------------------------------------------------------- synthetic inputs
__global__ void k_synth_images(uint8_t* dst, int n, int per, uint32_t k0, uint32_t k1, uint32_t first_image) {
  const int nch = (per + 15) / 16;
  const long long total = (long long)n * nch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = int(i / nch), ch = int(i - (long long)img * nch);
    const uint4 r = philox4x32_10(uint32_t(ch), first_image + uint32_t(img), 0u, stream_id(KIND_IMAGES), k0, k1);
    uint8_t* o = dst + (size_t)img * per + (size_t)ch * 16;
    const int cnt = min(16, per - ch * 16);
    if (cnt == 16 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      *reinterpret_cast<uint4*>(o) = r;
    } else {
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
      for (int b = 0; b < cnt; ++b) o[b] = uint8_t(w[b >> 2] >> (8 * (b & 3)));
    }
  }
}
__global__ void k_synth_labels(int32_t* dst, int n, int C, uint32_t k0, uint32_t k1, uint32_t first_image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint4 r = philox4x32_10(0u, first_image + uint32_t(i), 0u, stream_id(KIND_LABELS), k0, k1);
    dst[i] = int32_t(r.x % uint32_t(C));
  }
}

static inline int grid_for(long long work, int threads, int num_sms, int waves = 8) {
  long long b = (work + threads - 1) / threads;
  long long cap = (long long)num_sms * waves;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace fav

using namespace fav;

extern "C" size_t fav_corrupt_scratch_bytes(int corruption, int n, int height, int width) {
  if (corruption == FAV_CONTRAST) return (size_t)n * 3 * sizeof(unsigned long long);
  if (corruption == FAV_SNOW) return 2 * (size_t)n * height * width * sizeof(float);
  if (corruption == FAV_ELASTIC) return 7 * (size_t)n * height * width * sizeof(float);
  if (corruption == FAV_JPEG)      // reconstructed Y + 4:2:0 chroma planes, padded to whole 16x16 MCUs
    return (size_t)n * ((size_t)((height + 15) / 16 * 16) * ((width + 15) / 16 * 16) * 3 / 2);
  if (corruption == FAV_PIXELATE) return (size_t)n * height * width * 3;      // the down-sampled image (upper bound)
  if (corruption == FAV_FOG || corruption == FAV_FROST) {
    int m = 1;
    while (m < (height > width ? height : width)) m *= 2;
    const size_t stats = (((size_t)n * 16) + 255) / 256 * 256;
    return stats + (size_t)n * m * m * sizeof(float);
  }
  return 0;
}

extern "C" int fav_corrupt_normalize_ex(fav_handle h, const uint8_t* d_src, void* d_dst, int n, int height,
                                        int width, int corruption, int severity, const float* fparams,
                                        int n_fparams, const int32_t* iparams, int n_iparams,
                                        const void* d_table, size_t table_bytes, void* d_scratch,
                                        size_t scratch_bytes, uint64_t seed, uint64_t first_image,
                                        const float mean[3], const float std[3], unsigned flags, void* stream) {
  FAV_REQUIRE(h, "fav_corrupt_normalize: null handle");
  FAV_DEVICE(h);
  flags &= (FAV_SRC_BGR | FAV_OUT_F32 | FAV_NO_NORMALIZE);            // the profile bits only select host tables
  FAV_REQUIRE(n >= 0 && height > 0 && width > 0, "fav_corrupt_normalize: bad shape n=%d h=%d w=%d", n, height, width);
  FAV_REQUIRE(n == 0 || (d_src && d_dst), "fav_corrupt_normalize: null image pointer");
  FAV_REQUIRE(corruption == FAV_CLEAN || (severity >= 1 && severity <= 5), "severity must be 1..5 (got %d)", severity);
  FAV_REQUIRE(mean && std, "fav_corrupt_normalize: mean/std required");
  if (n == 0) return FAV_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int per = height * width * 3;
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  auto need_f = [&](int k) { return fparams && n_fparams >= k; };
  auto need_i = [&](int k) { return iparams && n_iparams >= k; };

  OutArgs out;
  out.dst = d_dst;
  out.flags = flags;
  for (int c = 0; c < 3; ++c) { out.mean[c] = mean[c]; out.inv_std[c] = 1.0f / std[c]; }

  PointwiseArgs a{};
  a.src = d_src; a.dst = d_dst; a.n = n; a.per = per; a.groups_per_image = (per + 47) / 48;
  a.gpi_magic = ((unsigned long long)n * a.groups_per_image + 1) * (unsigned long long)a.groups_per_image < (1ull << 32)
                    ? uint32_t((1ull << 32) / (unsigned long long)a.groups_per_image) + 1u : 0u;
  {
    const unsigned long long cpi = (unsigned long long)(per >> 4);
    a.cpi_m64 = (cpi > 1 && (unsigned long long)n * cpi < (1ull << 32)) ? ~0ull / cpi + 1ull : 0ull;       // ceil(2^64 / cpi) for cpi > 1
  }
  a.k0 = k0; a.k1 = k1; a.first_image = uint32_t(first_image);
  a.stream = stream_id(KIND_CORRUPT, corruption, severity);
  a.table = d_table; a.scratch = d_scratch; a.hw = height * width; a.width = width; a.flags = flags;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.inv_std[c] = 1.0f / std[c]; }
  const long long groups = (long long)n * a.groups_per_image;
  const int grid = grid_for(groups, 256, h->num_sms, 16);
  // one-chunk-per-thread kernels of the element-wise modes: whole 16-byte chunks, aligned, RGB source
  const bool chunk_ok = !h->k1_legacy && (per & 15) == 0 && !(flags & FAV_SRC_BGR) && (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(d_dst) & 15) == 0;
  const int grid16 = grid_for((long long)n * (per >> 4), 256, h->num_sms, 32);

  switch (corruption) {
    case FAV_CLEAN:
      if (chunk_ok) k1_chunk16<PW_CLEAN><<<grid16, 256, 0, st>>>(a);
      else k1_pointwise<PW_CLEAN><<<grid, 256, 0, st>>>(a);
      h->launches++; break;
    case FAV_GAUSSIAN_NOISE:
      FAV_REQUIRE(need_f(1), "gaussian_noise needs fparams[0]=sigma");
      a.f0 = fparams[0];
      if (chunk_ok) k1_chunk16<PW_GAUSS><<<grid16, 256, 0, st>>>(a);
      else k1_pointwise<PW_GAUSS><<<grid, 256, 0, st>>>(a);
      h->launches++; break;
    case FAV_SHOT_NOISE: {
      FAV_REQUIRE(need_f(1) && need_i(1) && d_table, "shot_noise needs fparams[0]=c, iparams[0]=width, table");
      FAV_REQUIRE(table_bytes >= 1024 + (size_t)iparams[0] * 1024 + 256 * 256 * 2, "shot_noise table too small");
      a.f0 = fparams[0]; a.u0 = uint32_t(iparams[0]);
      const size_t guide_off = n_iparams >= 2 ? size_t(iparams[1]) : 0;       // 0: no guide table (k does not fit 10 bits)
      if (guide_off && !h->k1_legacy && table_bytes >= guide_off + 131072 && (per % 48) == 0 && !(flags & FAV_SRC_BGR) &&
          (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 15) == 0 && (guide_off & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(d_table) & 15) == 0) {
        ShotArgs sa{};
        sa.src = d_src; sa.dst = d_dst; sa.n = n; sa.per = per; sa.groups_per_image = per / 48;
        sa.k0 = k0; sa.k1 = k1; sa.first_image = uint32_t(first_image); sa.stream = a.stream;
        sa.c = fparams[0]; sa.rc = 1.0f / fparams[0]; sa.width = iparams[0];
        sa.kmin = reinterpret_cast<const int*>(d_table);
        sa.thr = reinterpret_cast<const uint32_t*>(d_table) + 256;
        sa.guide = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(d_table) + guide_off);
        for (int c = 0; c < 3; ++c) { sa.mean[c] = mean[c]; sa.inv_std[c] = 1.0f / std[c]; }
        sa.flags = flags;
        if (!h->attr_shot) {
          FAV_CUDA_OK(cudaFuncSetAttribute(k1_shot_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024));
          h->attr_shot = true;
        }
        const long long ctas = (groups + 1023) / 1024;
        k1_shot_smem<<<int(ctas < h->num_sms ? ctas : h->num_sms), 1024, 131072 + 1024, st>>>(sa);
        h->launches++; break;
      }
      k1_pointwise<PW_SHOT><<<grid, 256, 0, st>>>(a); h->launches++; break;
    }
    case FAV_IMPULSE_NOISE:
      FAV_REQUIRE(need_i(2), "impulse_noise needs iparams[0..1]=pepper,salt thresholds");
      a.u0 = uint32_t(iparams[0]); a.u1 = uint32_t(iparams[1]);
      if (chunk_ok) k1_chunk16<PW_IMPULSE><<<grid16, 256, 0, st>>>(a);
      else k1_pointwise<PW_IMPULSE><<<grid, 256, 0, st>>>(a);
      h->launches++; break;
    case FAV_BRIGHTNESS:
      FAV_REQUIRE(need_f(1), "brightness needs fparams[0]=c");
      a.f0 = fparams[0];
      k1_pointwise<PW_BRIGHT><<<grid, 256, 0, st>>>(a); h->launches++; break;
    case FAV_CONTRAST: {
      FAV_REQUIRE(need_f(1), "contrast needs fparams[0]=c");
      FAV_REQUIRE(d_scratch && scratch_bytes >= fav_corrupt_scratch_bytes(corruption, n, height, width),
                  "contrast needs %zu scratch bytes", fav_corrupt_scratch_bytes(corruption, n, height, width));
      a.f0 = fparams[0];
      if ((per % 48) == 0 && per <= 48 * 64 && a.flags == 0 && (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0) {
        k1_contrast_small<<<(n + 7) / 8, 256, 0, st>>>(a);          // single pass: sums + apply from registers
        h->launches += 1;
        break;
      }
      if ((per % 48) == 0 && per <= 48 * 32 * 8 && (reinterpret_cast<uintptr_t>(d_src) & 15) == 0) {
        k1_channel_sums_small<<<(n + 7) / 8, 256, 0, st>>>(d_src, per, n, reinterpret_cast<unsigned long long*>(d_scratch));
      } else {
        FAV_CUDA_OK(cudaMemsetAsync(d_scratch, 0, (size_t)n * 24, st));
        const int bx = max(1, min(64, (height * width + 4095) / 4096));
        k1_channel_sums<<<dim3(n, bx), 256, 0, st>>>(d_src, per, reinterpret_cast<unsigned long long*>(d_scratch));
      }
      if (chunk_ok) k1_chunk16<PW_CONTRAST><<<grid16, 256, 0, st>>>(a);
      else k1_pointwise<PW_CONTRAST><<<grid, 256, 0, st>>>(a);
      h->launches += 2; break;
    }
    case FAV_FOG:
    case FAV_FROST: {
      const bool frost = corruption == FAV_FROST;
      FAV_REQUIRE(frost ? need_f(6) : need_f(2), frost ? "frost needs fparams = c0, c1, plasma decay, tint r, g, b"
                                                       : "fog needs fparams[0]=c, fparams[1]=wibble decay");
      int m = 1;
      while (m < max(height, width)) m *= 2;
      if (m <= 64 && !h->k1_legacy) {
        // small frames: plasma map in shared memory, one warp per image, map + statistics + application in ONE kernel
        PlasmaFusedArgs pf{};
        pf.src = d_src; pf.dst = d_dst; pf.n = n; pf.h = height; pf.w = width; pf.mapsize = m; pf.frost = frost ? 1 : 0;
        pf.f0 = fparams[0]; pf.f1 = frost ? fparams[1] : 0.f; pf.decay = frost ? fparams[2] : fparams[1];
        for (int c = 0; c < 3; ++c) { pf.tint[c] = frost ? fparams[3 + c] : 0.f; pf.mean[c] = mean[c]; pf.inv_std[c] = 1.0f / std[c]; }
        pf.k0 = k0; pf.k1 = k1; pf.first_image = uint32_t(first_image); pf.stream = a.stream; pf.flags = flags;
        if (m <= 32) {
          k1_plasma_fused<8><<<(n + 7) / 8, 256, (size_t)8 * m * m * 4, st>>>(pf);
        } else {
          if (!h->attr_plasma) {
            FAV_CUDA_OK(cudaFuncSetAttribute(k1_plasma_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * 64 * 4));
            h->attr_plasma = true;
          }
          k1_plasma_fused<4><<<(n + 3) / 4, 128, (size_t)4 * m * m * 4, st>>>(pf);
        }
        h->launches++; break;
      }
      FAV_REQUIRE(d_scratch && scratch_bytes >= fav_corrupt_scratch_bytes(corruption, n, height, width),
                  "fog / frost need %zu scratch bytes", fav_corrupt_scratch_bytes(corruption, n, height, width));
      a.f0 = fparams[0]; a.mapsize = m;
      if (frost) { a.f1 = fparams[1]; a.tint[0] = fparams[3]; a.tint[1] = fparams[4]; a.tint[2] = fparams[5]; }
      a.map_offset = (((size_t)n * 16) + 255) / 256 * 256;
      float* stats = reinterpret_cast<float*>(d_scratch);
      float* maps = reinterpret_cast<float*>(reinterpret_cast<char*>(d_scratch) + a.map_offset);
      k1_plasma<<<n, 1024, 0, st>>>(d_src, per, m, frost ? fparams[2] : fparams[1], k0, k1, uint32_t(first_image), a.stream, stats, maps);
      if (frost) k1_pointwise<PW_FROST><<<grid, 256, 0, st>>>(a);
      else k1_pointwise<PW_FOG><<<grid, 256, 0, st>>>(a);
      h->launches += 2; break;
    }
    case FAV_ELASTIC: {
      FAV_REQUIRE(need_f(5) && need_i(1) && d_table, "elastic_transform needs fparams[5], iparams[0]=radius and the Gaussian taps");
      FAV_REQUIRE(d_scratch && scratch_bytes >= fav_corrupt_scratch_bytes(corruption, n, height, width),
                  "elastic_transform needs %zu scratch bytes", fav_corrupt_scratch_bytes(corruption, n, height, width));
      ElasticArgs e{};
      e.src = d_src; e.out = out; e.n = n; e.h = height; e.w = width; e.radius = iparams[0];
      e.folded = (n_iparams >= 2 && iparams[1] != 0) ? 1 : 0;
      FAV_REQUIRE(e.radius >= 0 && table_bytes >= (e.folded ? ((size_t)width * width + (size_t)height * height) * 4 : (size_t)(2 * e.radius + 1) * 4),
                  "elastic_transform table too small");
      e.alpha = fparams[0]; e.mag = fparams[1]; e.c0 = fparams[2]; e.c1 = fparams[3]; e.sq = fparams[4];
      e.taps = reinterpret_cast<const float*>(d_table);
      e.k0 = k0; e.k1 = k1; e.first_image = uint32_t(first_image);
      e.stream = a.stream; e.aux_stream = stream_id(KIND_AUX, corruption, severity);
      e.scratch = reinterpret_cast<float*>(d_scratch); e.src_bgr = flags & FAV_SRC_BGR;
      const int g2 = grid_for((long long)n * height * width, 256, h->num_sms, 16);
      k1_elastic_prep<<<g2, 256, 0, st>>>(e);
      k1_elastic_xpass<<<g2, 256, 0, st>>>(e);
      k1_elastic_final<<<g2, 256, 0, st>>>(e);
      h->launches += 3; break;
    }
    case FAV_SNOW: {
      FAV_REQUIRE(need_f(6) && need_i(8) && d_table, "snow needs fparams[6], iparams[8] and the tap + zoom tables");
      FAV_REQUIRE(d_scratch && scratch_bytes >= fav_corrupt_scratch_bytes(corruption, n, height, width),
                  "snow needs %zu scratch bytes", fav_corrupt_scratch_bytes(corruption, n, height, width));
      SnowArgs sn{};
      sn.n = n; sn.h = height; sn.w = width;
      sn.loc = fparams[0]; sn.scale = fparams[1]; sn.thresh = fparams[2]; sn.ih_scale = fparams[5];
      sn.n_entries = iparams[0]; sn.max_taps = iparams[1];
      FAV_REQUIRE(table_bytes >= (size_t)iparams[7] + (size_t)(height + width) * 8, "snow table too small");
      sn.taps = reinterpret_cast<const uint8_t*>(d_table);
      sn.ztab = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(d_table) + iparams[7]);
      sn.k0 = k0; sn.k1 = k1; sn.first_image = uint32_t(first_image);
      sn.stream = a.stream; sn.aux_stream = stream_id(KIND_AUX, corruption, severity);
      sn.layer = reinterpret_cast<float*>(d_scratch);
      sn.blurred = sn.layer + (size_t)n * height * width;
      const int g2 = grid_for((long long)n * height * width, 256, h->num_sms, 16);
      k1_snow_layer<<<g2, 256, 0, st>>>(sn);
      k1_snow_blur<<<g2, 256, 0, st>>>(sn);
      a.f0 = fparams[3]; a.f1 = fparams[4];
      a.map_offset = (size_t)n * height * width * sizeof(float);       // pointwise reads the blurred plane
      k1_pointwise<PW_SNOW><<<grid, 256, 0, st>>>(a); h->launches += 3; break;
    }
    case FAV_JPEG: {
      FAV_REQUIRE(d_table && table_bytes >= 128 * 4, "jpeg_compression needs the 128-int quantisation table");
      FAV_REQUIRE(d_scratch && scratch_bytes >= fav_corrupt_scratch_bytes(corruption, n, height, width),
                  "jpeg_compression needs %zu scratch bytes", fav_corrupt_scratch_bytes(corruption, n, height, width));
      JpegArgs j{};
      j.src = d_src; j.out = out; j.n = n; j.h = height; j.w = width;
      j.mcu_x = (width + 15) / 16; j.mcu_y = (height + 15) / 16;
      FAV_REQUIRE(j.mcu_x * j.mcu_y <= 65535, "jpeg_compression: frame too large");
      j.table = reinterpret_cast<const int*>(d_table); j.src_bgr = flags & FAV_SRC_BGR;
      j.planes = reinterpret_cast<uint8_t*>(d_scratch);
      k1_jpeg_codec<<<dim3(n, j.mcu_x * j.mcu_y), 64, 0, st>>>(j);
      k1_jpeg_finish<<<grid_for((long long)n * height * width, 256, h->num_sms, 16), 256, 0, st>>>(j);
      h->launches += 2; break;
    }
    case FAV_GLASS_BLUR: {
      FAV_REQUIRE(need_i(3) && d_table, "glass_blur needs iparams = delta, iterations, radius and the tap table");
      GlassArgs g{};
      g.src = d_src; g.out = out; g.n = n; g.h = height; g.w = width;
      g.delta = iparams[0]; g.iters = iparams[1]; g.radius = iparams[2];
      FAV_REQUIRE(g.delta >= 1 && g.delta <= 7 && g.iters >= 1 && g.radius >= 0 && 2 * g.delta < min(height, width) &&
                  width - 2 * g.delta <= GLASS_CHUNK, "glass_blur: bad parameters");
      FAV_REQUIRE(table_bytes >= (size_t)(2 * g.radius + 1) * 8, "glass_blur tap table too small");
      g.table = reinterpret_cast<const int*>(d_table);
      g.k0 = k0; g.k1 = k1; g.first_image = uint32_t(first_image); g.stream = a.stream; g.src_bgr = flags & FAV_SRC_BGR;
      const size_t smem = (size_t)height * width * 4 + GLASS_CHUNK + (size_t)(2 * g.radius + 1) * 8;
      FAV_REQUIRE(smem <= 226 * 1024, "glass_blur keeps the image in shared memory: %dx%d is too large", height, width);
      // small images: too few (band, column) items to fill a CTA with the band scheme
      if (height * width > 64 * 64) {
        if (smem > 48 * 1024) FAV_CUDA_OK(cudaFuncSetAttribute(k1_glass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        k1_glass<true><<<n, 512, smem, st>>>(g);      // one CTA per SM there (the image fills shared memory): 16 warps for the blurs
      } else {
        if (smem > 48 * 1024) FAV_CUDA_OK(cudaFuncSetAttribute(k1_glass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        k1_glass<false><<<n, 256, smem, st>>>(g);
      }
      h->launches++; break;
    }
    case FAV_DEFOCUS_BLUR:
    case FAV_MOTION_BLUR: {
      FAV_REQUIRE(need_i(7) && d_table, "tap stencil needs iparams[0..6] and a tap table");
      TapArgs t{};
      t.src = d_src; t.out = out; t.n = n; t.h = height; t.w = width;
      t.n_entries = iparams[0]; t.max_taps = iparams[1]; t.border = iparams[2];
      t.dy_min = iparams[3]; t.dy_max = iparams[4]; t.dx_min = iparams[5]; t.dx_max = iparams[6];
      FAV_REQUIRE(t.n_entries >= 1 && t.max_taps >= 1 && t.dy_min <= 0 && t.dy_max >= 0 && t.dx_min <= 0 && t.dx_max >= 0,
                  "tap stencil: bad table geometry");
      FAV_REQUIRE(table_bytes >= (size_t)t.n_entries * (16 + 8 * (size_t)t.max_taps), "tap table too small");
      t.table = reinterpret_cast<const uint8_t*>(d_table);
      t.k0 = k0; t.k1 = k1; t.first_image = uint32_t(first_image);
      t.stream = stream_id(KIND_AUX, corruption, severity);
      t.tiles_x = (width + TAP_TILE - 1) / TAP_TILE; t.tiles_y = (height + TAP_TILE - 1) / TAP_TILE;
      t.src_bgr = flags & FAV_SRC_BGR;
      const bool src16 = (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (per & 15) == 0;
      t.raw_stage = (!h->k1_legacy && t.tiles_x * t.tiles_y == 1 && src16) ? 1 : 0;
      t.out_staged = (!h->k1_legacy && !(flags & FAV_OUT_F32) && ((size_t)width * 6) % 16 == 0 && ((size_t)per * 2) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(d_dst) & 15) == 0) ? 1 : 0;
      // register-tiled dense loop: 5 shared-memory wavefronts per (halo row, box column) of a warp against 17 per list tap;
      // it walks box_h + 3 halo rows, so a 3 x 3 box (twice the FMAs of its nine list taps) stays on the list loop
      // (measured: 3 x 3 list 0.52 ms vs dense 0.61 ms per 65 536 CIFAR frames; 5 x 5 dense 0.81 vs list 0.84-0.96;
      // 9 x 9 .. 21 x 21 at 224 x 224 dense x1.4 .. x2.2, tools/k1_stencil_bench.py)
      const long long box_w = t.dx_max - t.dx_min + 1, box_h = t.dy_max - t.dy_min + 1;
      t.dense = (!h->k1_legacy && !h->k1_list_stencil && t.n_entries == 1 && box_h >= 5 &&
                 5 * (box_h + 3) * box_w < 17LL * t.max_taps) ? 1 : 0;
      const size_t smem = (size_t)(TAP_TILE + t.dx_max - t.dx_min) * (TAP_TILE + t.dy_max - t.dy_min) * 16 + (size_t)t.max_taps * 8 +
                          (t.dense ? (size_t)(box_h + 3) * box_w * 16 : 0) +
                          (size_t)(2 * TAP_TILE + t.dx_max - t.dx_min + t.dy_max - t.dy_min) * 4 + (t.raw_stage ? (size_t)per + 16 : 0);
      FAV_REQUIRE(smem <= 200 * 1024, "tap stencil halo too large (%zu B of shared memory)", smem);
      if (smem > 48 * 1024) FAV_CUDA_OK(cudaFuncSetAttribute(k1_taps, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      k1_taps<<<dim3(n, t.tiles_x * t.tiles_y), 256, smem, st>>>(t); h->launches++; break;
    }
    case FAV_ZOOM_BLUR: {
      FAV_REQUIRE(need_i(1) && d_table, "zoom_blur needs iparams[0]=nz and a table");
      FAV_REQUIRE(table_bytes >= (size_t)iparams[0] * (height + width) * 8, "zoom table too small");
      ZoomArgs z{};
      z.src = d_src; z.out = out; z.n = n; z.h = height; z.w = width; z.nz = iparams[0];
      z.table = reinterpret_cast<const uint2*>(d_table); z.src_bgr = flags & FAV_SRC_BGR;
      k1_zoom<<<grid_for((long long)n * height * width, 256, h->num_sms, 16), 256, 0, st>>>(z); h->launches++; break;
    }
    case FAV_PIXELATE: {
      FAV_REQUIRE(need_i(4) && d_table, "pixelate needs iparams = sw, sh, kh, kv and the coefficient table");
      PixArgs p{};
      p.src = d_src; p.out = out; p.n = n; p.h = height; p.w = width;
      p.sw = iparams[0]; p.sh = iparams[1]; p.kh = iparams[2]; p.kv = iparams[3];
      FAV_REQUIRE(p.sw >= 1 && p.sw <= width && p.sh >= 1 && p.sh <= height && p.kh >= 1 && p.kv >= 1, "pixelate: bad geometry");
      FAV_REQUIRE(table_bytes >= ((size_t)p.sw * (2 + p.kh) + (size_t)p.sh * (2 + p.kv) + height + width) * 4, "pixelate table too small");
      FAV_REQUIRE(d_scratch && scratch_bytes >= (size_t)n * p.sh * p.sw * 3, "pixelate needs %zu scratch bytes", (size_t)n * p.sh * p.sw * 3);
      p.table = reinterpret_cast<const int*>(d_table); p.src_bgr = flags & FAV_SRC_BGR;
      p.small = reinterpret_cast<uint8_t*>(d_scratch);
      k1_pixelate_down<<<grid_for((long long)n * p.sh * p.sw, 256, h->num_sms, 16), 256, 0, st>>>(p);
      k1_pixelate_up<<<grid_for((long long)n * height * width, 256, h->num_sms, 16), 256, 0, st>>>(p);
      h->launches += 2; break;
    }
    default:
      set_error("corruption id %d is not implemented on the device yet", corruption);
      return FAV_E_UNSUPPORTED;
  }
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}

// The SURVEY.md 8(b) signature: everything per-cell is derived inside the library (tables.cu) and cached in the handle.
extern "C" int fav_corrupt_normalize(fav_handle h, const uint8_t* d_src, void* d_dst, int n, int height, int width,
                                     int corruption, int severity, uint64_t seed, uint64_t first_image,
                                     const float mean[3], const float std[3], unsigned flags, void* stream) {
  FAV_REQUIRE(h, "fav_corrupt_normalize: null handle");
  FAV_REQUIRE(height > 0 && width > 0, "fav_corrupt_normalize: bad shape h=%d w=%d", height, width);
  FAV_REQUIRE(corruption >= FAV_CLEAN && corruption <= FAV_JPEG, "unknown corruption id %d", corruption);
  FAV_REQUIRE(corruption == FAV_CLEAN || (severity >= 1 && severity <= 5), "severity must be 1..5 (got %d)", severity);
  FAV_DEVICE(h);
  if (corruption == FAV_CLEAN)
    return fav_corrupt_normalize_ex(h, d_src, d_dst, n, height, width, corruption, severity, nullptr, 0, nullptr, 0, nullptr, 0,
                                    nullptr, 0, seed, first_image, mean, std, flags, stream);
  const K1Entry* e = nullptr;
  int rc = k1_lookup(h, corruption, severity, height, width, profile_of(flags, height, width),
                     reinterpret_cast<cudaStream_t>(stream), &e);
  if (rc != FAV_OK) return rc;
  const size_t sb = fav_corrupt_scratch_bytes(corruption, n, height, width);
  void* scratch = nullptr;
  if (sb) { rc = k1_scratch(h, sb, &scratch); if (rc != FAV_OK) return rc; }
  return fav_corrupt_normalize_ex(h, d_src, d_dst, n, height, width, corruption, severity, e->p.fp.data(), int(e->p.fp.size()),
                                  e->p.ip.data(), int(e->p.ip.size()), e->d_table, e->p.table.size(), scratch, sb, seed,
                                  first_image, mean, std, flags, stream);
}

extern "C" int fav_synth_images(fav_handle h, uint8_t* d_dst, int n, int height, int width, uint64_t seed,
                                uint64_t first_image, void* stream) {
  FAV_REQUIRE(h && d_dst && n >= 0 && height > 0 && width > 0, "fav_synth_images: bad arguments");
  if (n == 0) return FAV_OK;
  const int per = height * width * 3;
  const long long work = (long long)n * ((per + 15) / 16);
  k_synth_images<<<grid_for(work, 256, h->num_sms, 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_dst, n, per, uint32_t(seed), uint32_t(seed >> 32), uint32_t(first_image));
  h->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}

extern "C" int fav_synth_labels(fav_handle h, int32_t* d_dst, int n, int C, uint64_t seed, uint64_t first_image,
                                void* stream) {
  FAV_REQUIRE(h && d_dst && n >= 0 && C > 0, "fav_synth_labels: bad arguments");
  if (n == 0) return FAV_OK;
  k_synth_labels<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_dst, n, C, uint32_t(seed), uint32_t(seed >> 32), uint32_t(first_image));
  h->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}
