// conv_flat.cu -- K2, flat-padded resident variant for 3x3 / stride 1 / pad 1 convolutions with Cin = Cout = 64 on small
// images (ResNet-18 layer1 at 8x8: a third of the sweep's time in the generic kernel, which is L2->SM bandwidth bound
// because every filter tap re-streams the activation tile and the weight tile: 216 KB per 128 output pixels).
//
// Here a group of G images is loaded ONCE by a single TMA box (64 ch, W+1, H+1, G) whose origin is (-1,-1): TMA's
// out-of-bounds zero fill writes one shared zero row above and one shared zero column left of every image, so shared
// memory holds a flat list of padded pixels, 128 B (64 bf16 channels) each, in exactly the canonical K-major
// SWIZZLE_128B operand layout.  Output "rows" of the GEMM are consecutive padded pixels (padding pixels included and
// discarded: (H+1)(W+1)/(HW) = 1.27x MMA work at 8x8), and filter tap (r,s) is the SAME shared-memory slab read through a
// UMMA descriptor whose start address is shifted by ((r-1)(W+1) + (s-1)) pixels.  The 9x64x64 weights (72 KB) are
// loaded once per CTA and stay resident.  L2->SM traffic drops from 216 KB to ~20 KB per 128 output pixels.
//
// Band mode (images larger than a slab, e.g. ResNet-50 layer1 at 56x56): a slab is a band of bh output rows of ONE image plus
// the row above and the row below -- one TMA box (64 ch, W+1, bh+2, 1) with origin (-1, y0-1); out-of-bounds rows / the
// column at x = -1 arrive as zeros.  The MMA tile starts at the band's second padded row (slab row W+1), taps shift exactly as
// above.  56x56: 228 padded pixels loaded per 112 outputs (29 KB instead of the generic kernel's 9 x 16 KB).
//
// Replaces (reference): nothing executable (see conv.cu header); oracle twin oracle/model.py.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "conv.cuh"
#include "tc_ptx.cuh"

namespace fav {

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box);

constexpr int FL_SLAB_ROWS = 256;                       // padded pixels per slab buffer (two 128-row MMA tiles)
constexpr int FL_SLAB_BYTES = FL_SLAB_ROWS * 128;       // 32 KiB
constexpr int FL_NSLAB = 4;
constexpr int FL_MARGIN = 16 * 128;                     // rows readable before slab 0 / after the last slab
constexpr int FL_W_BYTES = 9 * 64 * 128;                // resident weights: 9 taps x 64 cout x 64 cin bf16
constexpr int FL_EPI_WARPS = 16;
constexpr int FL_THREADS = 32 * (2 + FL_EPI_WARPS);
constexpr int FL_NACC = 4;                              // TMEM accumulators of 64 columns

struct FlatArgs {
  void* y;
  const float* bias;
  const __nv_bfloat16* res;
  int P, H, W, Wp;             // images, image size, padded pitch W + 1
  int rows_img, G, slab_rows;  // padded pixels per image, images per slab, G * rows_img (<= 256)
  int n_slabs, n_tiles;        // slabs in the problem, 128-row tiles per slab (1 or 2)
  int band, bh, bands, row0;   // band mode: output rows per band, bands per image, first slab row of the MMA tiles (= Wp)
  int relu, T, rep, drop;
  uint32_t drop_thr8, drop_add4, drop_hi4;   // round(p * 256); operands of dropout_keep4 (common.cuh)
  PhiloxKeys drop_keys;                // round keys of (k0, k1)
  float drop_scale;
  uint32_t k0, k1, first_image, drop_stream;
  uint32_t idesc;
  int base_offset_mode;        // 1: set the descriptor's matrix-base-offset field to (addr >> 7) & 7
  unsigned long long* stats;
};

__device__ __forceinline__ uint64_t flat_desc(uint32_t saddr, int base_offset_mode) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(FL_THREADS, 1)
conv3x3_flat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const FlatArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to_1k = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to_1k;
  const uint32_t smem_base = smem_u32(smem);
  // layout: [margin 2 KiB][slab 0..3, 32 KiB each][margin 2 KiB][weights 72 KiB][barriers]
  const uint32_t slab0 = smem_base + FL_MARGIN;
  const uint32_t w_base = slab0 + FL_NSLAB * FL_SLAB_BYTES + FL_MARGIN;     // 2048 + 131072 + 2048 = 1024-aligned
  const uint32_t bars = w_base + FL_W_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bars - smem_base) + (1 + 2 * FL_NSLAB + 2 * FL_NACC) * 8);
  const uint32_t w_full = bars;
  auto slab_full = [&](int s) { return bars + 8u * (1 + s); };
  auto slab_empty = [&](int s) { return bars + 8u * (1 + FL_NSLAB + s); };
  auto tfull = [&](int i) { return bars + 8u * (1 + 2 * FL_NSLAB + i); };
  auto tempty = [&](int i) { return bars + 8u * (1 + 2 * FL_NSLAB + FL_NACC + i); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float s_bias[64];
  if (threadIdx.x < 64) s_bias[threadIdx.x] = a.bias[threadIdx.x];

  // zero the margins and every slab's tail rows [slab_rows, 256): they act as the zero halo below the last image of a slab
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < FL_MARGIN / 16; i += blockDim.x) {
      z[i] = zero;
      z[(FL_MARGIN + FL_NSLAB * FL_SLAB_BYTES) / 16 + i] = zero;
    }
    const int tail16 = (FL_SLAB_ROWS - a.slab_rows) * 8;
    for (int s = 0; s < FL_NSLAB; ++s)
      for (int i = threadIdx.x; i < tail16; i += blockDim.x)
        z[(FL_MARGIN + s * FL_SLAB_BYTES + a.slab_rows * 128) / 16 + i] = zero;
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1u);
    for (int s = 0; s < FL_NSLAB; ++s) { mbar_init(slab_full(s), 1u); mbar_init(slab_empty(s), 1u); }
    for (int i = 0; i < FL_NACC; ++i) { mbar_init(tfull(i), 1u); mbar_init(tempty(i), uint32_t(FL_EPI_WARPS / 2)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 64 * FL_NACC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on we touch activations it produced (and buffers it may still be reading)
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================================================================= TMA issuer: weights once, then one box per slab
    {
      long long w_empty = 0;
      const long long t_begin = clock64();
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full, FL_W_BYTES);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_base + tap * 8192, &tmW, w_full, tap * 64, 0);
      }
      __syncwarp();
      int s = 0, ph = 0;
      const uint32_t bytes = uint32_t(a.slab_rows) * 128u;
      for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
        mbar_wait_timed(slab_empty(s), ph ^ 1, w_empty, a.stats != nullptr);
        if (elect_one()) {
          mbar_arrive_expect_tx(slab_full(s), bytes);
          if (a.band) {
            const int q = slab / a.bands, b = slab - q * a.bands;
            tma_load_4d(slab0 + s * FL_SLAB_BYTES, &tmA, slab_full(s), 0, -1, b * a.bh - 1, q);
          } else {
            tma_load_4d(slab0 + s * FL_SLAB_BYTES, &tmA, slab_full(s), 0, -1, -1, slab * a.G);
          }
        }
        __syncwarp();
        if (++s == FL_NSLAB) { s = 0; ph ^= 1; }
      }
      if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], (unsigned long long)w_empty);
        atomicAdd(&a.stats[1], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&a.stats[7], 1ull);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (whole warp converged; one elected lane issues)
    {
      long long w_full_t = 0, w_tempty = 0;
      const long long t_begin = clock64();
      mbar_wait(w_full, 0);
      int s = 0, ph = 0, ai = 0, aph = 0;
      for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
        mbar_wait_timed(slab_full(s), ph, w_full_t, a.stats != nullptr);
        for (int tile = 0; tile < a.n_tiles; ++tile) {
          mbar_wait_timed(tempty(ai), aph ^ 1, w_tempty, a.stats != nullptr);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(ai * 64);
          const uint32_t a_tile = slab0 + s * FL_SLAB_BYTES + (a.row0 + tile * 128) * 128;
          if (elect_one()) {
            uint32_t accumulate = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const uint64_t da_r = flat_desc(a_tile + ((r - 1) * a.Wp - 1) * 128, a.base_offset_mode);
              const uint64_t db_r = make_sw128_desc(w_base + r * 3 * 8192);
#pragma unroll
              for (int ss = 0; ss < 3; ++ss) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // +8 per pixel of shift, +512 per tap of weights, +2 per 32 bytes of K
                  umma_f16(d_tmem, da_r + uint64_t(ss * 8 + 2 * k), db_r + uint64_t(ss * 512 + 2 * k), a.idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            umma_commit(tfull(ai));
          }
          __syncwarp();
          if (++ai == FL_NACC) { ai = 0; aph ^= 1; }
        }
        if (elect_one()) umma_commit(slab_empty(s));
        __syncwarp();
        if (++s == FL_NSLAB) { s = 0; ph ^= 1; }
      }
      if (a.stats && lane == 0) {
        atomicAdd(&a.stats[2], (unsigned long long)w_full_t);
        atomicAdd(&a.stats[3], (unsigned long long)w_tempty);
        atomicAdd(&a.stats[4], (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ================================================================= epilogue: two groups of 8 warps take alternate tiles;
    // inside a group, warp = (TMEM lane quarter, 32-column half).  Two independent 16-column chunks per warp give the
    // residual loads / Philox chains of a tile some ILP, and the other group's tile overlaps them further.
    const int ew = warp - 2, group = ew >> 3, quarter = warp & 3, half = (ew >> 2) & 1;
    const int row = quarter * 32 + lane;
    const int hw_img = a.H * a.W, n_rep = a.rep > 1 ? a.rep : 1, Cout = 64;
    const int cbase = half * 32;
    long long w_tfull = 0;
    const long long t_begin = clock64();
    int tc = 0;                                         // tile sequence number inside this CTA (accumulator = tc % 4)
    for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
      for (int tile = 0; tile < a.n_tiles; ++tile, ++tc) {
        if ((tc & 1) != group) continue;
        const int ai = tc & (FL_NACC - 1), aph = (tc / FL_NACC) & 1;
        // everything that does not depend on the accumulator happens BEFORE the wait: row decode, residual loads,
        // the first replica's dropout masks -- so their latency overlaps the MMAs of this tile
        const int pi = tile * 128 + row;
        int q, hw;
        bool valid;
        if (a.band) {                                       // slab = (image, band of bh rows); tile row pi = padded pixel row0 + pi
          q = slab / a.bands;
          const int y0 = (slab - q * a.bands) * a.bh;
          const int pb = a.row0 + pi, ph_ = pb / a.Wp, pw_ = pb - ph_ * a.Wp, y = y0 + ph_ - 1;
          valid = ph_ >= 1 && ph_ <= a.bh && pw_ != 0 && y < a.H;
          hw = y * a.W + (pw_ - 1);
        } else {
          const int g = pi / a.rows_img, rem = pi - g * a.rows_img, ph_ = rem / a.Wp, pw_ = rem - ph_ * a.Wp;
          q = slab * a.G + g;
          valid = pi < a.slab_rows && ph_ != 0 && pw_ != 0 && q < a.P;
          hw = (ph_ - 1) * a.W + (pw_ - 1);
        }
        uint4 rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[i] = make_uint4(0, 0, 0, 0);
        if (valid && a.res) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + ((size_t)q * hw_img + hw) * Cout + cbase);
#pragma unroll
          for (int i = 0; i < 4; ++i) rv[i] = __ldg(rp + i);
        }
        const int n_img = a.rep > 1 ? q : q / a.T;
        const uint32_t e16 = uint32_t(hw * Cout + cbase) >> 4;
        // MC-dropout on the packed bf16 pairs: two Philox calls give thirty-two byte lanes (contract: common.cuh);
        // dropout_keep4 + PRMT sign replication make 0xFFFF-per-kept-channel masks that are ANDed on (dropped -> +0.0)
        auto keep_words = [&](int tt, uint32_t (&kw)[16]) {
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint4 r = philox4x32_10_keys(e16 + c2, a.first_image + uint32_t(n_img), uint32_t(tt), a.drop_stream, a.drop_keys);
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t k4 = dropout_keep4(rw[i], a.drop_add4, a.drop_hi4);
              kw[8 * c2 + 2 * i] = dropout_pair_lo(k4);
              kw[8 * c2 + 2 * i + 1] = dropout_pair_hi(k4);
            }
          }
        };
        uint32_t kw[16];                                     // first pass's keep-words: computed BEFORE the accumulator wait
#pragma unroll
        for (int i = 0; i < 16; ++i) kw[i] = 0xFFFFFFFFu;
        if (valid && a.drop) keep_words(a.rep > 1 ? 0 : q - n_img * a.T, kw);

        mbar_wait_timed(tfull(ai), aph, w_tfull, a.stats != nullptr);
        tc_fence_after();
        uint32_t acc[32];
        {
          uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&acc[0]);
          uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&acc[16]);
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(ai * 64 + cbase);
          tmem_ld16(taddr, lo);
          tmem_ld16(taddr + 16, hi);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(ai));          // accumulator is in registers: release it right away
        if (!valid) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t rw = reinterpret_cast<const uint32_t*>(rv)[i];
          v[2 * i] = __uint_as_float(acc[2 * i]) + s_bias[cbase + 2 * i] + bf16_lo(rw);
          v[2 * i + 1] = __uint_as_float(acc[2 * i + 1]) + s_bias[cbase + 2 * i + 1] + bf16_hi(rw);
        }
        if (a.drop) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= a.drop_scale;
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (a.relu) {                                        // after the pack: one instruction per pair (relu_bf16x2)
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = relu_bf16x2(pk[i]);
        }
        for (int rp = 0; rp < n_rep; ++rp) {
          const int p_out = a.rep > 1 ? q * a.rep + rp : q;
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = pk[i] & kw[i];
          if (a.drop && rp + 1 < n_rep) keep_words(rp + 1, kw);
          uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + ((size_t)p_out * hw_img + hw) * Cout + cbase);
#pragma unroll
          for (int k = 0; k < 4; ++k) yp[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        }
      }
    }
    if (a.stats && warp == 2 && lane == 0) {
      atomicAdd(&a.stats[5], (unsigned long long)w_tfull);
      atomicAdd(&a.stats[6], (unsigned long long)(clock64() - t_begin));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64 * FL_NACC);
}

// band mode: output rows per band -- the band plus its two halo rows must leave a zero tail row in the 256-row slab buffer and
// the band's padded pixels must fit one 128-row MMA tile
static int flat_band_rows(int w) {
  const int wp = w + 1;
  const int by_slab = (FL_SLAB_ROWS - 1) / wp - 2, by_tile = 128 / wp;
  return by_slab < by_tile ? by_slab : by_tile;
}

static int flat_mode_env() {
  static const int v = [] { const char* e = getenv("FAV_FLAT"); return e ? atoi(e) : 1; }();   // 0 disables, 3 = set the base-offset field
  return v;
}

bool conv_flat_applicable(const ConvCall& c) {
  const ConvLayer& L = *c.L;
  const bool layer_ok = L.r == 3 && L.s == 3 && L.stride == 1 && L.pad == 1 && L.cin == 64 && L.cout == 64 && L.k2pad == 0 &&
                        !L.fold && !L.cin_store && !c.out_f32 && L.bn == 64;
  const bool whole_ok = (c.h + 1) * (c.w + 1) <= FL_SLAB_ROWS && c.w + 1 <= 256 && c.h + 1 <= 256;
  static const int env_band = [] { const char* e = getenv("FAV_FLAT_BAND"); return e ? atoi(e) : 1; }();
  const bool band_ok = !whole_ok && env_band && flat_band_rows(c.w) >= 1 && (long long)c.p * c.h < (1ll << 30);
  if (!layer_ok || !(whole_ok || band_ok)) return false;
  if (c.a_mode == 4) return true;
  if (c.rep > 1) return false;      // the replica-writing epilogue dominates there and padding rows would idle half its lanes
  return c.a_mode < 0 && flat_mode_env() != 0;
}

int conv_flat_launch(Ctx* ctx, const ConvCall& c, cudaStream_t st) {
  const ConvLayer& L = *c.L;
  FAV_REQUIRE(L.tmap_ok, "conv: layer not finalized");
  FlatArgs a{};
  a.y = c.y; a.bias = L.bias; a.res = reinterpret_cast<const __nv_bfloat16*>(c.res);
  a.P = c.p; a.H = c.h; a.W = c.w; a.Wp = c.w + 1;
  a.rows_img = (c.h + 1) * (c.w + 1);
  a.band = a.rows_img > FL_SLAB_ROWS ? 1 : 0;
  if (a.band) {
    a.bh = flat_band_rows(c.w);
    FAV_REQUIRE(a.bh >= 1, "conv: image too wide for the flat-padded 3x3 kernel (W = %d)", c.w);
    a.bands = (c.h + a.bh - 1) / a.bh;
    a.G = 1;
    a.slab_rows = (a.bh + 2) * a.Wp;          // the TMA box; rows [slab_rows, 256) stay zero (halo right of the last halo row)
    a.n_slabs = c.p * a.bands;
    a.n_tiles = 1;
    a.row0 = a.Wp;
  } else {
    a.G = FL_SLAB_ROWS / a.rows_img;
    // small problems: fewer images per slab so that every SM gets the same number of slabs (the epilogue dominates there)
    if (a.G > 1 && c.p / a.G < 4 * ctx->num_sms) { int g = c.p / (4 * ctx->num_sms); a.G = g < 1 ? 1 : (g < a.G ? g : a.G); }
    if (a.G > c.p) a.G = c.p;
    if (a.G > 256) a.G = 256;
    a.slab_rows = a.G * a.rows_img;
    a.n_slabs = (c.p + a.G - 1) / a.G;
    a.n_tiles = (a.slab_rows + 127) / 128;
  }
  a.relu = c.relu; a.T = c.T > 0 ? c.T : 1; a.rep = c.rep > 1 ? c.rep : 1; a.drop = c.drop;
  if (c.drop) {
    FAV_REQUIRE(c.p_drop >= 0.f && c.p_drop < 1.f, "conv: p_drop must be in [0,1)");
    a.drop_thr8 = dropout_thr8(c.p_drop);
    a.drop_add4 = dropout_add4(a.drop_thr8); a.drop_hi4 = dropout_hi4(a.drop_thr8);
    a.drop_scale = dropout_scale8(a.drop_thr8);
    a.k0 = uint32_t(c.seed); a.k1 = uint32_t(c.seed >> 32); a.first_image = uint32_t(c.first_image);
    a.drop_keys = philox_keys(a.k0, a.k1);
    a.drop_stream = stream_id(KIND_DROPOUT, c.layer_id, 0);
  }
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(64 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  // measured on B200: the tensor core derives the 128B-swizzle phase from the absolute shared-memory address, so a
  // descriptor whose start is only 128-byte aligned must leave the matrix-base-offset field at 0 (setting it gives wrong results)
  a.base_offset_mode = flat_mode_env() == 3 ? 1 : 0;
  CUtensorMap tmA;
  memset(&tmA, 0, sizeof(tmA));
  {
    const cuuint64_t dims[4] = {64, (cuuint64_t)c.w, (cuuint64_t)c.h, (cuuint64_t)c.p};
    const cuuint64_t strides[3] = {128, (cuuint64_t)c.w * 128, (cuuint64_t)c.h * c.w * 128};
    const cuuint32_t box[4] = {64, (cuuint32_t)(c.w + 1), (cuuint32_t)(a.band ? a.bh + 2 : c.h + 1), (cuuint32_t)a.G};
    int rc = encode_map(&tmA, c.x, 4, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + 2 * FL_MARGIN + FL_NSLAB * FL_SLAB_BYTES + FL_W_BYTES + 256;
  if (!ctx->attr_flat) {
    FAV_CUDA_OK(cudaFuncSetAttribute(conv3x3_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));   // + 256 B static
    ctx->attr_flat = true;
  }
  const int grid = a.n_slabs < ctx->num_sms ? a.n_slabs : ctx->num_sms;
  cudaEvent_t e1 = nullptr;
  {
    const long long M = (long long)c.p * c.h * c.w;
    int rc = conv_timing_begin(ctx, st, float(2.0 * double(M) * 576.0 * 64.0 * 1e-9),
                               float((double(M) * 64 * 2 * (2 + (c.res ? 1 : 0)) + 576.0 * 64 * 2) * 1e-9), &e1, &a.stats);
    if (rc) return rc;
  }
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    FAV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv3x3_flat_kernel, tmA, *reinterpret_cast<const CUtensorMap*>(L.tmap_w), a));
  }
  if (e1) FAV_CUDA_OK(cudaEventRecord(e1, st));
  ctx->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
