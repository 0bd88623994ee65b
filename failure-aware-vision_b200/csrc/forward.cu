// forward.cu -- weights arena, layer plan and the batched MC-dropout forward (stem -> blocks -> pool -> fc).
//
// Replaces (reference): the ML-score slot documented at platform/backend/anomaly_simulator.py:34-77 ("No PyTorch
// dependency ... heuristic proxy") with a real classifier forward; topology = stock torchvision ResNet (oracle/model.py).
// MC-dropout: a mask after every residual block and on the pooled feature; everything before the
// first mask (stem, max-pool, block 0) is pass-invariant and runs once per image, block 0's epilogue then writes
// the T masked replicas, and all later layers run on N*T pass-images as one batched launch sequence.
#include <cstring>
#include "conv.cuh"

namespace fav {

struct BlockDesc { int n_convs; int ds; int conv0; bool fused_ds = false; };   // conv0 = index of the block's first conv; ds = index or -1

struct Plan {
  int model_id = 0, num_classes = 0, in_h = 0, in_w = 0;
  std::vector<ConvLayer> convs;      // [0] stem, blocks..., [last] fc (1x1)
  std::vector<BlockDesc> blocks;
  void* arena = nullptr;             // device weights
  size_t arena_bytes = 0;
  // workspace geometry (set by fav_reserve)
  int max_images = 0, max_T = 0;
  size_t buf_bytes = 0;              // each of the 5 activation buffers
};

void plan_destroy(Plan* p) {
  if (!p) return;
  if (p->arena) cudaFree(p->arena);
  delete p;
}

// ------------------------------------------------------------------------------------------ small kernels
// 3x3 / stride 2 / pad 1 max-pool over bf16 NHWC, 8 channels (16 B) per thread.  The maximum of bf16 values needs no
// unpacking (max.bf16x2 on the packed pairs is exact); indices are 32-bit (the launcher checks the element count).
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__global__ void __launch_bounds__(256) k_maxpool3x3s2(const uint4* __restrict__ x, uint4* __restrict__ y, int P, int H,
                                                      int W, int C8, int OH, int OW) {
  const uint32_t total = uint32_t(P) * uint32_t(OH) * uint32_t(OW) * uint32_t(C8);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t r = i / uint32_t(C8);
    const int c = int(i - r * uint32_t(C8));
    const uint32_t r2 = r / uint32_t(OW);
    const int ow = int(r - r2 * uint32_t(OW));
    const int p = int(r2 / uint32_t(OH));
    const int oh = int(r2 - uint32_t(p) * uint32_t(OH));
    uint4 m = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);      // -inf in every bf16 lane
    const uint4* img = x + (size_t)p * H * W * C8 + c;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = oh * 2 - 1 + dy;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int iw = ow * 2 - 1 + dx;
        if (iw < 0 || iw >= W) continue;
        const uint4 v = __ldg(img + (size_t)(ih * W + iw) * C8);
        m.x = max_bf16x2(m.x, v.x); m.y = max_bf16x2(m.y, v.y); m.z = max_bf16x2(m.z, v.z); m.w = max_bf16x2(m.w, v.w);
      }
    }
    y[i] = m;
  }
}

// global average pool + MC-dropout on the pooled feature: [P, HW, C] bf16 -> [P, C] bf16
__global__ void __launch_bounds__(256) k_pool_dropout(const uint4* __restrict__ x, uint4* __restrict__ y, int P, int HW,
                                                      int C8, int T, int drop, uint32_t thr8, float scale, uint32_t k0,
                                                      uint32_t k1, uint32_t first_image, uint32_t stream) {
  const long long total = (long long)P * C8;
  const float inv = 1.0f / float(HW);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C8), p = int(i / C8);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < HW; ++j) {
      const uint4 v = __ldg(x + ((size_t)p * HW + j) * C8 + c);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[2 * k] += bf16_lo(w4[k]); s[2 * k + 1] += bf16_hi(w4[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = HW == 1 ? s[k] : s[k] * inv;
    if (drop) {
      const int n_img = p / T, t = p - n_img * T;
      // this thread's 8 channels are bytes 8 (c & 1) .. + 7 of a 16-channel Philox chunk (contract: common.cuh)
      const uint4 r = philox4x32_10(uint32_t(c) >> 1, first_image + uint32_t(n_img), uint32_t(t), stream, k0, k1);
      const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
      const int w0 = 2 * (c & 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = ((rw[w0 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu) >= thr8 ? s[k] * scale : 0.f;
    }
    y[i] = make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
  }
}

// same for large feature maps (the batch-1 streaming gate: 15x20 pixels, one image): a CTA per (pass-image, 32 channel
// vectors); its 8 warps sum interleaved pixels, shared-memory reduce, warp 0 applies the mask
__global__ void __launch_bounds__(256) k_pool_dropout_wide(const uint4* __restrict__ x, uint4* __restrict__ y, int P, int HW,
                                                           int C8, int T, int drop, uint32_t thr8, float scale, uint32_t k0,
                                                           uint32_t k1, uint32_t first_image, uint32_t stream) {
  __shared__ float part[8][32][8];
  const int groups = (C8 + 31) / 32;
  const int p = blockIdx.x / groups, c = (blockIdx.x % groups) * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < C8) {
    for (int j = w; j < HW; j += 8) {
      const uint4 v = __ldg(x + ((size_t)p * HW + j) * C8 + c);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[2 * k] += bf16_lo(w4[k]); s[2 * k + 1] += bf16_hi(w4[k]); }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) part[w][threadIdx.x & 31][k] = s[k];
  __syncthreads();
  if (w != 0 || c >= C8) return;
  const float inv = 1.0f / float(HW);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += part[ww][threadIdx.x][k];
    s[k] = t * inv;
  }
  if (drop) {
    const int n_img = p / T, t = p - n_img * T;
    const uint4 r = philox4x32_10(uint32_t(c) >> 1, first_image + uint32_t(n_img), uint32_t(t), stream, k0, k1);
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
    const int w0 = 2 * (c & 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = ((rw[w0 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu) >= thr8 ? s[k] * scale : 0.f;
  }
  y[(size_t)p * C8 + c] = make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
}

// bf16 NHWC3 -> 2x2 space-to-depth of the zero-padded image, [n][hp][wp][16]: channel (dy*2 + dx)*4 + c of pixel (Y, X) is
// image pixel (2Y + dy - pad, 2X + dx - pad), zero outside the image and for c = 3 (conv_stem_padded_dims)
__global__ void __launch_bounds__(256) k_stem_s2d(const unsigned short* __restrict__ x, uint4* __restrict__ y, int n, int h, int w,
                                                  int hp, int wp, int pad) {
  const uint32_t total = uint32_t(n) * uint32_t(hp) * uint32_t(wp) * 2u;      // one thread per (pixel, dy): 16 bytes; < 2^31 (launcher)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dy = int(i & 1u);
    const uint32_t pix = i >> 1;                                // 32-bit divisions: the 64-bit ones cost more than the copy itself
    const uint32_t row = pix / uint32_t(wp);
    const int X = int(pix - row * uint32_t(wp));
    const long long img = row / uint32_t(hp);
    const int Y = int(row - uint32_t(img) * uint32_t(hp));
    const int yy = 2 * Y + dy - pad, x0 = 2 * X - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < h) {
      const unsigned short* p = x + 3 * ((img * h + yy) * (long long)w + x0);
      if (x0 >= 0 && x0 < w) { v.x = uint32_t(p[0]) | (uint32_t(p[1]) << 16); v.y = uint32_t(p[2]); }
      if (x0 + 1 >= 0 && x0 + 1 < w) { v.z = uint32_t(p[3]) | (uint32_t(p[4]) << 16); v.w = uint32_t(p[5]); }
    }
    y[i] = v;
  }
}

// bf16 NHWC3 -> NHWC4 (zero 4th channel): lets the stem gather 8-byte pixels, two filter taps per 16-byte chunk
__global__ void __launch_bounds__(256) k_pad_c3_c4(const unsigned short* __restrict__ x, uint2* __restrict__ y, long long npix) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const unsigned short* p = x + 3 * i;
    y[i] = make_uint2(uint32_t(p[0]) | (uint32_t(p[1]) << 16), uint32_t(p[2]));
  }
}

static inline int grid_for(long long work, int threads, int num_sms) {
  long long b = (work + threads - 1) / threads, cap = (long long)num_sms * 16;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

// geometry walk shared by reserve() and forward(): calls f(bytes) for every activation tensor produced
template <class F>
static void walk_activations(const Plan& pl, int n, int T, F&& f) {
  const ConvLayer& stem = pl.convs[0];
  int h = conv_out_dim(pl.in_h, stem.r, stem.stride, stem.pad), w = conv_out_dim(pl.in_w, stem.s, stem.stride, stem.pad);
  f((size_t)n * h * w * stem.cout * 2);
  {
    int hp = pl.in_h, wp = pl.in_w;
    f(conv_stem_padded_dims(stem, pl.in_h, pl.in_w, &hp, &wp) ? (size_t)n * hp * wp * 32 : (size_t)n * pl.in_h * pl.in_w * 8);
  }
  h = conv_out_dim(h, 3, 2, 1); w = conv_out_dim(w, 3, 2, 1);
  f((size_t)n * h * w * stem.cout * 2);
  long long P = n;
  for (size_t b = 0; b < pl.blocks.size(); ++b) {
    const BlockDesc& bd = pl.blocks[b];
    int hh = h, ww = w;
    for (int k = 0; k < bd.n_convs; ++k) {
      const ConvLayer& L = pl.convs[bd.conv0 + k];
      hh = conv_out_dim(hh, L.r, L.stride, L.pad); ww = conv_out_dim(ww, L.s, L.stride, L.pad);
      const bool last = k == bd.n_convs - 1;
      const long long Pout = (last && b == 0 && T > 1) ? P * T : P;
      f((size_t)Pout * hh * ww * (L.cout / (L.fold ? L.fold : 1)) * 2);
    }
    if (bd.ds >= 0) f((size_t)P * hh * ww * pl.convs[bd.ds].cout * 2);
    if (b == 0 && T > 1) P *= T;
    h = hh; w = ww;
  }
}

}  // namespace fav

using namespace fav;

extern "C" int fav_load_weights(fav_handle h, const void* blob, size_t nbytes, int model_id, int num_classes, int in_h,
                                int in_w) {
  FAV_REQUIRE(h && blob, "fav_load_weights: null pointer");
  FAV_DEVICE(h);
  FAV_REQUIRE(in_h > 0 && in_w > 0 && num_classes >= 2, "fav_load_weights: bad input size / classes");
  const uint8_t* p = reinterpret_cast<const uint8_t*>(blob);
  const uint8_t* end = p + nbytes;
  FAV_REQUIRE(nbytes >= 32 && memcmp(p, "FAVW1\0\0\0", 8) == 0, "fav_load_weights: bad magic (expected FAVW1)");
  const int32_t* hdr = reinterpret_cast<const int32_t*>(p + 8);
  const int n_convs = hdr[0], blob_classes = hdr[1], blob_model = hdr[2], n_blocks = hdr[3];
  FAV_REQUIRE(blob_model == model_id && blob_classes == num_classes, "fav_load_weights: blob is model %d / %d classes, asked %d / %d",
              blob_model, blob_classes, model_id, num_classes);
  FAV_REQUIRE(n_convs >= 3 && n_convs < 256 && n_blocks >= 1 && n_blocks < 64, "fav_load_weights: implausible header");
  p += 8 + 16;
  FAV_REQUIRE(p + 8 * (size_t)n_blocks <= end, "fav_load_weights: truncated block table");
  Plan* pl = new Plan();
  pl->model_id = model_id; pl->num_classes = num_classes; pl->in_h = in_h; pl->in_w = in_w;
  const int32_t* bt = reinterpret_cast<const int32_t*>(p);
  int ci = 1;
  for (int b = 0; b < n_blocks; ++b) {
    BlockDesc bd;
    bd.n_convs = bt[2 * b]; bd.conv0 = ci; ci += bd.n_convs;
    bd.ds = bt[2 * b + 1] ? ci++ : -1;
    pl->blocks.push_back(bd);
  }
  if (ci + 1 != n_convs) { delete pl; set_error("fav_load_weights: block table (%d convs) disagrees with header (%d)", ci + 1, n_convs); return FAV_E_ARG; }
  p += 8 * (size_t)n_blocks;
  p = reinterpret_cast<const uint8_t*>(blob) + (((p - reinterpret_cast<const uint8_t*>(blob)) + 15) / 16) * 16;

  // pass 1: sizes
  struct Rec { const uint8_t* w; const uint8_t* b; };
  std::vector<Rec> recs;
  size_t arena = 0;
  const uint8_t* q = p;
  for (int i = 0; i < n_convs; ++i) {
    if (q + 32 > end) { delete pl; set_error("fav_load_weights: truncated at conv %d", i); return FAV_E_ARG; }
    const int32_t* r = reinterpret_cast<const int32_t*>(q);
    ConvLayer L;
    L.cout = r[0]; L.cin = r[1]; L.r = r[2]; L.s = r[3]; L.stride = r[4]; L.pad = r[5];
    if (L.cout <= 0 || L.cin <= 0 || L.r <= 0 || L.s <= 0 || L.stride <= 0 || L.pad < 0) {
      delete pl; set_error("fav_load_weights: bad conv record %d", i); return FAV_E_ARG;
    }
    const int k_blob = L.r * L.s * L.cin;
    if (i == 0 && L.cin == 3) {
      // stem: stride 2 with R, S <= 8 -> 4x4/s1 conv over the space-to-depth input, A by TMA (conv a_mode 5); anything else
      // -> channels padded to 4 and S rounded up to even for the gather path (a_mode 3)
      static const int env_stem_tma = [] { const char* e = getenv("FAV_STEM_TMA"); return e ? atoi(e) : 1; }();   // 0: gather path
      if (env_stem_tma && L.stride == 2 && L.r <= 8 && L.s <= 8) { L.s2d = 1; L.cin_store = 16; L.s_store = 4; }
      else { L.cin_store = 4; L.s_store = (L.s + 1) & ~1; }
    }
    conv_layer_finalize(L);
    const size_t wb = ((size_t)L.cout * k_blob * 2 + 15) / 16 * 16, bb = ((size_t)L.cout * 4 + 15) / 16 * 16;
    if (q + 32 + wb + bb > end) { delete pl; set_error("fav_load_weights: truncated weights at conv %d", i); return FAV_E_ARG; }
    recs.push_back({q + 32, q + 32 + wb});
    q += 32 + wb + bb;
    pl->convs.push_back(L);
  }
  // fuse every TMA-able 1x1 downsample conv into the last conv of its block: its weights become extra K columns.
  // (stride-2 branches need even input dims for the parity-split tensor map, so the geometry is walked here)
  {
    const ConvLayer& stem0 = pl->convs[0];
    int gh = conv_out_dim(in_h, stem0.r, stem0.stride, stem0.pad), gw = conv_out_dim(in_w, stem0.s, stem0.stride, stem0.pad);
    gh = conv_out_dim(gh, 3, 2, 1); gw = conv_out_dim(gw, 3, 2, 1);
    for (BlockDesc& bd : pl->blocks) {
      const int bh = gh, bw = gw;               // block input size = input of the downsample branch
      for (int k = 0; k < bd.n_convs; ++k) {
        const ConvLayer& L = pl->convs[bd.conv0 + k];
        gh = conv_out_dim(gh, L.r, L.stride, L.pad); gw = conv_out_dim(gw, L.s, L.stride, L.pad);
      }
      if (bd.ds < 0) continue;
      const ConvLayer& D = pl->convs[bd.ds];
      ConvLayer& last = pl->convs[bd.conv0 + bd.n_convs - 1];
      const bool dims_ok = D.stride == 1 || (D.stride == 2 && (bh % 2) == 0 && (bw % 2) == 0);
      const bool last_tma = (last.cin % 64) == 0 && last.stride == 1 && last.r == last.s && 2 * last.pad == last.r - 1 && gw <= 128;
      if (D.r == 1 && D.s == 1 && D.pad == 0 && (D.cin % 64) == 0 && D.cout == last.cout && dims_ok && last_tma) {
        last.k2pad = D.kpad; last.cin2 = D.cin; last.stride2 = D.stride;
        bd.fused_ds = true;
      }
    }
  }
  // 3x3 / stride 1 / pad 1 convs whose input is 2x2: every output pixel sees every input pixel through exactly one tap,
  // so the conv is a dense GEMM [P, 4*Cin] x [4*Cin, 4*Cout] on the NHWC image itself (no im2col redundancy, no padding MACs)
  std::vector<int> fold_src(n_convs, 0);
  {
    const ConvLayer& stem0 = pl->convs[0];
    int gh = conv_out_dim(in_h, stem0.r, stem0.stride, stem0.pad), gw = conv_out_dim(in_w, stem0.s, stem0.stride, stem0.pad);
    gh = conv_out_dim(gh, 3, 2, 1); gw = conv_out_dim(gw, 3, 2, 1);
    for (BlockDesc& bd : pl->blocks) {
      for (int k = 0; k < bd.n_convs; ++k) {
        ConvLayer& L = pl->convs[bd.conv0 + k];
        if (gh == 2 && gw == 2 && L.r == 3 && L.s == 3 && L.stride == 1 && L.pad == 1 && L.k2pad == 0 && (L.cin % 16) == 0 &&
            (L.cout % 16) == 0) {
          fold_src[bd.conv0 + k] = 1;
          L.fold = 4; L.cin *= 4; L.cout *= 4; L.r = 1; L.s = 1; L.pad = 0;
          conv_layer_finalize(L);
        } else {
          gh = conv_out_dim(gh, L.r, L.stride, L.pad); gw = conv_out_dim(gw, L.s, L.stride, L.pad);
        }
      }
    }
  }
  for (const ConvLayer& L : pl->convs)
    arena += ((size_t)L.cout_pad * (L.kpad + L.k2pad) * 2 + 255) / 256 * 256 + ((size_t)L.cout_pad * 4 + 255) / 256 * 256;
  FAV_CUDA_OK(cudaSetDevice(h->device));
  cudaError_t e = cudaMalloc(&pl->arena, arena);
  if (e != cudaSuccess) { delete pl; set_error("fav_load_weights: cudaMalloc(%zu) failed: %s", arena, cudaGetErrorString(e)); return FAV_E_CUDA; }
  pl->arena_bytes = arena;
  cudaMemset(pl->arena, 0, arena);
  // pass 2: upload with row padding k -> kpad
  uint8_t* d = reinterpret_cast<uint8_t*>(pl->arena);
  for (int i = 0; i < n_convs; ++i) {
    ConvLayer& L = pl->convs[i];
    if (fold_src[i]) {
      // W2[(po*cout0 + co)][(pi*cin0 + ci)] = W[co][yi - yo + 1][xi - xo + 1][ci]   (po = yo*2 + xo, pi = yi*2 + xi)
      const int cin0 = L.cin / 4, cout0 = L.cout / 4;
      std::vector<uint16_t> tmp((size_t)L.cout * L.k);
      const uint16_t* src = reinterpret_cast<const uint16_t*>(recs[i].w);
      for (int po = 0; po < 4; ++po)
        for (int co = 0; co < cout0; ++co)
          for (int pi = 0; pi < 4; ++pi) {
            const int rr = (pi >> 1) - (po >> 1) + 1, ss = (pi & 1) - (po & 1) + 1;
            memcpy(&tmp[((size_t)po * cout0 + co) * L.k + (size_t)pi * cin0], &src[(((size_t)co * 3 + rr) * 3 + ss) * cin0], (size_t)cin0 * 2);
          }
      e = cudaMemcpy2D(d, (size_t)L.kpad * 2, tmp.data(), (size_t)L.k * 2, (size_t)L.k * 2, L.cout, cudaMemcpyHostToDevice);
    } else if (L.s2d) {
      // [cout][r][s][3] -> [cout][r' = r/2][s' = s/2][dy = r%2][dx = s%2][4], zero filled
      std::vector<uint16_t> tmp((size_t)L.cout * L.k, 0);
      const uint16_t* src = reinterpret_cast<const uint16_t*>(recs[i].w);
      for (int co = 0; co < L.cout; ++co)
        for (int rr = 0; rr < L.r; ++rr)
          for (int ss = 0; ss < L.s; ++ss)
            for (int c = 0; c < L.cin; ++c)
              tmp[(size_t)co * L.k + ((rr >> 1) * 4 + (ss >> 1)) * 16 + ((rr & 1) * 2 + (ss & 1)) * 4 + c] =
                  src[(((size_t)co * L.r + rr) * L.s + ss) * L.cin + c];
      e = cudaMemcpy2D(d, (size_t)L.kpad * 2, tmp.data(), (size_t)L.k * 2, (size_t)L.k * 2, L.cout, cudaMemcpyHostToDevice);
    } else if (L.cin_store) {
      // [cout][r][s][3] -> [cout][r][s_store][4], zero filled
      std::vector<uint16_t> tmp((size_t)L.cout * L.k, 0);
      const uint16_t* src = reinterpret_cast<const uint16_t*>(recs[i].w);
      for (int co = 0; co < L.cout; ++co)
        for (int rr = 0; rr < L.r; ++rr)
          for (int ss = 0; ss < L.s; ++ss)
            for (int c = 0; c < L.cin; ++c)
              tmp[((size_t)co * L.r + rr) * L.s_store * 4 + ss * 4 + c] = src[(((size_t)co * L.r + rr) * L.s + ss) * L.cin + c];
      e = cudaMemcpy2D(d, (size_t)L.kpad * 2, tmp.data(), (size_t)L.k * 2, (size_t)L.k * 2, L.cout, cudaMemcpyHostToDevice);
    } else {
      e = cudaMemcpy2D(d, (size_t)(L.kpad + L.k2pad) * 2, recs[i].w, (size_t)L.k * 2, (size_t)L.k * 2, L.cout, cudaMemcpyHostToDevice);
    }
    L.w = reinterpret_cast<const __nv_bfloat16*>(d);
    d += ((size_t)L.cout_pad * (L.kpad + L.k2pad) * 2 + 255) / 256 * 256;
    if (e == cudaSuccess) {
      if (fold_src[i]) {
        const int cout0 = L.cout / 4;
        for (int po = 0; po < 4 && e == cudaSuccess; ++po)
          e = cudaMemcpy(d + (size_t)po * cout0 * 4, recs[i].b, (size_t)cout0 * 4, cudaMemcpyHostToDevice);
      } else {
        e = cudaMemcpy(d, recs[i].b, (size_t)L.cout * 4, cudaMemcpyHostToDevice);
      }
    }
    L.bias = reinterpret_cast<const float*>(d);
    d += ((size_t)L.cout_pad * 4 + 255) / 256 * 256;
    if (e != cudaSuccess) {
      set_error("fav_load_weights: upload failed: %s", cudaGetErrorString(e));
      plan_destroy(pl);
      return FAV_E_CUDA;
    }
  }
  for (const BlockDesc& bd : pl->blocks) {
    if (!bd.fused_ds) continue;
    const ConvLayer& D = pl->convs[bd.ds];
    ConvLayer& last = pl->convs[bd.conv0 + bd.n_convs - 1];
    // extra K columns of the last conv <- downsample weights (device to device); bias <- bias + downsample bias
    e = cudaMemcpy2D(const_cast<__nv_bfloat16*>(last.w) + last.kpad, (size_t)(last.kpad + last.k2pad) * 2, D.w, (size_t)D.kpad * 2,
                     (size_t)D.kpad * 2, D.cout, cudaMemcpyDeviceToDevice);
    std::vector<float> b0(last.cout), b1(D.cout);
    if (e == cudaSuccess) e = cudaMemcpy(b0.data(), last.bias, (size_t)last.cout * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(b1.data(), D.bias, (size_t)D.cout * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < last.cout; ++i) b0[i] += b1[i];
    if (e == cudaSuccess) e = cudaMemcpy(const_cast<float*>(last.bias), b0.data(), (size_t)last.cout * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("fav_load_weights: downsample fusion failed: %s", cudaGetErrorString(e)); plan_destroy(pl); return FAV_E_CUDA; }
  }
  for (ConvLayer& L : pl->convs) {
    int rc = conv_layer_finalize(L);
    if (rc) { plan_destroy(pl); return rc; }
  }
  const ConvLayer& fc = pl->convs.back();
  if (fc.r != 1 || fc.s != 1 || fc.cout != num_classes) { plan_destroy(pl); set_error("fav_load_weights: last record must be the 1x1 fc"); return FAV_E_ARG; }
  if (h->plan) plan_destroy(h->plan);
  h->plan = pl;
  if (h->ws) { cudaFree(h->ws); h->ws = nullptr; h->ws_bytes = 0; }
  return FAV_OK;
}

extern "C" int fav_reserve(fav_handle h, int max_images, int T) {
  FAV_REQUIRE(h && h->plan, "fav_reserve: load weights first");
  FAV_DEVICE(h);
  FAV_REQUIRE(max_images > 0 && T >= 1, "fav_reserve: bad max_images/T");
  Plan& pl = *h->plan;
  size_t mx = 0;
  walk_activations(pl, max_images, T, [&](size_t b) { if (b > mx) mx = b; });
  mx = (mx + 1023) / 1024 * 1024;
  const size_t need = 5 * mx;
  if (need > h->ws_bytes) {
    if (h->ws) { FAV_CUDA_OK(cudaFree(h->ws)); h->ws = nullptr; h->ws_bytes = 0; }
    FAV_CUDA_OK(cudaMalloc(&h->ws, need));
    h->ws_bytes = need;
  }
  pl.buf_bytes = h->ws_bytes / 5 / 1024 * 1024;
  pl.max_images = max_images; pl.max_T = T;
  return FAV_OK;
}

extern "C" int fav_forward_mc(fav_handle h, const void* d_x, float* d_logits, int n, int T, float p_drop, uint64_t seed,
                              uint64_t first_image, void* stream) {
  FAV_REQUIRE(h && h->plan, "fav_forward_mc: load weights first");
  FAV_DEVICE(h);
  if (!h->plan) return FAV_E_STATE;
  FAV_REQUIRE(n >= 0 && T >= 1 && (n == 0 || (d_x && d_logits)), "fav_forward_mc: bad arguments");
  FAV_REQUIRE(T == 1 || (p_drop >= 0.f && p_drop < 1.f), "fav_forward_mc: p_drop must be in [0,1)");
  if (n == 0) return FAV_OK;
  Plan& pl = *h->plan;
  size_t mx = 0;
  walk_activations(pl, n, T, [&](size_t b) { if (b > mx) mx = b; });
  if (!h->ws || mx > pl.buf_bytes) {
    int rc = fav_reserve(h, n > pl.max_images ? n : pl.max_images, T > pl.max_T ? T : pl.max_T);
    if (rc) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* base = reinterpret_cast<uint8_t*>(h->ws);
  void* X[2] = {base, base + pl.buf_bytes};
  void* Y1 = base + 2 * pl.buf_bytes;
  void* Y2 = base + 3 * pl.buf_bytes;
  void* DS = base + 4 * pl.buf_bytes;
  const bool mc = T > 1;
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);

  auto run = [&](const ConvLayer& L, const void* x, void* y, const void* res, int P, int hh, int ww, int relu, int drop,
                 int rep, int layer_id, int out_f32, const void* x2 = nullptr, int h2 = 0, int w2 = 0, int a_mode = -1,
                 int drop2_layer = -1) -> int {
    ConvCall c;
    c.L = &L; c.x = x; c.y = y; c.res = res; c.p = P; c.h = hh; c.w = ww; c.relu = relu; c.out_f32 = out_f32;
    c.x2 = x2; c.h2 = h2; c.w2 = w2; c.a_mode = a_mode; c.drop2_layer = drop2_layer;
    c.T = T; c.rep = rep; c.drop = drop; c.p_drop = p_drop; c.seed = seed; c.first_image = first_image; c.layer_id = layer_id;
    return conv_launch(h, c, st);
  };

  // stem + max-pool (pass-invariant)
  const ConvLayer& stem = pl.convs[0];
  const void* stem_in = d_x;
  int stem_mode = -1, hp = 0, wp = 0;
  if (conv_stem_padded_dims(stem, pl.in_h, pl.in_w, &hp, &wp)) {
    const long long npix = (long long)n * hp * wp;
    FAV_REQUIRE((size_t)npix * 32 <= pl.buf_bytes, "workspace too small for the space-to-depth input");
    FAV_REQUIRE(npix < (1ll << 30), "space-to-depth input: too many pixels in one call (%lld)", npix);
    k_stem_s2d<<<grid_for(npix * 2, 256, h->num_sms), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(d_x),
                                                                    reinterpret_cast<uint4*>(DS), n, pl.in_h, pl.in_w, hp, wp, stem.pad);
    FAV_CUDA_OK(cudaGetLastError());
    h->launches++;
    stem_in = DS;
    stem_mode = 5;
  } else if (stem.cin_store == 4) {
    const long long npix = (long long)n * pl.in_h * pl.in_w;
    FAV_REQUIRE((size_t)npix * 8 <= pl.buf_bytes, "workspace too small for the padded input");
    k_pad_c3_c4<<<grid_for(npix, 256, h->num_sms), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(d_x),
                                                                 reinterpret_cast<uint2*>(DS), npix);
    h->launches++;
    stem_in = DS;
  }
  int rc = run(stem, stem_in, Y1, nullptr, n, pl.in_h, pl.in_w, 1, 0, 1, 0, 0, nullptr, 0, 0, stem_mode);
  if (rc) return rc;
  int hh = conv_out_dim(pl.in_h, stem.r, stem.stride, stem.pad), ww = conv_out_dim(pl.in_w, stem.s, stem.stride, stem.pad);
  {
    const int oh = conv_out_dim(hh, 3, 2, 1), ow = conv_out_dim(ww, 3, 2, 1);
    FAV_REQUIRE((stem.cout & 7) == 0, "stem Cout must be a multiple of 8");
    const long long work = (long long)n * oh * ow * (stem.cout / 8);
    FAV_REQUIRE(work < (1ll << 31), "max-pool: too many outputs in one call (%lld)", work);
    k_maxpool3x3s2<<<grid_for(work, 256, h->num_sms), 256, 0, st>>>(reinterpret_cast<const uint4*>(Y1), reinterpret_cast<uint4*>(X[0]),
                                                                    n, hh, ww, stem.cout / 8, oh, ow);
    h->launches++;
    hh = oh; ww = ow;
  }
  int cur = 0, P = n, ch = stem.cout;
  bool fused_fc_drop = false;
  for (size_t b = 0; b < pl.blocks.size(); ++b) {
    const BlockDesc& bd = pl.blocks[b];
    const void* ident = X[cur];
    int oh = hh, ow = ww;
    if (bd.ds >= 0 && !bd.fused_ds) {
      const ConvLayer& D = pl.convs[bd.ds];
      rc = run(D, X[cur], DS, nullptr, P, hh, ww, 0, 0, 1, 0, 0);
      if (rc) return rc;
      ident = DS;
    }
    const void* in = X[cur];
    void* tmp[2] = {Y1, Y2};
    int ih = hh, iw = ww;
    for (int k = 0; k < bd.n_convs; ++k) {
      const ConvLayer& L = pl.convs[bd.conv0 + k];
      const bool last = k == bd.n_convs - 1;
      oh = conv_out_dim(ih, L.r, L.stride, L.pad); ow = conv_out_dim(iw, L.s, L.stride, L.pad);
      const int ch_ = L.fold ? 1 : ih, cw_ = L.fold ? 1 : iw;      // folded: the whole 2x2 image is one GEMM row
      if (!last) {
        rc = run(L, in, tmp[k & 1], nullptr, P, ch_, cw_, 1, 0, 1, 0, 0);
        in = tmp[k & 1];
      } else {
        const int rep = (mc && b == 0) ? T : 1;
        // 1x1 final feature map: the pooled feature is this output, so the dropout before fc is a second mask in this epilogue
        fused_fc_drop = mc && rep == 1 && b + 1 == pl.blocks.size() && oh * ow == 1 && !L.fold;
        const int d2 = fused_fc_drop ? 255 : -1;
        if (bd.fused_ds) rc = run(L, in, X[cur ^ 1], nullptr, P, ih, iw, 1, mc ? 1 : 0, rep, int(b), 0, X[cur], hh, ww, -1, d2);
        else rc = run(L, in, X[cur ^ 1], ident, P, ch_, cw_, 1, mc ? 1 : 0, rep, int(b), 0, nullptr, 0, 0, -1, d2);
        ch = L.cout / (L.fold ? L.fold : 1);
      }
      if (rc) return rc;
      ih = oh; iw = ow;
    }
    if (mc && b == 0) P *= T;
    cur ^= 1; hh = oh; ww = ow;
  }
  // global average pool + dropout on the pooled feature, then fc as a 1x1 conv with fp32 output [n, T, C]
  FAV_REQUIRE((ch & 7) == 0, "feature width must be a multiple of 8");
  const void* fc_in = Y1;
  if (hh * ww == 1 && (fused_fc_drop || !mc)) {
    fc_in = X[cur];                               // nothing to pool; the mask (if any) was applied by the last conv's epilogue
  } else {
    const long long work = (long long)P * (ch / 8);
    FAV_REQUIRE(!mc || (ch & 15) == 0, "MC-dropout on the pooled feature needs a width that is a multiple of 16");
    const uint32_t thr = mc ? dropout_thr8(p_drop) : 0u;
    const int groups = (ch / 8 + 31) / 32;
    if (hh * ww >= 64 && (long long)P * groups <= (1 << 20))
      k_pool_dropout_wide<<<P * groups, 256, 0, st>>>(
          reinterpret_cast<const uint4*>(X[cur]), reinterpret_cast<uint4*>(Y1), P, hh * ww, ch / 8, T, mc ? 1 : 0, thr,
          dropout_scale8(thr), k0, k1, uint32_t(first_image), stream_id(KIND_DROPOUT, 255, 0));
    else
      k_pool_dropout<<<grid_for(work, 256, h->num_sms), 256, 0, st>>>(
          reinterpret_cast<const uint4*>(X[cur]), reinterpret_cast<uint4*>(Y1), P, hh * ww, ch / 8, T, mc ? 1 : 0, thr,
          dropout_scale8(thr), k0, k1, uint32_t(first_image), stream_id(KIND_DROPOUT, 255, 0));
    h->launches++;
  }
  const ConvLayer& fc = pl.convs.back();
  FAV_REQUIRE(fc.cin == ch, "fc input width %d does not match the trunk (%d)", fc.cin, ch);
  rc = run(fc, fc_in, d_logits, nullptr, P, 1, 1, 0, 0, 1, 0, 1);
  if (rc) return rc;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}
