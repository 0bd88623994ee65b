// conv_dev.cuh -- device-side pieces shared by the conv kernels (conv.cu, conv_pair.cu): kernel arguments, tile
// decoding, the division-free k-block walk and the fused epilogue.
#pragma once
#include "conv.cuh"
#include "tc_ptx.cuh"

namespace fav {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;          // 16 KiB

// unsigned division by a launch-time constant: q = umulhi(n, floor(2^32 / d) + 1), exact while n * d < 2^32 (checked on the
// host: ConvArgs::fastdiv); one IMAD.HI instead of the ~20-instruction division sequence in every role's tile / row decode
struct FastDiv {
  uint32_t m, d;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.m = d > 1 ? uint32_t((1ull << 32) / d) + 1u : 0u;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f, bool fast) {
  if (!fast) return n / f.d;
  return f.d == 1 ? n : __umulhi(n, f.m);
}

struct ConvArgs {
  const __nv_bfloat16* x;
  void* y;
  const float* bias;
  const __nv_bfloat16* res;
  int P, H, W, Cin, OH, OW, Cout;
  int R, S, stride, pad;
  int K, num_kb, M, BN, stages;
  int relu, out_f32, a_mode;
  int bw, bh, bn_img, tiles_w, tiles_h, cin_blocks;
  int ntiles, total_tiles;     // N tiles per M tile, all CTA tiles (persistent scheduler)
  int mt_per_tile, mtiles;     // 128-row M tiles per CTA tile (1 or 2), number of 128-row M tiles
  int s_store;                 // a_mode 3: filter-row slots (S padded to an even count), Cin stored as 4
  int stem_tma;                // a_mode 0 on the pre-padded NHWC4 stem input: one 5-D TMA box = one filter row x 16 taps x 4 ch per k-block
  int T, rep, drop;
  uint32_t drop_thr8;          // keep a channel iff its Philox byte >= drop_thr8 (= round(p * 256), common.cuh)
  uint32_t drop_add4, drop_hi4;   // operands of dropout_keep4 (common.cuh)
  PhiloxKeys drop_keys;        // round keys of (k0, k1)
  // flattened 1x1 convolution (conv_launch): the launch sees ONE image of 1 x (P*OH*OW) pixels, so tiles are 128 consecutive
  // pixels whatever the image size; the dropout counters still need (image, pixel in image) = divmod(pixel, flat_ohw)
  int flat_ohw;                // 0 = not flattened
  unsigned long long flat_m64; // ceil(2^64 / flat_ohw): q = umul64hi(n, m) is exact for every 32-bit n
  float drop_scale;
  uint32_t k0, k1, first_image, drop_stream;
  uint32_t drop_stream2;       // second mask stream (ConvCall::drop2_layer), used when drop2 != 0
  int drop2;
  uint32_t tmem_cols, idesc;
  int pair_tiles, nkb_tot;     // 2-SM variant: CTA-pair tiles (two 128-row M tiles each), resident W k-blocks (main + fused branch)
  int kb2, stride2, Cin2;      // fused second source (the block's 1x1 downsample branch): extra k-blocks after the main taps
  int stg_bytes;               // > 0: epilogue stages bf16 output in shared memory (two 128 x 64 slabs) and writes it with TMA stores
  int res_mma;                 // the residual is added by the tensor core: extra k-blocks A = residual tile (tmR), B = a 64x64 identity
                               // resident in shared memory, N = 64 MMAs into the matching accumulator columns (a.res is null then)
  uint32_t idesc64;            // instruction descriptor of those N = 64 MMAs
  int fastdiv;                 // 1: every (dividend, divisor) pair of the decode functions satisfies n * d < 2^32
  FastDiv fd_ntiles, fd_tw, fd_th, fd_twth, fd_perimg, fd_bw, fd_ksplit, fd_ohw, fd_ow;
  int ksplit;                  // > 1: each output tile is computed by ksplit CTA tiles taking interleaved k-blocks; every slice stores
  float* acc32;                //      its fp32 partial tile in acc32 [tile][slice][128][BN]; the LAST slice to finish (tickets[tile]) adds the
  int* tickets;                //      partials in slice order (deterministic) and applies bias / residual / ReLU
  int ablate;                  // tuning aid (env FAV_CONV_ABLATE): 1 skip A loads, 2 skip B loads, 4 skip MMAs, 8 skip epilogue math
  unsigned long long* stats;   // optional per-launch role timing (8 counters), see fav_conv_stats_read
};

constexpr int EPI_WARP0 = 2;
constexpr int IDENT_BYTES = 64 * 128;              // resident identity tile of the residual MMAs
constexpr int STG_SLAB_BYTES = 128 * 128;         // staged epilogue: 128 pixels x 64 bf16 channels
constexpr int THREADS_TMA1 = 32 * (2 + 8);        // MT = 1: 8 epilogue warps, two CTAs per SM
constexpr int THREADS_TMA2 = 32 * (2 + 16);       // MT = 2: 16 epilogue warps, one CTA per SM
constexpr int THREADS_GATHER = 32 * (2 + 8 + 8);  // gather variant: 8 epilogue + 8 gather warps (two groups on alternate k-blocks)

struct Tile { int mt, nt, q0, oh0, ow0, ks, tb; };   // mt = index of the 128-row M tile, ks = split-K slice, tb = output tile index

__device__ __forceinline__ Tile decode_tile(const ConvArgs& a, int tile_in, int u = 0) {
  Tile t;
  const bool fast = a.fastdiv != 0;
  uint32_t tile = uint32_t(tile_in);
  t.ks = 0;
  if (a.ksplit > 1) {                                              // slices of one output tile run side by side
    const uint32_t q = fdiv(tile, a.fd_ksplit, fast);
    t.ks = int(tile - q * a.fd_ksplit.d);
    tile = q;
  }
  t.tb = int(tile);
  const uint32_t tq = fdiv(tile, a.fd_ntiles, fast);
  t.nt = int(tile - tq * a.fd_ntiles.d);
  t.mt = int(tq) * a.mt_per_tile + u;
  t.q0 = 0; t.oh0 = 0; t.ow0 = 0;
  if (a.a_mode == 0) {
    const uint32_t mt = uint32_t(t.mt);
    const uint32_t r1 = fdiv(mt, a.fd_tw, fast), tw = mt - r1 * a.fd_tw.d;            // mt = (tn * tiles_h + th) * tiles_w + tw
    const uint32_t tn = fdiv(r1, a.fd_th, fast), th = r1 - tn * a.fd_th.d;
    t.q0 = int(tn) * a.bn_img; t.oh0 = int(th) * a.bh; t.ow0 = int(tw) * a.bw;
  }
  return t;
}

// Split-K interleave: slice ks takes the visited k-blocks number ks, ks + ksplit, ...  `cd` is a countdown (no modulo in the
// single issuing thread): take the k-block when it reaches 0.
__device__ __forceinline__ bool kb_mine(const ConvArgs& a, int& cd) {
  if (a.ksplit <= 1) return true;
  if (cd == 0) { cd = a.ksplit - 1; return true; }
  --cd;
  return false;
}

// walks the k-blocks of tile t (all-padding filter taps are skipped); returns the split-K countdown after the last one
template <class F>
__device__ __forceinline__ int for_each_kb(const ConvArgs& a, const Tile& t, F&& f) {
  int cd = t.ks;
  if (a.a_mode != 0 || a.stem_tma) {
    for (int kb = 0; kb < a.num_kb; ++kb)
      if (kb_mine(a, cd)) f(kb, 0, 0, 0);
    return cd;
  }
  const int oh_last = min(t.oh0 + a.bh, a.OH) - 1, ow_last = min(t.ow0 + a.bw, a.OW) - 1;
  int kb = 0;
  for (int r = 0; r < a.R; ++r) {
    const bool row_ok = !(oh_last * a.stride + r - a.pad < 0 || t.oh0 * a.stride + r - a.pad >= a.H);
    for (int ss = 0; ss < a.S; ++ss, kb += a.cin_blocks) {
      if (!row_ok || ow_last * a.stride + ss - a.pad < 0 || t.ow0 * a.stride + ss - a.pad >= a.W) continue;
      for (int cb = 0; cb < a.cin_blocks; ++cb)
        if (kb_mine(a, cd)) f(kb + cb, r, ss, cb);
    }
  }
  return cd;
}

// output pixel owned by A-tile row `row` of tile t
__device__ __forceinline__ bool decode_row(const ConvArgs& a, const Tile& t, int row, int& q, int& oh, int& ow) {
  const bool fast = a.fastdiv != 0;
  if (a.a_mode == 0) {
    const uint32_t nl = fdiv(uint32_t(row), a.fd_perimg, fast), rem = uint32_t(row) - nl * a.fd_perimg.d;
    const uint32_t hl = fdiv(rem, a.fd_bw, fast), wl = rem - hl * a.fd_bw.d;
    q = t.q0 + int(nl); oh = t.oh0 + int(hl); ow = t.ow0 + int(wl);
    return int(nl) < a.bn_img && q < a.P && oh < a.OH && ow < a.OW;
  }
  const long long m = (long long)t.mt * BM + row;
  q = 0; oh = 0; ow = 0;
  if (m >= a.M) return false;
  const uint32_t qq = fdiv(uint32_t(m), a.fd_ohw, fast), rem = uint32_t(m) - qq * a.fd_ohw.d;
  const uint32_t hh = fdiv(rem, a.fd_ow, fast);
  q = int(qq); oh = int(hh); ow = int(rem - hh * a.fd_ow.d);
  return true;
}

// MC-dropout on 16 packed bf16 channels starting at channel offset e16*16 of the image: ONE Philox call gives sixteen byte
// lanes (contract: common.cuh); dropout_keep4 compares four bytes at a time, PRMT sign replication turns the flags into
// 0xFFFF-per-kept-channel masks that are ANDed onto the packed pairs (dropped channels become +0.0)
__device__ __forceinline__ void dropout_and16(const ConvArgs& a, uint32_t e16, uint32_t image, uint32_t tt, const uint32_t (&pk)[8],
                                              uint32_t (&o)[8], uint32_t stream) {
  const uint4 r = philox4x32_10_keys(e16, image, tt, stream, a.drop_keys);
  const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t k4 = dropout_keep4(rw[i], a.drop_add4, a.drop_hi4);
    o[2 * i] = pk[2 * i] & dropout_pair_lo(k4);
    o[2 * i + 1] = pk[2 * i + 1] & dropout_pair_hi(k4);
  }
}

// (image, pixel in image) of output row (q, hw) for the dropout counters: identity unless the launch is a flattened 1x1 conv
__device__ __forceinline__ void dropout_pixel(const ConvArgs& a, int q, int hw, int& q_img, int& hw_img) {
  q_img = q; hw_img = hw;
  if (a.flat_ohw) {
    q_img = int(__umul64hi((unsigned long long)uint32_t(hw), a.flat_m64));
    hw_img = hw - q_img * a.flat_ohw;
  }
}

// Epilogue of one 128-row sub-tile for one warp: TMEM (lanes of this warp's quarter, columns of accumulator `trow`) ->
// bias + residual + ReLU + MC-dropout mask (+ T masked replicas) -> bf16 NHWC / fp32.  The warp handles the 16-column
// chunks j = sub_w, sub_w + wpq, ...; residual loads and dropout masks are issued before the TMEM load they combine with.
__device__ __forceinline__ void conv_epilogue_subtile(const ConvArgs& a, const Tile& t, uint32_t trow, int row, int sub_w, int WPQ) {
  const int ohw = a.OH * a.OW, n_rep = a.rep > 1 ? a.rep : 1;
  int q, oh, ow;
  const bool valid = decode_row(a, t, row, q, oh, ow);
  const int hw = oh * a.OW + ow;
  const size_t res_off = ((size_t)q * ohw + hw) * a.Cout;
  const bool vec_io = (a.Cout & 7) == 0;
  int qd, hwd;                                  // row of the pass-image tensor and pixel inside it, as the mask counters see them
  dropout_pixel(a, q, hw, qd, hwd);
  const int n_img = a.rep > 1 ? qd : qd / a.T;
  const int tt0 = a.rep > 1 ? 0 : qd - n_img * a.T;
  // residual of chunk j (two 16-byte vectors), prefetched one chunk ahead so its L2 latency overlaps the previous chunk
  auto load_res = [&](int j, uint4& r0, uint4& r1) {
    r0 = make_uint4(0, 0, 0, 0); r1 = r0;
    const int c0 = t.nt * a.BN + j * 16;
    if (a.res && vec_io && valid && j < a.BN / 16 && c0 < a.Cout) {
      const uint4* rp = reinterpret_cast<const uint4*>(a.res + res_off + c0);
      r0 = __ldg(rp);
      if (c0 + 8 < a.Cout) r1 = __ldg(rp + 1);
    }
  };
  uint4 rn0, rn1;
  load_res(sub_w, rn0, rn1);
  for (int j = sub_w; j < a.BN / 16; j += WPQ) {
    const int c0 = t.nt * a.BN + j * 16;
    const uint4 rv0 = rn0, rv1 = rn1;
    load_res(j + WPQ, rn0, rn1);
    uint32_t acc[16];
    tmem_ld16(trow + uint32_t(j * 16), acc);       // warp-collective: executed by every lane, valid or not
    tmem_ld_wait();
    if (!valid || c0 >= a.Cout || (a.ablate & 8)) continue;
    float v[16];
    {
      const float4* bp = reinterpret_cast<const float4*>(a.bias + c0);     // bias is padded to cout_pad
      const uint32_t rw[8] = {rv0.x, rv0.y, rv0.z, rv0.w, rv1.x, rv1.y, rv1.z, rv1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = __ldg(bp + i);
        v[4 * i] = __uint_as_float(acc[4 * i]) + b.x;
        v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b.y;
        v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b.z;
        v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b.w;
      }
      if (a.res && vec_io) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo(rw[i]); v[2 * i + 1] += bf16_hi(rw[i]); }
      }
    }
    if (a.res && !vec_io) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < a.Cout) v[i] += __bfloat162float(a.res[res_off + c0 + i]);
    }
    const bool packed_relu = a.relu && !a.out_f32 && vec_io;     // ReLU after the pack (relu_bf16x2), same result
    if (a.relu && !packed_relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (a.drop) {                            // (bf16 output, Cout % 16 == 0: checked at launch)
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i] * a.drop_scale, v[2 * i + 1] * a.drop_scale);
      if (a.relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = relu_bf16x2(pk[i]);
      }
      const uint32_t e8 = uint32_t((size_t)hwd * a.Cout + c0) >> 4;
      for (int rp = 0; rp < n_rep; ++rp) {
        const int p_out = a.rep > 1 ? q * a.rep + rp : q;
        uint32_t o[8];
        dropout_and16(a, e8, a.first_image + uint32_t(n_img), uint32_t(a.rep > 1 ? rp : tt0), pk, o, a.drop_stream);
        if (a.drop2) {                       // dropout before fc on a 1x1 feature map: bf16 value x scale -> bf16, second mask
          uint32_t p2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) p2[i] = pack_bf16x2(bf16_lo(o[i]) * a.drop_scale, bf16_hi(o[i]) * a.drop_scale);
          dropout_and16(a, e8, a.first_image + uint32_t(n_img), uint32_t(tt0), p2, o, a.drop_stream2);
        }
        uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + ((size_t)p_out * ohw + hw) * a.Cout + c0);
        yp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        yp[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
      continue;
    }
    for (int rp = 0; rp < n_rep; ++rp) {
      const int p_out = a.rep > 1 ? q * a.rep + rp : q;
      const size_t off = ((size_t)p_out * ohw + hw) * a.Cout + c0;
      if (a.out_f32) {
        float* yp = reinterpret_cast<float*>(a.y) + off;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < a.Cout) yp[i] = v[i];
      } else if (vec_io) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (a.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = relu_bf16x2(pk[i]);
        }
        uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + off);
        yp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (c0 + 8 < a.Cout) yp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + off;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < a.Cout) yp[i] = __float2bfloat16_rn(v[i]);
      }
    }
  }
}

// ---- split-K epilogue (a.ksplit > 1; bf16 output, no dropout / replicas) ----
// part 1: this slice's fp32 accumulator tile -> acc32[tile][slice][row][col] (plain 16-byte stores)
template <int WPQ>
__device__ __forceinline__ void splitk_store_partials(const ConvArgs& a, const Tile& t, uint32_t trow, int row, int sub_w) {
  float* mine = a.acc32 + (((size_t)t.tb * a.ksplit + t.ks) * BM + row) * a.BN;
  for (int j = sub_w; j < a.BN / 16; j += WPQ) {
    uint32_t acc[16];
    tmem_ld16(trow + uint32_t(j * 16), acc);
    tmem_ld_wait();
    float4* d = reinterpret_cast<float4*>(mine + j * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      d[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
  }
}
// part 2: take a ticket for the output tile; the last slice to arrive sums all partials in slice order and finishes the tile
template <int WPQ>
__device__ __forceinline__ void splitk_fixup(const ConvArgs& a, const Tile& t, int row, int sub_w, volatile int* s_flag, bool leader) {
  __threadfence();                                   // this thread's partial stores are visible device-wide ...
  named_bar_sync(1, 128 * WPQ);                      // ... before the CTA's ticket is taken
  if (leader) *s_flag = atomicAdd(&a.tickets[t.tb], 1);
  named_bar_sync(1, 128 * WPQ);
  if (*s_flag != a.ksplit - 1) return;
  __threadfence();
  int q, oh, ow;
  const bool valid = decode_row(a, t, row, q, oh, ow);
  const size_t off = ((size_t)q * (a.OH * a.OW) + oh * a.OW + ow) * a.Cout;
  const bool vec_io = (a.Cout & 7) == 0;
  if (valid) {
    for (int j = sub_w; j < a.BN / 16; j += WPQ) {
      const int c0 = t.nt * a.BN + j * 16;
      if (c0 >= a.Cout) continue;
      float v[16];
      const float4* p0 = reinterpret_cast<const float4*>(a.acc32 + ((size_t)t.tb * a.ksplit * BM + row) * a.BN + j * 16);
      const size_t slice = (size_t)BM * a.BN / 4;    // float4 per slice
#pragma unroll
      for (int i = 0; i < 4; ++i) {                  // one 16-byte column at a time, all slices' loads in flight together
        float4 x[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) x[s] = s < a.ksplit ? __ldcg(p0 + s * slice + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // via L2
        float4 acc = x[0];
#pragma unroll
        for (int s = 1; s < 8; ++s) { acc.x += x[s].x; acc.y += x[s].y; acc.z += x[s].z; acc.w += x[s].w; }   // fixed slice order
        v[4 * i] = acc.x; v[4 * i + 1] = acc.y; v[4 * i + 2] = acc.z; v[4 * i + 3] = acc.w;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + c0 + i);
      if (a.res) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < a.Cout) v[i] += __bfloat162float(a.res[off + c0 + i]);
      }
      if (a.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + off + c0;
      if (vec_io) {
        reinterpret_cast<uint4*>(yp)[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        if (c0 + 8 < a.Cout)
          reinterpret_cast<uint4*>(yp)[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < a.Cout) yp[i] = __float2bfloat16_rn(v[i]);
      }
    }
  }
  if (leader) a.tickets[t.tb] = 0;                   // ready for the next launch
}

// Staged epilogue (a.stg_bytes > 0): as above, but the bf16 results go to a SWIZZLE_128B shared-memory slab of 128 pixels x
// 64 channels and one thread writes the slab with a 5-D TMA store (Cout, OW, OH, replica, image) -- whole 128-byte lines
// per pixel instead of 16 bytes per lane at a Cout*2-byte stride (32 L1 wavefronts per store instruction).  The box is the
// A-tile's pixel rectangle, so rows outside the image or beyond P are clipped by the tensor bounds.  Two slabs alternate:
// the leader waits for the previous store to finish reading before the barrier that releases the next slab's writers.
template <int WPQ, bool REP>   // REP: block 0 of an MC sweep writes a.rep masked replicas of every output row
__device__ __forceinline__ void conv_epilogue_staged(const ConvArgs& a, const CUtensorMap* tmY, const Tile& t, uint32_t trow, int row,
                                                     int sub_w, uint8_t* stg, int& seq, bool leader) {
  constexpr int NCH = 4 / WPQ;                         // 16-column chunks of a 64-channel slab per warp
  const int ohw = a.OH * a.OW, n_rep = REP ? a.rep : 1;
  int q, oh, ow;
  const bool valid = decode_row(a, t, row, q, oh, ow);
  const int hw = oh * a.OW + ow;
  const size_t res_off = ((size_t)q * ohw + hw) * a.Cout;
  int qd, hwd;                                  // row of the pass-image tensor and pixel inside it, as the mask counters see them
  dropout_pixel(a, q, hw, qd, hwd);
  const int n_img = REP ? qd : qd / a.T;
  const int tt0 = REP ? 0 : qd - n_img * a.T;
  const uint32_t stg_u32 = smem_u32(stg);
  const bool masked = a.drop && valid;
  // this row's two 16-byte slots of every chunk inside a slab (SWIZZLE_128B: slot index ^ (row & 7))
  uint32_t slot[NCH][2];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int jj = sub_w + ch * WPQ;
    slot[ch][0] = uint32_t(row * 128 + (((2 * jj) ^ (row & 7)) << 4));
    slot[ch][1] = uint32_t(row * 128 + (((2 * jj + 1) ^ (row & 7)) << 4));
  }
  for (int hf = 0; hf < a.BN / 64; ++hf) {
    uint32_t pk[NCH][8];                               // finished values (bias, residual, dropout scale, ReLU) as bf16 pairs
    uint4 rv[NCH][2];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {               // residual loads of all chunks first: their latency overlaps the TMEM loads
      const int c0 = t.nt * a.BN + hf * 64 + (sub_w + ch * WPQ) * 16;
      rv[ch][0] = make_uint4(0, 0, 0, 0); rv[ch][1] = rv[ch][0];
      if (a.res && valid && !(a.ablate & 16)) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + res_off + c0);
        rv[ch][0] = __ldg(rp); rv[ch][1] = __ldg(rp + 1);
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int jj = sub_w + ch * WPQ, c0 = t.nt * a.BN + hf * 64 + jj * 16;
      uint32_t acc[16];
      tmem_ld16(trow + uint32_t(hf * 64 + jj * 16), acc);       // warp-collective: executed by every lane, valid or not
      tmem_ld_wait();
      const float4* bp = reinterpret_cast<const float4*>(a.bias + c0);
      const uint32_t rw[8] = {rv[ch][0].x, rv[ch][0].y, rv[ch][0].z, rv[ch][0].w, rv[ch][1].x, rv[ch][1].y, rv[ch][1].z, rv[ch][1].w};
      float v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = __ldg(bp + i);
        v[4 * i] = __uint_as_float(acc[4 * i]) + b.x;
        v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b.y;
        v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b.z;
        v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b.w;
      }
      if (a.res) {                           // (null when the residual went through the identity MMAs)
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo(rw[i]); v[2 * i + 1] += bf16_hi(rw[i]); }
      }
      if (a.drop) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= a.drop_scale;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[ch][i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      if (a.relu) {                          // after the pack: one instruction per pair (relu_bf16x2)
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[ch][i] = relu_bf16x2(pk[ch][i]);
      }
    }
    const uint32_t e16 = uint32_t((size_t)hwd * a.Cout + t.nt * a.BN + hf * 64 + sub_w * 16) >> 4;   // chunk ch: + ch * WPQ
    if (!REP && masked) {                              // single pass: the masks go onto the values in place
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        dropout_and16(a, e16 + uint32_t(ch * WPQ), a.first_image + uint32_t(n_img), uint32_t(tt0), pk[ch], pk[ch], a.drop_stream);
    }
    for (int rp = 0; rp < n_rep; ++rp) {             // one slab (and one TMA store) per masked replica
      uint8_t* slab = stg + (seq & 1) * STG_SLAB_BYTES;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (REP && masked) {
          uint32_t o[8];
          dropout_and16(a, e16 + uint32_t(ch * WPQ), a.first_image + uint32_t(n_img), uint32_t(rp), pk[ch], o, a.drop_stream);
          *reinterpret_cast<uint4*>(slab + slot[ch][0]) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(slab + slot[ch][1]) = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
          *reinterpret_cast<uint4*>(slab + slot[ch][0]) = make_uint4(pk[ch][0], pk[ch][1], pk[ch][2], pk[ch][3]);
          *reinterpret_cast<uint4*>(slab + slot[ch][1]) = make_uint4(pk[ch][4], pk[ch][5], pk[ch][6], pk[ch][7]);
        }
      }
      fence_proxy_async();                             // generic-proxy smem writes -> visible to the TMA store
      if (leader) bulk_wait_read_all();                // the store that last used the OTHER slab has drained it
      named_bar_sync(1, 128 * WPQ);
      if (leader && !(a.ablate & 32)) {
        tma_store_5d(tmY, stg_u32 + uint32_t((seq & 1) * STG_SLAB_BYTES), t.nt * a.BN + hf * 64, t.ow0, t.oh0, rp, t.q0);
        bulk_commit_group();
      }
      ++seq;
    }
  }
}

}  // namespace fav
