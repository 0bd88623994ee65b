// conv_pair.cu -- K2, 2-SM variant for 128-output-channel convolutions (ResNet-18 layer2).
//
// The generic kernel streams a 16 KB activation tile AND a 16 KB weight tile per k-block per CTA and is bound by the
// 64 B/clk/SM L2->SM port (<= 50 % tensor pipe for 128x128 tiles).  Here two CTAs on one TPC form a cluster and issue
// tcgen05.mma.cta_group::2 with M = 256 (128 output pixels per CTA), N = 128: each CTA loads only HALF of every weight
// tile (64 of the 128 output channels), i.e. 24 KB instead of 32 KB per k-block, in a 9-deep ring.  MEASURED: correct but not
// faster -- the leader's M=256 MMAs issue at ~119 cycles each (role timers), so one CTA pair does less per SM than two
// independent 128x128 CTAs; kept as an opt-in variant (FAV_PAIR=1) and as the tested base for N=256 tiles.  (A weights-
// stationary layout -- the 147 KB half resident, activations only streaming -- was measured slower: it leaves room for
// just four 16 KB activation slots, too few bytes in flight to cover the L2 latency.)  The leader CTA issues the MMAs;
// tcgen05.commit multicasts completion to both CTAs; each CTA drains its own 128 TMEM lanes in the shared epilogue.
//
// Replaces (reference): nothing executable (see conv.cu header); oracle twin oracle/model.py.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "conv_dev.cuh"

namespace fav {

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box);

constexpr int PAIR_EPI_WARPS = 16;
constexpr int PAIR_THREADS = 32 * (2 + PAIR_EPI_WARPS);

__global__ void __launch_bounds__(PAIR_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to_1k = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to_1k;
  const uint32_t smem_base = smem_u32(smem);
  // [ring: stages x {A 16 KB, this CTA's half of the W tile: BN/2 output channels x 64 k}][barriers]
  const int w_half_bytes = (a.BN / 2) * 128;
  const int STAGE = A_TILE_BYTES + w_half_bytes;
  const uint32_t bars = smem_base + a.stages * STAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bars - smem_base) + (2 * a.stages + 4) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + 2 + i); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();
  const bool leader = cr == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // pair tile pt -> (pair of M tiles, N tile), N fastest so the CTA pairs running side by side share the A tiles in L2
  auto pair_tile = [&](int pt, int which) { return decode_tile(a, (2 * (pt / a.ntiles) + which) * a.ntiles + pt % a.ntiles); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (a.kb2 > 0) tma_prefetch_desc(&tmA2);
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 2u); mbar_init(empty_bar(s), 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1u); mbar_init(tempty_bar(i), uint32_t(2 * PAIR_EPI_WARPS)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), a.tmem_cols);
  tc_fence_before();
  cluster_sync_all();                         // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t a_bytes = uint32_t(a.bn_img * a.bh * a.bw) * 128u;

  if (warp == 0) {
    // ================================================================= TMA issuer (both CTAs)
    int stage = 0, phase = 0;
    long long w_empty = 0;
    const long long t_begin = clock64();
    for (int pt = pair; pt < a.pair_tiles; pt += n_pairs) {
      const Tile t = pair_tile(pt, int(cr));
      const int w_row = t.nt * a.BN + int(cr) * (a.BN / 2);          // this CTA's half of the W tile's rows
      auto arm = [&]() {
        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * (a_bytes + uint32_t(w_half_bytes)));
        else mbar_arrive_remote(full_bar(stage), 0);
      };
      for_each_kb(a, t, [&](int kb, int r, int ss, int cb) {
        mbar_wait_timed(empty_bar(stage), phase ^ 1, w_empty, a.stats != nullptr);
        if (elect_one()) {
          arm();
          const uint32_t sa = smem_base + stage * STAGE;
          tma_load_2d_2sm(sa + A_TILE_BYTES, &tmB, full_bar(stage), kb * BK, w_row);
          if (a.stride == 1) {
            tma_load_4d_2sm(sa, &tmA, full_bar(stage), cb * 64, t.ow0 + ss - a.pad, t.oh0 + r - a.pad, t.q0);
          } else {
            const int v = r - a.pad, w = ss - a.pad;
            tma_load_5d_2sm(sa, &tmA, full_bar(stage), (w & 1) * a.Cin + cb * 64, t.ow0 + (w >> 1), v & 1, t.oh0 + (v >> 1), t.q0);
          }
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      });
      for (int cb = 0; cb < a.kb2; ++cb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          arm();
          const uint32_t sa = smem_base + stage * STAGE;
          tma_load_2d_2sm(sa + A_TILE_BYTES, &tmB, full_bar(stage), (a.num_kb + cb) * BK, w_row);
          if (a.stride2 == 1) tma_load_4d_2sm(sa, &tmA2, full_bar(stage), cb * 64, t.ow0, t.oh0, t.q0);
          else tma_load_5d_2sm(sa, &tmA2, full_bar(stage), cb * 64, t.ow0, 0, t.oh0, t.q0);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (a.stats && lane == 0 && leader) {
      atomicAdd(&a.stats[0], (unsigned long long)w_empty);
      atomicAdd(&a.stats[1], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&a.stats[7], 1ull);
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (leader CTA only)
    if (leader) {
      int stage = 0, phase = 0, ti = 0;
      long long w_full = 0, w_tempty = 0;
      const long long t_begin = clock64();
      const uint64_t desc_a0 = make_sw128_desc(smem_base);
      for (int pt = pair; pt < a.pair_tiles; pt += n_pairs, ++ti) {
        const Tile t = pair_tile(pt, 0);
        const int acc = ti & 1;
        mbar_wait_timed(tempty_bar(acc), ((ti >> 1) & 1) ^ 1, w_tempty, a.stats != nullptr);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * a.BN);
        uint32_t accumulate = 0;
        auto issue = [&]() {
          mbar_wait_timed(full_bar(stage), phase, w_full, a.stats != nullptr);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc_a0 + uint64_t(stage) * uint64_t(STAGE >> 4);
            const uint64_t db = da + uint64_t(A_TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(d_tmem, da + 2u * k, db + 2u * k, a.idesc, k > 0 ? 1u : accumulate);
            umma_commit_2sm_mc(empty_bar(stage), 3);
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        };
        for_each_kb(a, t, [&](int, int, int, int) { issue(); });
        for (int cb = 0; cb < a.kb2; ++cb) issue();
        if (elect_one()) umma_commit_2sm_mc(tfull_bar(acc), 3);
        __syncwarp();
      }
      if (a.stats && lane == 0) {
        atomicAdd(&a.stats[2], (unsigned long long)w_full);
        atomicAdd(&a.stats[3], (unsigned long long)w_tempty);
        atomicAdd(&a.stats[4], (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ================================================================= epilogue warps (both CTAs, own 128 TMEM lanes)
    const int quarter = warp & 3, sub_w = (warp - EPI_WARP0) >> 2;
    const int row = quarter * 32 + lane;
    int ti = 0;
    long long w_tfull = 0;
    const long long t_begin = clock64();
    for (int pt = pair; pt < a.pair_tiles; pt += n_pairs, ++ti) {
      const int acc = ti & 1;
      const Tile t = pair_tile(pt, int(cr));
      mbar_wait_timed(tfull_bar(acc), (ti >> 1) & 1, w_tfull, a.stats != nullptr);
      tc_fence_after();
      conv_epilogue_subtile(a, t, tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * a.BN), row, sub_w, PAIR_EPI_WARPS / 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_remote(tempty_bar(acc), 0);
      }
    }
    if (a.stats && warp == EPI_WARP0 && lane == 0 && leader) {
      atomicAdd(&a.stats[5], (unsigned long long)w_tfull);
      atomicAdd(&a.stats[6], (unsigned long long)(clock64() - t_begin));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, a.tmem_cols);
}

static int pair_env() {
  // -1 (default): N = 256 pair tiles for layers with >= 256 (padded) output channels; the N = 128 form stays opt-in
  // (FAV_PAIR=1: measured ~15 % slower than two independent 128x128 CTAs per SM on ResNet-18 layer2,
  // profiles/r01k_pair_vs_generic.txt).  0 disables the 2-SM variant.
  static const int v = [] { const char* e = getenv("FAV_PAIR"); return e ? atoi(e) : -1; }();
  return v;
}

// pair-tile width for this layer (0 = not applicable).  a: fully prepared ConvArgs of the generic TMA path (a_mode 0).
// Both CTAs of a pair must walk the same k-blocks (the leader issues the MMAs for both), so tiles are whole images.
int conv_pair_bn(const ConvLayer& L, const ConvArgs& a, int force) {
  // (a flattened 1x1 conv of small images walks every k-block in every tile too: no taps to skip)
  const bool whole_images = a.tiles_w == 1 && a.tiles_h == 1;
  // larger images too when K is long enough to hide this kernel's direct epilogue (measured, ResNet-50 at T = 30: 1024 -> 256 on
  // 14x14 239 -> 189 us, 1024 -> 512 448 -> 364 us; but 64 -> 256 + downsample on 56x56, two k-blocks, 140 -> 217 us)
  const bool flat_small = a.flat_ohw > 0 && (a.flat_ohw <= BM || a.num_kb + a.kb2 >= 8) && L.r == 1 && L.s == 1;
  if (a.a_mode != 0 || a.stem_tma || L.bn != 128 || !(whole_images || flat_small)) return 0;
  const int env = pair_env();
  if (force == 0 && env == 0) return 0;
  // 1x1 convs (no fused downsample branch, not a folded 3x3) with a residual, a dropout mask or fp32 logits to write are
  // epilogue-bound: the generic kernel's staged TMA-store epilogue with the residual through the identity MMAs beats this
  // kernel's direct epilogue (ResNet-50 layer4 at T = 30: 512 -> 2048 + residual + dropout 465 -> 322 us, fc 50 -> 38 us), while
  // the plain 2048 -> 512 reduce conv is faster here (165 vs 193 us).  Decided from layer / sweep properties only.
  if (force == 0 && L.r == 1 && L.s == 1 && L.k2pad == 0 && !L.fold && (a.res != nullptr || a.drop || a.out_f32)) return 0;
  if ((L.cout_pad % 256) == 0 && (force || env != 0)) return 256;
  if (L.cout_pad == 128 && L.tmap64_ok && (force || env > 0)) return 128;
  return 0;
}
bool conv_pair_applicable(const ConvLayer& L, const ConvArgs& a, int force) { return conv_pair_bn(L, a, force) != 0; }

int conv_pair_launch(Ctx* ctx, const ConvLayer& L, ConvArgs a, const CUtensorMap& tmA, const CUtensorMap& tmA2, cudaStream_t st) {
  const int bn = conv_pair_bn(L, a, 1);
  FAV_REQUIRE(bn != 0, "conv: the 2-SM variant does not fit this layer");
  a.BN = bn;
  a.ntiles = L.cout_pad / bn; a.mt_per_tile = 1;
  a.fd_ntiles = make_fastdiv(uint32_t(a.ntiles));
  a.nkb_tot = a.num_kb + a.kb2;
  a.stg_bytes = 0;
  a.pair_tiles = ((a.mtiles + 1) / 2) * a.ntiles;
  const int stage_bytes = A_TILE_BYTES + (bn / 2) * 128;                   // 24 KB (N = 128) or 32 KB (N = 256) per k-block and CTA
  a.stages = (227 * 1024 - 1024 - 512) / stage_bytes;
  a.tmem_cols = 2 * bn;                                                    // two accumulators of N columns
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(256 >> 4) << 24);   // M = 256, N = bn
  const size_t smem = (size_t)a.stages * stage_bytes + 1024 + 512;
  if (!ctx->attr_pair) {
    FAV_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->attr_pair = true;
  }
  const int max_pairs = ctx->num_sms / 2;
  const int n_pairs = a.pair_tiles < max_pairs ? a.pair_tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n_pairs); cfg.blockDim = dim3(PAIR_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  // N = 256: each CTA loads 128 of the tile's 256 W rows with the layer's standard 128-row box; N = 128: the 64-row box
  const CUtensorMap& tmW = *reinterpret_cast<const CUtensorMap*>(bn == 256 ? L.tmap_w : L.tmap_w64);
  FAV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_pair_kernel, tmA, tmW, tmA2, a));
  return FAV_OK;
}

}  // namespace fav
