// conv.cu -- K2: implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 x bf16 -> fp32.
//
// Replaces (reference): nothing executable -- the reference names "image classification" as its perception
// task (README.md:15-20) and lists torchvision in requirements.txt:2 but ships no classifier; this is the
// forward of stock torchvision ResNet-18/50 with BN folded (oracle twin: oracle/model.py).
//
// GEMM view:  D[M = pixels, N = Cout] = A[M, K = R*S*Cin] * W[N, K]^T
//   A tile  128 x 64 bf16, K-major, SWIZZLE_128B in shared memory
//        a_mode 0: TMA tiled load of the NHWC activation with a 4-D box (64 ch, bw, bh, bn images) per filter
//                  tap; the box origin is shifted by (s - pad, r - pad) and TMA's out-of-bounds zero fill
//                  supplies the padding halo (stride-1 convs, Cin % 64 == 0)
//        a_mode 1: 128 producer threads gather 16-byte channel vectors (any stride / padding, Cin % 8 == 0)
//        a_mode 2: scalar gather (the 7x7 stem, Cin = 3)
//   W tile  BN x 64 bf16 by TMA (2-D box, SWIZZLE_128B), BN in {16..256}
//   D       128 lanes x BN fp32 columns in TMEM, one tcgen05.mma (M=128, N=BN, K=16) per 32 bytes of K
//   epilogue: tcgen05.ld -> +bias (+residual) -> ReLU -> MC-dropout mask from Philox (optionally T masked
//             replicas of a pass-invariant tile) -> bf16 NHWC (or fp32 logits)
// Persistent CTAs, warp-specialised (TMA issuer, MMA issuer, 8 epilogue warps, 4 optional gather warps); smem ring of
// `stages` {A,B} slots with full/empty mbarriers, tcgen05.commit frees slots; two TMEM accumulators so the epilogue of
// one tile overlaps the MMAs of the next.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "conv_dev.cuh"

namespace fav {

bool conv_pair_applicable(const ConvLayer& L, const ConvArgs& a, int force);
int conv_pair_launch(Ctx* ctx, const ConvLayer& L, ConvArgs a, const CUtensorMap& tmA, const CUtensorMap& tmA2, cudaStream_t st);

// ------------------------------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: each CTA walks tiles (tile = blockIdx.x + i * gridDim.x; N-tile fastest so CTAs
// running side by side share the activation tile in L2).  Two TMEM accumulators let the epilogue of tile i overlap
// the MMAs of tile i+1.
//   warp 0      TMA issuer (one lane)            warp 1      MMA issuer (one lane) + TMEM alloc/dealloc
//   warps 2-9   epilogue: TMEM lane quarter = warp % 4, the two warps of a quarter split the 16-column chunks
//   warps 10-13 gather producers (a_mode 1/2/3 only; not launched in a_mode 0)
//   MT = 2: each CTA tile is 256 output pixels (two A tiles sharing one W tile per k-block, two accumulators) --
//           1.36x fewer L2->smem bytes per FLOP; the conv kernels are L2-bandwidth-bound, so this is the main lever.
template <bool GATHER, int MT, int EPI_WARPS>
__device__ __forceinline__ void conv_igemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                                                const CUtensorMap& tmY, const CUtensorMap& tmR, const ConvArgs& a) {
  constexpr int GATHER_WARP0 = EPI_WARP0 + EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to_1k = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to_1k;
  const int stage_bytes = MT * A_TILE_BYTES + a.BN * 128;
  const uint32_t smem_base = smem_u32(smem);
  uint8_t* stg = smem + a.stages * stage_bytes;                  // staged epilogue: two 16 KB output slabs (a.stg_bytes, may be 0)
  const int ident_bytes = a.res_mma ? IDENT_BYTES : 0;            // resident 64x64 bf16 identity (B operand of the residual MMAs)
  uint8_t* ident = stg + a.stg_bytes;
  const int fixed = a.stages * stage_bytes + a.stg_bytes + ident_bytes;
  const uint32_t bars = smem_base + fixed;                       // full[s], empty[s], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + fixed + (2 * a.stages + 4) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + 2 + i); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (a.a_mode == 0) tma_prefetch_desc(&tmA);
    if (a.kb2 > 0) tma_prefetch_desc(&tmA2);
    if (a.stg_bytes) tma_prefetch_desc(&tmY);
    if (a.res_mma) tma_prefetch_desc(&tmR);
    const uint32_t full_count = a.a_mode == 0 ? 1u : 1u + 128u;
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), full_count); mbar_init(empty_bar(s), 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1u); mbar_init(tempty_bar(i), uint32_t(EPI_WARPS)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
  if (a.res_mma && warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // identity in the canonical K-major SWIZZLE_128B layout: row n = 128 bytes, 1.0 at k = n, 16-byte chunk index ^= n % 8
    const int i = threadIdx.x - 32 * EPI_WARP0;                    // 0..127: two threads per row (64 rows), 64 bytes each
    const int n = i >> 1, half = i & 1;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int chunk = half * 4 + ch;                             // logical chunk: k = 8 * chunk .. 8 * chunk + 7
      uint4 v = make_uint4(0, 0, 0, 0);
      if ((n >> 3) == chunk) {
        const uint32_t one = 0x3F80u << (16 * (n & 1));
        const int w = (n & 7) >> 1;
        v.x = w == 0 ? one : 0u; v.y = w == 1 ? one : 0u; v.z = w == 2 ? one : 0u; v.w = w == 3 ? one : 0u;
      }
      *reinterpret_cast<uint4*>(ident + n * 128 + ((chunk ^ (n & 7)) << 4)) = v;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on we touch activations it produced (and buffers it may still be reading)
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================================================================= TMA issuer (warp converged, one elected lane issues)
    {
      const uint32_t a_bytes = a.a_mode == 0 ? uint32_t(a.bn_img * a.bh * a.bw) * 128u : 0u;
      int stage = 0, phase = 0;
      long long w_empty = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        Tile t[MT];
        int n_sub = 0;
#pragma unroll
        for (int u = 0; u < MT; ++u) { t[u] = decode_tile(a, tile, u); if (t[u].mt < a.mtiles) n_sub = u + 1; }
        const uint32_t tx = ((a.ablate & 1) ? 0u : a_bytes * n_sub) + ((a.ablate & 2) ? 0u : uint32_t(a.BN) * 128u);
        const int n_col = t[0].nt * a.BN;
        int cd = for_each_kb(a, t[0], [&](int kb, int r, int ss, int cb) {
          mbar_wait_timed(empty_bar(stage), phase ^ 1, w_empty, a.stats != nullptr);
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), tx);
          if (a.a_mode == 0 && !(a.ablate & 1)) {
#pragma unroll
            for (int u = 0; u < MT; ++u) {
              if (u >= n_sub) break;
              if (a.stem_tma) {
                // space-to-depth input viewed as (64 = 4 taps x 16 ch, tap row, ow [32-byte step], oh, image):
                // k-block kb = tap row kb -> 128 contiguous bytes per output pixel, no bounds handling needed
                tma_load_5d(sa + u * A_TILE_BYTES, &tmA, full_bar(stage), 0, kb, t[u].ow0, t[u].oh0, t[u].q0);
              } else if (a.stride == 1) {
                tma_load_4d(sa + u * A_TILE_BYTES, &tmA, full_bar(stage), cb * 64, t[u].ow0 + ss - a.pad, t[u].oh0 + r - a.pad, t[u].q0);
              } else {
                // stride 2: input row 2*oh + v (v = r - pad) = 2*(oh + (v >> 1)) + (v & 1); the tensor map views the
                // activation as (2*Cin [w parity folded into channels], W/2, 2 [h parity], H/2, P)
                const int v = r - a.pad, w = ss - a.pad;
                tma_load_5d(sa + u * A_TILE_BYTES, &tmA, full_bar(stage), (w & 1) * a.Cin + cb * 64, t[u].ow0 + (w >> 1), v & 1,
                            t[u].oh0 + (v >> 1), t[u].q0);
              }
            }
          }
          if (!(a.ablate & 2)) tma_load_2d(sa + MT * A_TILE_BYTES, &tmB, full_bar(stage), kb * BK, n_col);
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        });
        for (int cb = 0; cb < a.kb2; ++cb) {        // fused downsample branch: 1x1 taps of the block input, same output tile
          if (!kb_mine(a, cd)) continue;
          mbar_wait_timed(empty_bar(stage), phase ^ 1, w_empty, a.stats != nullptr);
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), tx);
#pragma unroll
          for (int u = 0; u < MT; ++u) {
            if (u >= n_sub) break;
            if (a.stride2 == 1) tma_load_4d(sa + u * A_TILE_BYTES, &tmA2, full_bar(stage), cb * 64, t[u].ow0, t[u].oh0, t[u].q0);
            else tma_load_5d(sa + u * A_TILE_BYTES, &tmA2, full_bar(stage), cb * 64, t[u].ow0, 0, t[u].oh0, t[u].q0);
          }
          tma_load_2d(sa + MT * A_TILE_BYTES, &tmB, full_bar(stage), (a.num_kb + cb) * BK, n_col);
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (a.res_mma) {                             // residual rows of this tile, 64 channels per k-block (no W tile: B is resident)
          for (int j = 0; j < a.BN / 64; ++j) {
            mbar_wait_timed(empty_bar(stage), phase ^ 1, w_empty, a.stats != nullptr);
            const uint32_t sa = smem_base + stage * stage_bytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar(stage), a_bytes * n_sub);
#pragma unroll
              for (int u = 0; u < MT; ++u)
                if (u < n_sub) tma_load_4d(sa + u * A_TILE_BYTES, &tmR, full_bar(stage), n_col + j * 64, t[u].ow0, t[u].oh0, t[u].q0);
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], (unsigned long long)w_empty);
        atomicAdd(&a.stats[1], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&a.stats[7], 1ull);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (warp converged, one elected lane issues)
    {
      int stage = 0, phase = 0, ti = 0;
      const uint64_t desc0 = make_sw128_desc(smem_base);
      const uint64_t desc_stage = uint64_t(stage_bytes >> 4);      // descriptor address field is in 16-byte units
      long long w_full = 0, w_tempty = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        const Tile t = decode_tile(a, tile, 0);
        const int n_sub = (MT == 2 && decode_tile(a, tile, 1).mt < a.mtiles) ? 2 : 1;
        const int acc = ti & 1;
        mbar_wait_timed(tempty_bar(acc), ((ti >> 1) & 1) ^ 1, w_tempty, a.stats != nullptr);   // epilogue drained this accumulator set
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * MT * a.BN);
        uint32_t accumulate = 0;
        auto issue_stage = [&]() {
          mbar_wait_timed(full_bar(stage), phase, w_full, a.stats != nullptr);
          tc_fence_after();
          const uint64_t da0 = desc0 + uint64_t(stage) * desc_stage, db = da0 + uint64_t(MT * (A_TILE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
          for (int u = 0; u < MT; ++u) {
            if (u >= n_sub) break;
            const uint64_t da = da0 + uint64_t(u * (A_TILE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)          // 32 bytes of K per MMA: +2 in the (addr >> 4) field
              if (!(a.ablate & 4)) umma_f16(d_tmem + uint32_t(u * a.BN), da + 2u * k, db + 2u * k, a.idesc, k > 0 ? 1u : accumulate);
          }
          umma_commit(empty_bar(stage));
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        };
        int cd = for_each_kb(a, t, [&](int, int, int, int) { issue_stage(); });
        for (int cb = 0; cb < a.kb2; ++cb)
          if (kb_mine(a, cd)) issue_stage();
        if (a.res_mma) {
          const uint64_t d_ident = make_sw128_desc(smem_u32(ident));
          for (int j = 0; j < a.BN / 64; ++j) {      // acc[:, 64j .. 64j+63] += residual[:, 64j ..] x I
            mbar_wait_timed(full_bar(stage), phase, w_full, a.stats != nullptr);
            tc_fence_after();
            const uint64_t da0 = desc0 + uint64_t(stage) * desc_stage;
            if (elect_one()) {
#pragma unroll
              for (int u = 0; u < MT; ++u) {
                if (u >= n_sub) break;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_f16(d_tmem + uint32_t(u * a.BN + j * 64), da0 + uint64_t(u * (A_TILE_BYTES >> 4)) + 2u * k, d_ident + 2u * k,
                           a.idesc64, 1u);
              }
              umma_commit(empty_bar(stage));
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(tfull_bar(acc));
        __syncwarp();
      }
      if (a.stats && lane == 0) {
        atomicAdd(&a.stats[2], (unsigned long long)w_full);
        atomicAdd(&a.stats[3], (unsigned long long)w_tempty);
        atomicAdd(&a.stats[4], (unsigned long long)(clock64() - t_begin));
      }
    }
  } else if (GATHER && warp >= GATHER_WARP0) {
    // ================================================================= gather producers (a_mode 1/2/3)
    // two groups of 4 warps take alternate k-blocks, so the load latency of one k-block overlaps the stores of the other
    const int gtid = threadIdx.x - 32 * GATHER_WARP0;
    const int row = gtid & 127, group = gtid >> 7;        // row = A-tile row
    int stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const Tile t = decode_tile(a, tile);
      int q, oh, ow;
      const bool valid = decode_row(a, t, row, q, oh, ow);
      const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
      const __nv_bfloat16* ximg = a.x + (size_t)q * a.H * a.W * a.Cin;
      for (int kb = 0; kb < a.num_kb; ++kb, ++it) {
        if ((it & 1) != group) {
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_wait(empty_bar(stage), phase ^ 1);
        uint8_t* rowp = smem + stage * stage_bytes + row * 128;
        if (a.a_mode == 3) {
          // channel-padded stem: a 16-byte chunk = two horizontally adjacent filter taps x 4 channels.  All 16 loads of
          // the k-block are issued before any store (clamped addresses + predicated zeroing: no branches in between).
          uint2 lo[8], hi[8];
          const int half = a.s_store >> 1;
          const uint2* xi = reinterpret_cast<const uint2*>(ximg);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int pair = (kb * BK + c * 8) >> 3, r = pair / half, ss = (pair - r * half) * 2;
            const int ih = ih0 + r, iw = iw0 + ss;
            const bool okr = valid && r < a.R && ih >= 0 && ih < a.H;
            const bool ok0 = okr && iw >= 0 && iw < a.W && ss < a.S, ok1 = okr && iw + 1 >= 0 && iw + 1 < a.W && ss + 1 < a.S;
            const size_t base = (size_t)min(max(ih, 0), a.H - 1) * a.W;
            lo[c] = __ldg(xi + base + min(max(iw, 0), a.W - 1));
            hi[c] = __ldg(xi + base + min(max(iw + 1, 0), a.W - 1));
            if (!ok0) lo[c] = make_uint2(0, 0);
            if (!ok1) hi[c] = make_uint2(0, 0);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = make_uint4(lo[c].x, lo[c].y, hi[c].x, hi[c].y);
        } else if (a.a_mode == 1) {
          uint4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k = kb * BK + c * 8;
            const int tap = k / a.Cin, ci = k - tap * a.Cin, r = tap / a.S, ss = tap - r * a.S;
            const int ih = ih0 + r, iw = iw0 + ss;
            const bool ok = valid && k < a.K && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
            const size_t off = ((size_t)min(max(ih, 0), a.H - 1) * a.W + min(max(iw, 0), a.W - 1)) * a.Cin + min(ci, a.Cin - 8);
            v[c] = __ldg(reinterpret_cast<const uint4*>(ximg + off));
            if (!ok) v[c] = make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = v[c];
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k = kb * BK + c * 8;
            uint32_t w4[4] = {0, 0, 0, 0};
            if (valid && k < a.K) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int kk = k + e;
                if (kk < a.K) {
                  const int tap = kk / a.Cin, ci = kk - tap * a.Cin, r = tap / a.S, ss = tap - r * a.S;
                  const int ih = ih0 + r, iw = iw0 + ss;
                  if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) {
                    const uint32_t b = __ldg(reinterpret_cast<const unsigned short*>(ximg) + ((size_t)ih * a.W + iw) * a.Cin + ci);
                    w4[e >> 1] |= b << (16 * (e & 1));
                  }
                }
              }
            }
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);      // 128B swizzle: chunk ^= row % 8
          }
        }
        fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(full_bar(stage));
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================================================================= epilogue warps
    const int quarter = warp & 3;                        // TMEM lanes 32*quarter .. +31 (hardware: warp id % 4)
    const int sub_w = (warp - EPI_WARP0) >> 2;           // which of the EPI_WARPS/4 warps of this quarter
    constexpr int WPQ = EPI_WARPS / 4;
    const int row = quarter * 32 + lane;
    int ti = 0, seq = 0;
    const bool leader = warp == EPI_WARP0 && lane == 0;      // issues the staged epilogue's TMA stores
    long long w_tfull = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      const int acc_i = ti & 1;
      mbar_wait_timed(tfull_bar(acc_i), (ti >> 1) & 1, w_tfull, a.stats != nullptr);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < MT; ++u) {
      const Tile t = decode_tile(a, tile, u);
      if (t.mt >= a.mtiles) break;
      const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t((acc_i * MT + u) * a.BN);
      if (!GATHER && MT == 1 && a.ksplit > 1) splitk_store_partials<WPQ>(a, t, trow, row, sub_w);
      else if (!GATHER && a.stg_bytes) {
        if (a.rep > 1) conv_epilogue_staged<WPQ, true>(a, &tmY, t, trow, row, sub_w, stg, seq, leader);
        else conv_epilogue_staged<WPQ, false>(a, &tmY, t, trow, row, sub_w, stg, seq, leader);
      }
      else conv_epilogue_subtile(a, t, trow, row, sub_w, WPQ);
      }   // sub-tiles
      // this warp has finished reading the accumulators: hand them back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc_i));
      if (!GATHER && MT == 1 && a.ksplit > 1)          // accumulator already released: the fix-up only touches global memory
        splitk_fixup<WPQ>(a, decode_tile(a, tile, 0), row, sub_w, reinterpret_cast<volatile int*>(tmem_slot + 1), leader);
    }
    if (leader && a.stg_bytes) bulk_wait_all();              // shared memory must outlive the last store's reads
    if (a.stats && warp == EPI_WARP0 && lane == 0) {
      atomicAdd(&a.stats[5], (unsigned long long)w_tfull);
      atomicAdd(&a.stats[6], (unsigned long long)(clock64() - t_begin));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// TMA-fed variants: MT=1 (320 threads, two CTAs per SM) and MT=2 (576 threads, one CTA per SM); gather variant.
__global__ void __launch_bounds__(THREADS_TMA1, 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmR, const ConvArgs a) {
  conv_igemm_body<false, 1, 8>(tmA, tmB, tmA2, tmY, tmR, a);
}
__global__ void __launch_bounds__(THREADS_TMA2, 1)
conv_igemm_m256_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmR, const ConvArgs a) {
  conv_igemm_body<false, 2, 16>(tmA, tmB, tmA2, tmY, tmR, a);
}
__global__ void __launch_bounds__(THREADS_GATHER, 1)
conv_igemm_gather_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmR, const ConvArgs a) {
  conv_igemm_body<true, 1, 8>(tmA, tmB, tmA2, tmY, tmR, a);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return FAV_E_CUDA; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", int(r), rank); return FAV_E_CUDA; }
  return FAV_OK;
}

// optional per-launch timing (bench roofline / tuning): records the start event, returns the stop event to record after
// the launch and a zeroed 8-counter stats slot for the kernel's role timers
int conv_timing_begin(Ctx* ctx, cudaStream_t st, float gflop, float gbyte, cudaEvent_t* stop, unsigned long long** stats) {
  *stop = nullptr;
  *stats = nullptr;
  if (!ctx->timing) return FAV_OK;
  while (ctx->ev_pool.size() < ctx->ev_used + 2) {
    cudaEvent_t e;
    FAV_CUDA_OK(cudaEventCreate(&e));
    ctx->ev_pool.push_back(e);
  }
  cudaEvent_t e0 = ctx->ev_pool[ctx->ev_used];
  *stop = ctx->ev_pool[ctx->ev_used + 1];
  ctx->ev_used += 2;
  ctx->ev_gflop.push_back(gflop);
  ctx->ev_gbyte.push_back(gbyte);
  if (!ctx->stats_buf) FAV_CUDA_OK(cudaMalloc(&ctx->stats_buf, 512 * 8 * sizeof(unsigned long long)));
  const size_t li = ctx->ev_used / 2 - 1;
  if (li < 512) {
    *stats = reinterpret_cast<unsigned long long*>(ctx->stats_buf) + 8 * li;
    FAV_CUDA_OK(cudaMemsetAsync(*stats, 0, 8 * sizeof(unsigned long long), st));
  }
  FAV_CUDA_OK(cudaEventRecord(e0, st));
  return FAV_OK;
}

int conv_out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }
int conv_pick_bn(int cout) { return cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : 128; }

// a_mode 5 (stride-2 stem, R, S <= 8): the zero-padded image is stored 2x2 space-to-depth, x = [p][hp][wp][16] bf16 with
// channel (dy*2 + dx)*4 + c of pixel (Y, X) = padded image (2Y + dy, 2X + dx, c).  The conv becomes 4x4 / stride 1 over
// 16 channels: one k-block = one row of 4 taps x 16 channels = 128 contiguous bytes starting at pixel (oh + r', ow), so
// A is one 5-D TMA box per k-block with no bounds handling (windows of neighbouring pixels overlap, 32-byte step).
// A SWIZZLE_128B box narrower than 128 bytes faults on sm_100 (tools/probe/tma_probe.cu), which rules out feeding
// 8-tap NHWC4 rows directly.
bool conv_stem_padded_dims(const ConvLayer& L, int h, int w, int* hp, int* wp) {
  if (!L.s2d) return false;
  const int oh = conv_out_dim(h, L.r, L.stride, L.pad), ow = conv_out_dim(w, L.s, L.stride, L.pad);
  if (oh <= 0 || ow <= 0) return false;
  *hp = oh + 3;
  *wp = ow + 3;
  return true;
}

int conv_layer_finalize(ConvLayer& L) {
  L.k = L.s2d ? 256 : L.r * (L.s_store ? L.s_store : L.s) * (L.cin_store ? L.cin_store : L.cin);
  L.kpad = (L.k + BK - 1) / BK * BK;
  L.bn = conv_pick_bn(L.cout);
  L.cout_pad = (L.cout + L.bn - 1) / L.bn * L.bn;
  if (L.w) {
    const cuuint64_t dims[2] = {(cuuint64_t)(L.kpad + L.k2pad), (cuuint64_t)L.cout_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)(L.kpad + L.k2pad) * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)L.bn};
    int rc = encode_map(reinterpret_cast<CUtensorMap*>(L.tmap_w), L.w, 2, dims, strides, box);
    if (rc) return rc;
    L.tmap_ok = true;
    if (L.bn == 128) {
      const cuuint32_t box64[2] = {BK, 64}, box32[2] = {BK, 32};
      rc = encode_map(reinterpret_cast<CUtensorMap*>(L.tmap_w64), L.w, 2, dims, strides, box64);
      if (rc) return rc;
      L.tmap64_ok = true;
      rc = encode_map(reinterpret_cast<CUtensorMap*>(L.tmap_w32), L.w, 2, dims, strides, box32);
      if (rc) return rc;
      L.tmap32_ok = true;
    }
  }
  return FAV_OK;
}

int conv_launch(Ctx* ctx, const ConvCall& c_in, cudaStream_t st) {
  ConvCall c = c_in;
  const ConvLayer& L = *c.L;
  FAV_REQUIRE(L.tmap_ok, "conv: layer not finalized");
  FAV_REQUIRE(c.p > 0 && c.h > 0 && c.w > 0, "conv: bad shape p=%d h=%d w=%d", c.p, c.h, c.w);
  // 1x1 / stride 1 convolutions on images larger than one tile are pixel-local: present the whole batch as ONE image of
  // 1 x (P*H*W) pixels, so every tile is 128 consecutive pixels instead of a (bw x bh <= 128) rectangle that has to divide the
  // image (56x56 -> 56x2 = 112 of 128 rows, 14x14 -> 14x9 + 14x5 = 77 %, 7x7 -> two images = 98 rows).  NHWC is contiguous, so x / y / residual addresses are
  // unchanged; only the dropout counters need the (image, pixel) split back (ConvArgs::flat_ohw).  Results are bit-identical.
  static const int env_flat1 = [] { const char* e = getenv("FAV_FLAT1X1"); return e ? atoi(e) : 1; }();
  int flat_ohw = 0;
  if (env_flat1 && L.r == 1 && L.s == 1 && L.stride == 1 && L.pad == 0 && !L.cin_store && !L.s2d && !L.fold && (L.cin % 64) == 0 &&
      c.a_mode <= 0 && c.rep <= 1 && c.drop2_layer < 0 && (c.h * c.w > BM || BM % (c.h * c.w) != 0) &&
      (long long)c.p * c.h * c.w < (1ll << 31) &&
      (L.k2pad == 0 || (L.stride2 == 1 && c.h2 == c.h && c.w2 == c.w))) {
    flat_ohw = c.h * c.w;
    c.w = c.p * c.h * c.w; c.h = 1; c.p = 1;
    if (L.k2pad) { c.w2 = c.w; c.h2 = 1; }
  }
  if (conv_flat_applicable(c)) return conv_flat_launch(ctx, c, st);
  FAV_REQUIRE(c.a_mode != 4, "conv: a_mode 4 (flat-padded 3x3) needs 3x3/s1/p1, Cin = Cout = 64 and (H+1)(W+1) <= 256 or 3 (W+1) <= 255 (band mode)");
  ConvArgs a{};
  a.x = reinterpret_cast<const __nv_bfloat16*>(c.x); a.y = c.y; a.bias = L.bias;
  a.res = reinterpret_cast<const __nv_bfloat16*>(c.res);
  a.P = c.p; a.H = c.h; a.W = c.w; a.Cin = L.cin_store ? L.cin_store : L.cin; a.Cout = L.cout;
  a.s_store = L.s_store;
  a.R = L.r; a.S = L.s; a.stride = L.stride; a.pad = L.pad;
  a.OH = conv_out_dim(c.h, L.r, L.stride, L.pad); a.OW = conv_out_dim(c.w, L.s, L.stride, L.pad);
  FAV_REQUIRE(a.OH > 0 && a.OW > 0, "conv: empty output");
  a.K = L.k; a.num_kb = L.kpad / BK; a.BN = L.bn;
  const long long M = (long long)c.p * a.OH * a.OW;
  FAV_REQUIRE(M < (1ll << 31), "conv: too many output pixels (%lld)", M);
  a.M = int(M);
  a.relu = c.relu; a.out_f32 = c.out_f32;
  a.T = c.T > 0 ? c.T : 1; a.rep = c.rep > 1 ? c.rep : 1; a.drop = c.drop;
  if (c.drop) {
    FAV_REQUIRE((L.cout & 15) == 0 && !c.out_f32, "conv: dropout epilogue needs Cout %% 16 == 0 and bf16 output");
    FAV_REQUIRE(c.p_drop >= 0.f && c.p_drop < 1.f, "conv: p_drop must be in [0,1)");
    a.drop_thr8 = dropout_thr8(c.p_drop);
    a.drop_add4 = dropout_add4(a.drop_thr8); a.drop_hi4 = dropout_hi4(a.drop_thr8);
    a.drop_scale = dropout_scale8(a.drop_thr8);
    a.k0 = uint32_t(c.seed); a.k1 = uint32_t(c.seed >> 32); a.first_image = uint32_t(c.first_image);
    a.drop_keys = philox_keys(a.k0, a.k1);
    a.drop_stream = stream_id(KIND_DROPOUT, c.layer_id, 0);
    if (c.drop2_layer >= 0) {
      FAV_REQUIRE(a.rep == 1 && a.OH * a.OW == 1, "conv: the fused second dropout needs a 1x1 output map and no replicas");
      a.drop2 = 1;
      a.drop_stream2 = stream_id(KIND_DROPOUT, c.drop2_layer, 0);
    }
  }
  // operand-A mode
  const bool tma_s1 = L.stride == 1 && 2 * L.pad == L.r - 1;
  const bool tma_s2 = L.stride == 2 && (c.h % 2) == 0 && (c.w % 2) == 0;
  const bool tma_ok = !L.cin_store && (L.cin % 64) == 0 && L.r == L.s && (tma_s1 || tma_s2);
  int mode = c.a_mode;
  const bool stem_tma = c.a_mode == 5;
  if (stem_tma) {
    int hp, wp;
    FAV_REQUIRE(conv_stem_padded_dims(L, c.h, c.w, &hp, &wp), "conv: a_mode 5 needs the space-to-depth stem layout (stride 2, R, S <= 8)");
    mode = 0;
  } else if (L.cin_store) {
    FAV_REQUIRE(!L.s2d, "conv: a space-to-depth stem runs in a_mode 5 only");
    mode = 3;
  }
  if (mode < 0) mode = tma_ok ? 0 : ((L.cin % 8) == 0 ? 1 : 2);
  FAV_REQUIRE(mode != 0 || tma_ok || stem_tma, "conv: a_mode 0 (TMA) needs Cin %% 64 == 0 and stride 1 with 'same' padding or stride 2 with even H, W");
  FAV_REQUIRE(mode != 1 || (L.cin % 8) == 0, "conv: a_mode 1 needs Cin %% 8 == 0");
  FAV_REQUIRE(mode != 3 || (L.cin_store == 4 && (L.s_store % 2) == 0), "conv: a_mode 3 needs the channel-padded stem layout");
  a.a_mode = mode;
  a.flat_ohw = flat_ohw;
  a.flat_m64 = flat_ohw ? ~0ull / (unsigned long long)flat_ohw + 1ull : 0ull;
  CUtensorMap tmA;
  memset(&tmA, 0, sizeof(tmA));
  int mtiles;
  a.stem_tma = stem_tma ? 1 : 0;
  if (mode == 0) {
    if (a.OH * a.OW <= BM) { a.bw = a.OW; a.bh = a.OH; a.bn_img = BM / (a.OH * a.OW); }
    else {
      // the (bw x bh <= 128) pixel rectangle that covers the output with the fewest tiles (ties: the widest one)
      long long best = -1;
      for (int bw = 1; bw <= (a.OW < BM ? a.OW : BM); ++bw) {
        const int bh = (BM / bw) < a.OH ? (BM / bw) : a.OH;
        const long long tiles = (long long)((a.OW + bw - 1) / bw) * ((a.OH + bh - 1) / bh);
        if (best < 0 || tiles <= best) { best = tiles; a.bw = bw; a.bh = bh; }
      }
      a.bn_img = 1;
    }
    if (a.bn_img > c.p) a.bn_img = c.p;
    a.tiles_w = (a.OW + a.bw - 1) / a.bw; a.tiles_h = (a.OH + a.bh - 1) / a.bh;
    const int tiles_n = (c.p + a.bn_img - 1) / a.bn_img;
    a.cin_blocks = L.cin / 64;
    mtiles = a.tiles_w * a.tiles_h * tiles_n;
    int rc;
    if (stem_tma) {
      int hp, wp;
      conv_stem_padded_dims(L, c.h, c.w, &hp, &wp);
      const cuuint64_t row = (cuuint64_t)wp * 32;      // bytes per space-to-depth row (16 bf16 per pixel)
      const cuuint64_t dims[5] = {64, (cuuint64_t)a.num_kb, (cuuint64_t)a.OW, (cuuint64_t)a.OH, (cuuint64_t)c.p};
      const cuuint64_t strides[4] = {row, 32, row, (cuuint64_t)hp * row};
      const cuuint32_t box[5] = {64, 1, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
      rc = encode_map(&tmA, c.x, 5, dims, strides, box);
    } else if (L.stride == 1) {
      const cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)c.w, (cuuint64_t)c.h, (cuuint64_t)c.p};
      const cuuint64_t strides[3] = {(cuuint64_t)L.cin * 2, (cuuint64_t)c.w * L.cin * 2, (cuuint64_t)c.h * c.w * L.cin * 2};
      const cuuint32_t box[4] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
      rc = encode_map(&tmA, c.x, 4, dims, strides, box);
    } else {
      // parity-split view for stride 2: (2*Cin [w parity folded in], W/2, 2 [h parity], H/2, P)
      const cuuint64_t dims[5] = {(cuuint64_t)L.cin * 2, (cuuint64_t)c.w / 2, 2, (cuuint64_t)c.h / 2, (cuuint64_t)c.p};
      const cuuint64_t strides[4] = {(cuuint64_t)L.cin * 4, (cuuint64_t)c.w * L.cin * 2, (cuuint64_t)c.w * L.cin * 4,
                                     (cuuint64_t)c.h * c.w * L.cin * 2};
      const cuuint32_t box[5] = {64, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
      rc = encode_map(&tmA, c.x, 5, dims, strides, box);
    }
    if (rc) return rc;
  } else {
    mtiles = int((M + BM - 1) / BM);
  }
  // fused downsample branch (second A source): 1x1 conv of the block input onto the same output tile
  CUtensorMap tmA2;
  memset(&tmA2, 0, sizeof(tmA2));
  if (L.k2pad > 0) {
    FAV_REQUIRE(mode == 0 && c.x2, "conv: the fused downsample branch needs the TMA path and its input");
    FAV_REQUIRE((L.cin2 % 64) == 0 && (L.stride2 == 1 || (L.stride2 == 2 && (c.h2 % 2) == 0 && (c.w2 % 2) == 0)),
                "conv: unsupported downsample geometry");
    FAV_REQUIRE(conv_out_dim(c.h2, 1, L.stride2, 0) == a.OH && conv_out_dim(c.w2, 1, L.stride2, 0) == a.OW,
                "conv: downsample branch output does not match the main branch");
    a.kb2 = L.k2pad / BK; a.stride2 = L.stride2; a.Cin2 = L.cin2;
    int rc;
    if (L.stride2 == 1) {
      const cuuint64_t dims[4] = {(cuuint64_t)L.cin2, (cuuint64_t)c.w2, (cuuint64_t)c.h2, (cuuint64_t)c.p};
      const cuuint64_t strides[3] = {(cuuint64_t)L.cin2 * 2, (cuuint64_t)c.w2 * L.cin2 * 2, (cuuint64_t)c.h2 * c.w2 * L.cin2 * 2};
      const cuuint32_t box[4] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
      rc = encode_map(&tmA2, c.x2, 4, dims, strides, box);
    } else {
      const cuuint64_t dims[5] = {(cuuint64_t)L.cin2 * 2, (cuuint64_t)c.w2 / 2, 2, (cuuint64_t)c.h2 / 2, (cuuint64_t)c.p};
      const cuuint64_t strides[4] = {(cuuint64_t)L.cin2 * 4, (cuuint64_t)c.w2 * L.cin2 * 2, (cuuint64_t)c.w2 * L.cin2 * 4,
                                     (cuuint64_t)c.h2 * c.w2 * L.cin2 * 2};
      const cuuint32_t box[5] = {64, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
      rc = encode_map(&tmA2, c.x2, 5, dims, strides, box);
    }
    if (rc) return rc;
  }
  static const int env_ablate = [] { const char* e = getenv("FAV_CONV_ABLATE"); return e ? atoi(e) : 0; }();
  a.ablate = env_ablate;
  // Small launches (fewer tiles than SMs): one CTA per SM is limited by the 64 B/clk L2->SM ingress of its own SM, so narrower
  // N tiles -- more CTAs, fewer bytes per CTA and k-block (A 16 KB + W 4 / 8 KB instead of 16 KB) -- shorten the K loop.  The
  // per-element accumulation order does not depend on the N tiling, and whether the residual goes through the identity
  // MMAs (which changes the order of the bias and residual adds) is decided from layer properties only (layer_epi_bound),
  // so results are bit-identical whatever tile width a launch gets.
  static const int env_narrow = [] { const char* e = getenv("FAV_NARROW_N"); return e ? atoi(e) : 1; }();
  static const int env_rmma = [] { const char* e = getenv("FAV_RES_MMA"); return e ? atoi(e) : -1; }();   // -1 auto, 0 off, 1 all eligible
  const bool layer_epi_bound = (a.num_kb + a.kb2) <= 8 || a.rep > 1 || L.bn == 64;       // independent of the launch's row count
  const bool want_res_mma = env_rmma != 0 && (env_rmma > 0 || layer_epi_bound) && mode == 0 && c.res && !c.out_f32 &&
                            (L.cout % 64) == 0 && (L.bn % 64) == 0;
  const CUtensorMap* tmW_sel = reinterpret_cast<const CUtensorMap*>(L.tmap_w);
  if (env_narrow && mode == 0 && L.bn == 128 && c.force_mt == 0 && !conv_pair_applicable(L, a, 0)) {
    const long long tiles128 = (long long)mtiles * (L.cout_pad / 128);
    if (L.tmap32_ok && !want_res_mma && 4 * tiles128 <= ctx->num_sms) { a.BN = 32; tmW_sel = reinterpret_cast<const CUtensorMap*>(L.tmap_w32); }
    else if (L.tmap64_ok && 2 * tiles128 <= ctx->num_sms) { a.BN = 64; tmW_sel = reinterpret_cast<const CUtensorMap*>(L.tmap_w64); }
  }
  a.ntiles = L.cout_pad / a.BN;
  a.mtiles = mtiles;
  // M = 256 per CTA when there is enough work to fill the machine twice over with one CTA per SM
  const int pair_tiles = ((mtiles + 1) / 2) * a.ntiles;
  static const int env_mt = [] { const char* e = getenv("FAV_FORCE_MT"); return e ? atoi(e) : 0; }();   // tuning aid
  const int force_mt = c.force_mt ? c.force_mt : env_mt;
  (void)pair_tiles;   // measured on B200: two CTAs x 128-pixel tiles beat one CTA x 256-pixel tiles on every ResNet-18 layer
  const int MT = (mode == 0 && force_mt == 2) ? 2 : 1;
  a.mt_per_tile = MT;
  a.total_tiles = ((mtiles + MT - 1) / MT) * a.ntiles;
  // split-K (opt-in: FAV_SPLITK=1 or fav_set_option(h, "splitk", 1)).  MEASURED on the 640x480 batch-1 gate: correct and
  // deterministic, but each split launch is 1.5-2x SLOWER than the unsplit one (layer4 24-37 us -> 52-63 us, ncu launch
  // list), so nothing enables it by default.  Role timers (tools/splitk_dbg.py, 512->512 3x3 on 15x20 pixels): the 96 slice CTAs
  // only pull ~18 B/clk each from L2 (they hammer the same weight lines), so the MMA span shrinks 23 -> 9 us instead of 3 us,
  // and store + fence + ticket + fix-up add ~7 us per CTA.
  // a launch that fills less than half the machine but has long K loops (batch-1 layer3/4) is cut into ksplit slices per
  // output tile; every slice keeps at least one k-block (the centre tap is never skipped).  Each slice stores its fp32
  // partial tile; the last one to finish adds them in slice order, so results are deterministic.  Not used by the sweep:
  // whether a launch splits depends on its row count, and the sweep's results must not depend on the block size.
  static const int env_splitk = [] { const char* e = getenv("FAV_SPLITK"); return e ? atoi(e) : -1; }();
  const bool splitk_on = env_splitk >= 0 ? env_splitk != 0 : ctx->allow_splitk;
  a.ksplit = 0;
  const bool pair_ok = MT == 1 && conv_pair_applicable(L, a, c.force_mt == 3 ? 1 : 0) && c.force_mt != 1;
  if (splitk_on && mode == 0 && !stem_tma && MT == 1 && !pair_ok && !c.drop && a.rep == 1 && !c.out_f32 &&
      a.cin_blocks >= 2 && a.num_kb + a.kb2 >= 16 && 2 * a.total_tiles <= ctx->num_sms) {
    int S = (2 * ctx->num_sms) / a.total_tiles;
    if (S > a.cin_blocks) S = a.cin_blocks;
    if (S > 8) S = 8;
    if (S >= 2) {
      const size_t tick_bytes = ((size_t)a.total_tiles * 4 + 255) / 256 * 256;
      const size_t need = tick_bytes + (size_t)a.total_tiles * S * BM * a.BN * 4;
      if (need > ctx->splitk_bytes) {
        if (ctx->splitk_buf) FAV_CUDA_OK(cudaFree(ctx->splitk_buf));
        ctx->splitk_buf = nullptr; ctx->splitk_bytes = 0;
        FAV_CUDA_OK(cudaMalloc(&ctx->splitk_buf, need));
        FAV_CUDA_OK(cudaMemset(ctx->splitk_buf, 0, need));          // tickets start at zero; each launch leaves them at zero
        ctx->splitk_bytes = need;
      }
      a.ksplit = S;
      a.tickets = reinterpret_cast<int*>(ctx->splitk_buf);
      a.acc32 = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->splitk_buf) + tick_bytes);
      a.total_tiles *= S;
    }
  }
  // launch-time constants of the tile / row decode as multiply-high divisors
  a.fd_ntiles = make_fastdiv(uint32_t(a.ntiles));
  a.fd_ksplit = make_fastdiv(uint32_t(a.ksplit > 1 ? a.ksplit : 1));
  a.fd_ohw = make_fastdiv(uint32_t(a.OH * a.OW));
  a.fd_ow = make_fastdiv(uint32_t(a.OW));
  if (mode == 0) {
    a.fd_tw = make_fastdiv(uint32_t(a.tiles_w)); a.fd_th = make_fastdiv(uint32_t(a.tiles_h));
    a.fd_twth = make_fastdiv(uint32_t(a.tiles_w * a.tiles_h));
    a.fd_perimg = make_fastdiv(uint32_t(a.bw * a.bh)); a.fd_bw = make_fastdiv(uint32_t(a.bw));
  } else {
    a.fd_tw = a.fd_th = a.fd_twth = a.fd_perimg = a.fd_bw = make_fastdiv(1u);
  }
  {
    // exactness of the multiply-high divisions (n * d < 2^32), per division: tile / ksplit and tile / ntiles see n <= total_tiles,
    // m-tile / tiles_w and / tiles_h see n <= mtiles (a flattened 1x1 conv has tiles_w = mtiles in the tens of thousands)
    const unsigned long long nmax = (unsigned long long)a.total_tiles + 1, mmax = (unsigned long long)M + BM;
    const unsigned long long mtmax = (unsigned long long)mtiles + 2;
    unsigned long long d1 = (unsigned long long)a.ntiles, d2 = 1;
    if ((unsigned long long)a.ksplit > d1) d1 = a.ksplit;
    if (mode == 0) d2 = (unsigned long long)(a.tiles_w > a.tiles_h ? a.tiles_w : a.tiles_h);
    a.fastdiv = (nmax * d1 < (1ull << 32) && mtmax * d2 < (1ull << 32) &&
                 (mode == 0 || mmax * (unsigned long long)(a.OH * a.OW) < (1ull << 32))) ? 1 : 0;
  }
  const int stage_bytes = MT * A_TILE_BYTES + a.BN * 128;
  const int ctas_per_sm = (mode == 0 && MT == 1) ? 2 : 1;
  // staged epilogue (TMA stores from two shared-memory slabs): bf16 NHWC output of the TMA-tiled modes.  The default
  // keeps it for the layers whose epilogue is the bottleneck (few k-blocks per tile, or T masked replicas per tile).
  static const int env_stg = [] { const char* e = getenv("FAV_EPI_TMA"); return e ? atoi(e) : -1; }();   // -1 auto, 0 off, 1 all eligible
  const bool stg_ok = mode == 0 && !c.out_f32 && (L.cout % 64) == 0 && (a.BN % 64) == 0 && a.ksplit <= 1 && !a.drop2;
  const bool stg_auto = layer_epi_bound || a.BN == 64;       // (staged and direct epilogues compute identical values)
  a.stg_bytes = (stg_ok && env_stg != 0 && (env_stg > 0 || stg_auto)) ? 2 * STG_SLAB_BYTES : 0;
  CUtensorMap tmY;
  memset(&tmY, 0, sizeof(tmY));
  if (a.stg_bytes) {
    const cuuint64_t px = (cuuint64_t)L.cout * 2;
    const cuuint64_t dims[5] = {(cuuint64_t)L.cout, (cuuint64_t)a.OW, (cuuint64_t)a.OH, (cuuint64_t)a.rep, (cuuint64_t)c.p};
    const cuuint64_t strides[4] = {px, px * a.OW, px * a.OW * a.OH, px * a.OW * a.OH * a.rep};
    const cuuint32_t box[5] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, 1, (cuuint32_t)a.bn_img};
    int rc = encode_map(&tmY, c.y, 5, dims, strides, box);
    if (rc) return rc;
  }
  // residual added by the tensor core (identity MMAs): the residual tile arrives by TMA through the operand pipeline,
  // far ahead of its use, and the epilogue loses its residual loads, bf16 unpacking and adds
  // Default: only where the epilogue is the bottleneck (layer_epi_bound); on long-K layers the two extra k-blocks per tile
  // cost more shared-memory bandwidth than the epilogue saves.  FAV_RES_MMA=0 off, 1 everywhere.
  CUtensorMap tmR;
  memset(&tmR, 0, sizeof(tmR));
  a.res_mma = (want_res_mma && (a.BN % 64) == 0 && a.ksplit <= 1 && !pair_ok) ? 1 : 0;
  a.idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(64 >> 3) << 17) | (uint32_t(BM >> 4) << 24);
  if (a.res_mma) {
    a.res = nullptr;
    const cuuint64_t px = (cuuint64_t)L.cout * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)L.cout, (cuuint64_t)a.OW, (cuuint64_t)a.OH, (cuuint64_t)c.p};
    const cuuint64_t strides[3] = {px, px * a.OW, px * a.OW * a.OH};
    const cuuint32_t box[4] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bn_img};
    int rc = encode_map(&tmR, c.res, 4, dims, strides, box);
    if (rc) return rc;
  }
  // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2 = 113 KB each, minus 1 KB alignment slack and the barrier block
  int stages = ((ctas_per_sm == 2 ? 113 * 1024 - 1280 : 200 * 1024) - a.stg_bytes - (a.res_mma ? IDENT_BYTES : 0)) / stage_bytes;
  stages = stages < 2 ? 2 : (stages > 8 ? 8 : stages);
  a.stages = stages;
  uint32_t cols = 32;
  while (cols < uint32_t(2 * MT * a.BN)) cols *= 2;       // two accumulator sets (double-buffered epilogue)
  a.tmem_cols = cols;
  // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a.BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
  const size_t smem = (size_t)stages * stage_bytes + a.stg_bytes + (a.res_mma ? IDENT_BYTES : 0) + 1024 + 256;
  if (!ctx->attr_conv) {
    FAV_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FAV_CUDA_OK(cudaFuncSetAttribute(conv_igemm_m256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FAV_CUDA_OK(cudaFuncSetAttribute(conv_igemm_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->attr_conv = true;
  }
  const int threads = mode != 0 ? THREADS_GATHER : (MT == 2 ? THREADS_TMA2 : THREADS_TMA1);
  const int grid = a.total_tiles < ctas_per_sm * ctx->num_sms ? a.total_tiles : ctas_per_sm * ctx->num_sms;
  cudaEvent_t e1 = nullptr;
  {
    const float gf = L.fold ? float(2.0 * double(M) * 36.0 * (L.cin / 4) * (L.cout / 4) * 1e-9)   // nominal (stock PyTorch) FLOPs
                            : float(2.0 * double(M) * (L.r * L.s * L.cin + L.cin2) * L.cout * 1e-9);
    // algorithmic bytes: every operand read once (a strided 1x1 conv touches only the pixels it samples), result written once
    const double in1 = (L.r == 1 && L.s == 1 && L.stride > 1) ? double(M) * L.cin * 2 : double(c.p) * c.h * c.w * (L.cin_store ? L.cin_store : L.cin) * 2;
    const double in2 = L.k2pad > 0 ? double(M) * L.cin2 * 2 : 0.0;
    const double outb = double(M) * L.cout * (c.out_f32 ? 4 : 2) * (c.rep > 1 ? c.rep : 1);
    const double resb = c.res ? double(M) * L.cout * 2 : 0.0;
    const double wb = double(L.kpad + L.k2pad) * L.cout_pad * 2;
    int rc = conv_timing_begin(ctx, st, gf, float((in1 + in2 + outb + resb + wb) * 1e-9), &e1, &a.stats);
    if (rc) return rc;
  }
  if (pair_ok) {
    int rc = conv_pair_launch(ctx, L, a, tmA, tmA2, st);
    if (rc) return rc;
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const CUtensorMap& tmW = *tmW_sel;
    if (mode == 0 && MT == 2) FAV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_igemm_m256_kernel, tmA, tmW, tmA2, tmY, tmR, a));
    else if (mode == 0) FAV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_igemm_kernel, tmA, tmW, tmA2, tmY, tmR, a));
    else FAV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_igemm_gather_kernel, tmA, tmW, tmA2, tmY, tmR, a));
  }
  if (e1) FAV_CUDA_OK(cudaEventRecord(e1, st));
  ctx->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav

using namespace fav;

// Unit-test / tooling entry: one convolution with caller-provided (unpadded) device weights.
extern "C" int fav_conv2d(fav_handle h, const void* d_x, const void* d_w, const float* d_bias, const void* d_res,
                          void* d_y, int p, int height, int width, int cin, int cout, int r, int s, int stride,
                          int pad, int relu, int out_f32, int a_mode, void* stream) {
  FAV_REQUIRE(h && d_x && d_w && d_y, "fav_conv2d: null pointer");
  FAV_DEVICE(h);
  FAV_REQUIRE(cin > 0 && cout > 0 && r > 0 && s > 0 && stride > 0 && pad >= 0, "fav_conv2d: bad conv geometry");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ConvLayer L;
  L.cin = cin; L.cout = cout; L.r = r; L.s = s; L.stride = stride; L.pad = pad;
  conv_layer_finalize(L);                          // sizes only (w == nullptr)
  __nv_bfloat16* wpad = nullptr;
  float* bpad = nullptr;
  FAV_CUDA_OK(cudaMalloc(&wpad, (size_t)L.cout_pad * L.kpad * 2));
  FAV_CUDA_OK(cudaMalloc(&bpad, (size_t)L.cout_pad * 4));
  FAV_CUDA_OK(cudaMemsetAsync(wpad, 0, (size_t)L.cout_pad * L.kpad * 2, st));
  FAV_CUDA_OK(cudaMemsetAsync(bpad, 0, (size_t)L.cout_pad * 4, st));
  FAV_CUDA_OK(cudaMemcpy2DAsync(wpad, (size_t)L.kpad * 2, d_w, (size_t)L.k * 2, (size_t)L.k * 2, cout,
                                cudaMemcpyDeviceToDevice, st));
  if (d_bias) FAV_CUDA_OK(cudaMemcpyAsync(bpad, d_bias, (size_t)cout * 4, cudaMemcpyDeviceToDevice, st));
  L.w = wpad; L.bias = bpad;
  int rc = conv_layer_finalize(L);
  if (rc == FAV_OK) {
    ConvCall c;
    c.L = &L; c.x = d_x; c.y = d_y; c.res = d_res; c.p = p; c.h = height; c.w = width;
    c.relu = relu; c.out_f32 = out_f32;
    c.a_mode = a_mode < 0 ? -1 : (a_mode & 0xFF);          // bits 8-9: force the CTA tile height (1 -> 128, 2 -> 256 pixels)
    c.force_mt = a_mode < 0 ? 0 : ((a_mode >> 8) & 3);
    rc = conv_launch(h, c, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(wpad);
  cudaFree(bpad);
  if (rc == FAV_OK && e != cudaSuccess) { set_error("fav_conv2d: %s", cudaGetErrorString(e)); return FAV_E_CUDA; }
  return rc;
}
