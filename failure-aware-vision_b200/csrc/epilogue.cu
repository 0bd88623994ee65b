// epilogue.cu -- K3 uncertainty epilogue and K4 calibration / detection aggregates.
//
// Replaces (reference): nothing executable -- README.md:2 ("uncertainty estimation"), README.md:22-24
// (failure = incorrect prediction with high confidence; threshold left open -> tau is an argument).
// Nearest reference code: histogram entropy of a gray frame, platform/backend/signal_analyzer.py:100-112,
// and the summary counters of failure_attributor.py:93-108.  Definitions: oracle/uncertainty.py,
// oracle/metrics.py.
//
// K3: one warp per sample; lanes stride over classes; warp-shuffle reductions for max / sum / entropy;
// the pass-mean probabilities live in registers (C <= 1024) so logits are read from HBM exactly once.
// K4: per-CTA shared-memory histograms (ECE bins + AUROC buckets, u32) flushed with one 64-bit global
// atomic per non-empty slot; all accumulators are integers (counts, Q32 fixed-point sums).
#include "common.cuh"

namespace fav {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// softmax / entropy transcendentals on the SFU: ex2.approx and lg2.approx have <= 2 ulp relative error, which keeps
// confidence within 1e-6 and entropy / MI within 1e-5 of the fp32 oracle (tests/test_gpu_parity.py tolerances) at a
// fraction of the instruction count of expf / logf / IEEE division
__device__ __forceinline__ float fast_exp(float d) {          // exp(d), d <= 0 (exp(-inf) = 0)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float fast_log(float x) {          // ln(x), x > 0
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.6931471805599453f;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct SampleOut {
  float conf, H, mi;
  int pred;
};

// NC = ceil(C / 32) register slots per lane
template <int NC>
__device__ __forceinline__ SampleOut sample_uncertainty(const float* __restrict__ z, int T, int C, int lane) {
  float pbar[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) pbar[i] = 0.f;
  float hsum = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* zt = z + (size_t)t * C;
    float v[NC];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? zt[c] : -INFINITY;
      m = fmaxf(m, v[i]);
    }
    m = warp_max(m);
    float s = 0.f, w = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float d = v[i] - m;                      // -inf for padding lanes
      v[i] = fast_exp(d);                            // exp(-inf) = 0
      s += v[i];
      w += v[i] > 0.f ? v[i] * d : 0.f;
    }
    s = warp_sum(s);
    w = warp_sum(w);
    const float inv_s = fast_rcp(s);
#pragma unroll
    for (int i = 0; i < NC; ++i) pbar[i] += v[i] * inv_s;
    hsum += fast_log(s) - w * inv_s;                     // H(p_t) = ln S - sum_i e_i (z_i - m) / S   (no per-class log)
  }
  const float invT = 1.0f / float(T);
  float best = -1.f;
  int arg = 0;                                      // NaN logits never win a compare: pred stays a valid class
  float H = 0.f;
  const bool single = T == 1;                       // one pass: the mean IS that pass, H = H(p_1) is already in hsum, MI = 0
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    const float p = pbar[i] * invT;
    if (c < C) {
      if (p > best) { best = p; arg = c; }          // ascending c inside a lane: first max kept
      if (!single && p > 0.f) H -= p * fast_log(p);
    }
  }
  H = single ? hsum : warp_sum(H);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {                // argmax, lowest index on ties
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  SampleOut r;
  r.conf = best; r.pred = arg; r.H = H;
  r.mi = fmaxf(H - hsum * invT, 0.f);
  return r;
}

// T == 1 (deterministic MSP path, e.g. the ImageNet-shaped config C3): the mean IS the only pass, so nothing has to be kept
// per class -- conf = 1 / S (the largest exponential is exp(0) = 1), pred = argmax of the logits (lowest index on ties),
// H = ln S - sum e (z - m) / S, MI = 0.  Half the instructions and half the registers of the general path.
template <int NC>
__device__ __forceinline__ SampleOut sample_uncertainty_single(const float* __restrict__ z, int C, int lane) {
  float v[NC];
  float m = -INFINITY;
  int arg = lane < C ? lane : 0;                      // a valid class even when every logit is NaN / -inf
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? z[c] : -INFINITY;
    if (v[i] > m) { m = v[i]; arg = c; }              // ascending c inside a lane: first max kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {                  // max + argmax, lowest index on ties
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
  }
  float s = 0.f, w = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const float d = v[i] - m;
    const float e = fast_exp(d);
    s += e;
    w += e > 0.f ? e * d : 0.f;
  }
  s = warp_sum(s);
  w = warp_sum(w);
  const float inv_s = fast_rcp(s);
  SampleOut r;
  r.conf = inv_s; r.pred = arg; r.H = fast_log(s) - w * inv_s; r.mi = 0.f;
  return r;
}

// C <= 16: lanes run over the MC passes (lane t owns pass t, t + 32, ...), each lane does its softmax serially in
// registers; only the pass-mean needs cross-lane reductions.  ~3x fewer shuffles than lanes-over-classes at C = 10.
__device__ __forceinline__ SampleOut sample_uncertainty_small(const float* __restrict__ z, int T, int C, int lane) {
  float pbar[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) pbar[c] = 0.f;
  float hsum = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float* zt = z + (size_t)t * C;
    float v[16];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) { v[c] = c < C ? zt[c] : -INFINITY; m = fmaxf(m, v[c]); }
    float s = 0.f, w = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float d = v[c] - m;
      v[c] = fast_exp(d);
      s += v[c];
      w += v[c] > 0.f ? v[c] * d : 0.f;
    }
    const float inv_s = fast_rcp(s);
#pragma unroll
    for (int c = 0; c < 16; ++c) pbar[c] += v[c] * inv_s;
    hsum += fast_log(s) - w * inv_s;
  }
  const float invT = 1.0f / float(T);
  hsum = warp_sum(hsum);
  float best = -1.f, H = 0.f;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float p = warp_sum(pbar[c]) * invT;       // identical in every lane
    if (c < C) {
      if (p > best) { best = p; arg = c; }          // ascending c: lowest index wins ties
      if (p > 0.f) H -= p * fast_log(p);
    }
  }
  SampleOut r;
  r.conf = best; r.pred = arg; r.H = H;
  r.mi = fmaxf(H - hsum * invT, 0.f);
  return r;
}

struct HistGeom {
  int C, n_bins, n_buckets;
  float tau, inv_lnC;
};

__device__ __forceinline__ unsigned long long q32(float x) {
  return __float2ull_rn(x * 4294967296.0f);
}
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// shared histogram: [3*n_bins (count, correct, pad)] + [6*n_buckets]; Q32 sums go straight to global.
// s_binsum / s_cm (optional, small-C kernel): per-CTA copies of the per-bin confidence sums (u64) and of the C x C confusion
// matrix (u32); without them those two go straight to global atomics, which serialise on a handful of hot addresses
__device__ __forceinline__ void accumulate_sample(const HistGeom& g, unsigned* s_hist, unsigned long long* hist,
                                                  float conf, float H, float mi, int pred, int label,
                                                  unsigned long long* s_binsum = nullptr, unsigned* s_cm = nullptr) {
  const bool correct = pred == label;
  int b = int(ceilf(conf * float(g.n_bins))) - 1;
  b = min(max(b, 0), g.n_bins - 1);
  atomicAdd(&s_hist[2 * b], 1u);
  if (correct) atomicAdd(&s_hist[2 * b + 1], 1u);
  if (s_binsum) atomicAdd(&s_binsum[b], q32(conf));
  else atomicAdd(&hist[FAV_HIST_HDR + 3 * b + 1], q32(conf));
  const float s0 = clip01(1.0f - conf), s1 = clip01(H * g.inv_lnC), s2 = clip01(mi * g.inv_lnC);
  const float sc[3] = {s0, s1, s2};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int k = int(floorf(sc[s] * float(g.n_buckets)));
    k = min(max(k, 0), g.n_buckets - 1);
    atomicAdd(&s_hist[2 * g.n_bins + (s * g.n_buckets + k) * 2 + (correct ? 0 : 1)], 1u);
  }
  const size_t cb = FAV_HIST_HDR + 3 * (size_t)g.n_bins + 6 * (size_t)g.n_buckets;
  if (s_cm) {
    atomicAdd(&s_cm[label * g.C + pred], 1u);
  } else if (g.C <= 100) {
    atomicAdd(&hist[cb + (size_t)label * g.C + pred], 1ull);
  } else {
    atomicAdd(&hist[cb + 2 * (size_t)label], 1ull);
    if (correct) atomicAdd(&hist[cb + 2 * (size_t)label + 1], 1ull);
  }
}

// block-level header sums: n, correct, flag, sum conf / H / MI (Q32) reduced in shared then one atomic each
struct BlockSums {
  unsigned long long v[6];
};

template <int NC, bool FUSED, bool SINGLE = false>
__global__ void __launch_bounds__(256) k34_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                  int n, int T, HistGeom g, unsigned long long* __restrict__ hist,
                                                  float* __restrict__ o_conf, float* __restrict__ o_H,
                                                  float* __restrict__ o_mi, int32_t* __restrict__ o_pred,
                                                  uint8_t* __restrict__ o_flag,
                                                  // K4-only inputs (FUSED == false and logits == nullptr)
                                                  const float* __restrict__ i_conf, const float* __restrict__ i_H,
                                                  const float* __restrict__ i_mi, const int32_t* __restrict__ i_pred) {
  extern __shared__ unsigned s_hist[];
  __shared__ unsigned long long s_sums[6];
  const int n_slots = 2 * g.n_bins + 6 * g.n_buckets;
  const bool do_hist = hist != nullptr;
  if (do_hist) {
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < 6) s_sums[threadIdx.x] = 0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  unsigned long long my[6] = {0, 0, 0, 0, 0, 0};
  if (logits) {
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
      const SampleOut r = NC == 0 ? sample_uncertainty_small(logits + (size_t)i * T * g.C, T, g.C, lane)
                                  : (SINGLE ? sample_uncertainty_single<(NC > 0 ? NC : 1)>(logits + (size_t)i * g.C, g.C, lane)
                                            : sample_uncertainty<(NC > 0 ? NC : 1)>(logits + (size_t)i * T * g.C, T, g.C, lane));
      if (lane == 0) {
        const int label = labels ? labels[i] : -1;
        const bool flag = labels && r.pred != label && r.conf >= g.tau;
        if (o_conf) o_conf[i] = r.conf;
        if (o_H) o_H[i] = r.H;
        if (o_mi) o_mi[i] = r.mi;
        if (o_pred) o_pred[i] = r.pred;
        if (o_flag) o_flag[i] = flag ? 1 : 0;
        if (do_hist) {
          if (unsigned(label) >= unsigned(g.C)) {          // label outside [0, C): counted, kept out of every histogram
            atomicAdd(&hist[FAV_HIST_NINVALID], 1ull);
          } else {
            accumulate_sample(g, s_hist, hist, r.conf, r.H, r.mi, r.pred, label);
            my[0] += 1; my[1] += (r.pred == label); my[2] += flag;
            my[3] += q32(r.conf); my[4] += q32(clip01(r.H * g.inv_lnC)); my[5] += q32(clip01(r.mi * g.inv_lnC));
          }
        }
      }
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const float conf = i_conf[i], H = i_H[i], mi = i_mi[i];
      const int pred = i_pred[i], label = labels[i];
      if (unsigned(label) >= unsigned(g.C) || unsigned(pred) >= unsigned(g.C)) { atomicAdd(&hist[FAV_HIST_NINVALID], 1ull); continue; }
      accumulate_sample(g, s_hist, hist, conf, H, mi, pred, label);
      my[0] += 1; my[1] += (pred == label); my[2] += (pred != label && conf >= g.tau);
      my[3] += q32(conf); my[4] += q32(clip01(H * g.inv_lnC)); my[5] += q32(clip01(mi * g.inv_lnC));
    }
  }
  if (!do_hist) return;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (my[k]) atomicAdd(&s_sums[k], my[k]);
  __syncthreads();
  if (threadIdx.x < 6 && s_sums[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_sums[threadIdx.x]);
  for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
    const unsigned v = s_hist[i];
    if (!v) continue;
    size_t dst;
    if (i < 2 * g.n_bins) dst = FAV_HIST_HDR + 3 * (size_t)(i >> 1) + ((i & 1) ? 2 : 0);
    else dst = FAV_HIST_HDR + 3 * (size_t)g.n_bins + (size_t)(i - 2 * g.n_bins);
    atomicAdd(&hist[dst], (unsigned long long)v);
  }
}


// ---- C <= 16 (CIFAR-shaped sweeps): TPS threads per sample, 512 / TPS samples per tile ----
// A sample's T*C logits are one contiguous run and so is a whole tile: it is copied global -> shared with 16-byte
// cp.async (fully coalesced), double-buffered so the next tile streams in while this one is reduced.  Thread q of a
// sample softmaxes passes q, q + TPS, ... serially in registers (no shuffles, no idle class lanes); the pass-mean is a
// log2(TPS)-stage butterfly; thread 0 of the sample finishes entropy / argmax / flag and accumulates.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_le1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

constexpr int K34S_THREADS = 512;
constexpr int K34D_THREADS = 128;      // DIRECT variant: small launches, many small CTAs

// DIRECT (small launches, e.g. one 4096-sample sweep step): no per-CTA shared-memory histograms -- zeroing and flushing
// 98 KB of them per CTA costs more than the few thousand samples themselves; every sample's handful of counters goes
// straight to global atomics (spread over the 25 k slots of the arena row: no contention to speak of).
template <int TPS, int CT, bool DIRECT = false>   // CT > 0: C == CT exactly (no predicated class slots); CT == 0: any C <= 16
__global__ void __launch_bounds__(DIRECT ? K34D_THREADS : K34S_THREADS) k34_small_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                        int n, int T, HistGeom g, unsigned long long* __restrict__ hist,
                                                        int hist_bytes, float* __restrict__ o_conf, float* __restrict__ o_H,
                                                        float* __restrict__ o_mi, int32_t* __restrict__ o_pred,
                                                        uint8_t* __restrict__ o_flag) {
  constexpr int NTHR = DIRECT ? K34D_THREADS : K34S_THREADS;
  constexpr int S = NTHR / TPS, CM = CT > 0 ? CT : 16;
  extern __shared__ __align__(16) unsigned char k34_smem[];
  // [n_slots u32 | pad to 8 | n_bins u64 confidence sums | C*C u32 confusion | pad to 16] = hist_bytes, then the two tiles
  unsigned* s_hist = reinterpret_cast<unsigned*>(k34_smem);
  float* tiles = reinterpret_cast<float*>(k34_smem + hist_bytes);
  __shared__ unsigned long long s_sums[6];
  const int C = CT > 0 ? CT : g.C;
  const int n_slots = 2 * g.n_bins + 6 * g.n_buckets;
  unsigned long long* s_binsum = reinterpret_cast<unsigned long long*>(k34_smem + ((size_t(n_slots) * 4 + 7) & ~size_t(7)));
  unsigned* s_cm = reinterpret_cast<unsigned*>(s_binsum + g.n_bins);
  const bool do_hist = hist != nullptr;
  if (do_hist) {
    if (!DIRECT) {
      for (int i = threadIdx.x; i < n_slots; i += blockDim.x) s_hist[i] = 0;
      for (int i = threadIdx.x; i < g.n_bins; i += blockDim.x) s_binsum[i] = 0;
      for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cm[i] = 0;
    }
    if (threadIdx.x < 6) s_sums[threadIdx.x] = 0;
  }
  const size_t bin0 = FAV_HIST_HDR, bk0 = FAV_HIST_HDR + 3 * (size_t)g.n_bins, cm0 = bk0 + 6 * (size_t)g.n_buckets;
  const int row = T * C, tile_floats = S * row;
  const int ntiles = (n + S - 1) / S;
  auto issue = [&](int tile, int buf) {
    const size_t base = (size_t)tile * tile_floats;
    const int cnt = min(S, n - tile * S) * row;
    float* dst = tiles + (size_t)buf * tile_floats;
    for (int v = threadIdx.x; v < (cnt >> 2); v += NTHR) cp_async16(dst + 4 * v, logits + base + 4 * v);
    if (threadIdx.x < (cnt & 3)) dst[(cnt & ~3) + threadIdx.x] = logits[base + (cnt & ~3) + threadIdx.x];
  };
  int tile = blockIdx.x, buf = 0;
  if (tile < ntiles) issue(tile, 0);
  cp_async_commit();
  const int sidx = threadIdx.x / TPS, q = threadIdx.x % TPS;
  unsigned long long my[6] = {0, 0, 0, 0, 0, 0};
  for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait_le1();                 // this tile's group has landed (the prefetch may still be in flight)
    __syncthreads();
    const int i = tile * S + sidx;
    const float* z = tiles + (size_t)buf * tile_floats + (size_t)sidx * row;
    float pbar[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) pbar[c] = 0.f;
    float hsum = 0.f;
    if (i < n) {
      for (int t = q; t < T; t += TPS) {
        const float* zt = z + t * C;
        float v[CM];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CM; ++c) { v[c] = (CT > 0 || c < C) ? zt[c] : -INFINITY; m = fmaxf(m, v[c]); }
        float sum = 0.f, w = 0.f;
#pragma unroll
        for (int c = 0; c < CM; ++c) {
          const float d = v[c] - m;
          v[c] = fast_exp(d);
          sum += v[c];
          w += v[c] > 0.f ? v[c] * d : 0.f;
        }
        const float inv_s = fast_rcp(sum);
#pragma unroll
        for (int c = 0; c < CM; ++c) pbar[c] += v[c] * inv_s;
        hsum += fast_log(sum) - w * inv_s;
      }
    }
#pragma unroll
    for (int o = 1; o < TPS; o <<= 1) {          // butterfly over the sample's TPS lanes (aligned lane groups)
      hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
#pragma unroll
      for (int c = 0; c < CM; ++c) pbar[c] += __shfl_xor_sync(0xffffffffu, pbar[c], o);
    }
    if (i < n) {
      // every lane of the sample holds the full pass-sum: all finish redundantly (same warp time as one lane would
      // take), then lane q does its share of the accumulation: 0 -> outputs + ECE bin + header, 1..3 -> one AUROC score each
      const float invT = 1.0f / float(T);
      float best = -1.f, H = 0.f;
      int arg = 0;
#pragma unroll
      for (int c = 0; c < CM; ++c) {
        const float pc = pbar[c] * invT;
        if (CT > 0 || c < C) {
          if (pc > best) { best = pc; arg = c; }          // ascending c: lowest index wins ties
          if (pc > 0.f) H -= pc * fast_log(pc);
        }
      }
      const float mi = T == 1 ? 0.f : fmaxf(H - hsum * invT, 0.f);       // one pass: the mean is that pass, MI is exactly 0
      const int label = labels ? labels[i] : -1;
      const bool correct = arg == label;
      const bool flag = labels && !correct && best >= g.tau;
      if (q == 0) {
        if (o_conf) o_conf[i] = best;
        if (o_H) o_H[i] = H;
        if (o_mi) o_mi[i] = mi;
        if (o_pred) o_pred[i] = arg;
        if (o_flag) o_flag[i] = flag ? 1 : 0;
      }
      if (do_hist && unsigned(label) >= unsigned(C)) {     // label outside [0, C): counted, kept out of every histogram
        if (q == 0) atomicAdd(&hist[FAV_HIST_NINVALID], 1ull);
      } else if (do_hist) {
        const float sc0 = clip01(1.0f - best), sc1 = clip01(H * g.inv_lnC), sc2 = clip01(mi * g.inv_lnC);
        if (q == 0) {
          int b = int(ceilf(best * float(g.n_bins))) - 1;
          b = min(max(b, 0), g.n_bins - 1);
          if (DIRECT) {
            atomicAdd(&hist[bin0 + 3 * b], 1ull);
            if (correct) atomicAdd(&hist[bin0 + 3 * b + 2], 1ull);
            atomicAdd(&hist[bin0 + 3 * b + 1], q32(best));
            atomicAdd(&hist[cm0 + label * C + arg], 1ull);
          } else {
            atomicAdd(&s_hist[2 * b], 1u);
            if (correct) atomicAdd(&s_hist[2 * b + 1], 1u);
            atomicAdd(&s_binsum[b], q32(best));
            atomicAdd(&s_cm[label * C + arg], 1u);
          }
          my[0] += 1; my[1] += correct; my[2] += flag;
          my[3] += q32(best); my[4] += q32(sc1); my[5] += q32(sc2);
        }
        // AUROC buckets: with >= 4 lanes per sample lanes 1..3 take one score each, otherwise lane 0 does all three
#pragma unroll
        for (int si = 0; si < 3; ++si) {
          if (q != (TPS >= 4 ? si + 1 : 0)) continue;
          const float sc = si == 0 ? sc0 : (si == 1 ? sc1 : sc2);
          int k = int(floorf(sc * float(g.n_buckets)));
          k = min(max(k, 0), g.n_buckets - 1);
          if (DIRECT) atomicAdd(&hist[bk0 + (size_t)(si * g.n_buckets + k) * 2 + (correct ? 0 : 1)], 1ull);
          else atomicAdd(&s_hist[2 * g.n_bins + (si * g.n_buckets + k) * 2 + (correct ? 0 : 1)], 1u);
        }
      }
    }
    __syncthreads();                       // every reader is done with `buf` before the next iteration refills it
  }
  if (!do_hist) return;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (my[k]) atomicAdd(&s_sums[k], my[k]);
  __syncthreads();
  if (threadIdx.x < 6 && s_sums[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_sums[threadIdx.x]);
  if constexpr (!DIRECT) {
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
      const unsigned v = s_hist[i];
      if (!v) continue;
      size_t dst;
      if (i < 2 * g.n_bins) dst = FAV_HIST_HDR + 3 * (size_t)(i >> 1) + ((i & 1) ? 2 : 0);
      else dst = FAV_HIST_HDR + 3 * (size_t)g.n_bins + (size_t)(i - 2 * g.n_bins);
      atomicAdd(&hist[dst], (unsigned long long)v);
    }
    for (int i = threadIdx.x; i < g.n_bins; i += blockDim.x)
      if (s_binsum[i]) atomicAdd(&hist[FAV_HIST_HDR + 3 * (size_t)i + 1], s_binsum[i]);
    const size_t cb = FAV_HIST_HDR + 3 * (size_t)g.n_bins + 6 * (size_t)g.n_buckets;
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (s_cm[i]) atomicAdd(&hist[cb + i], (unsigned long long)s_cm[i]);
  }
}

}  // namespace fav

using namespace fav;

extern "C" size_t fav_hist_words(int C, int n_bins, int n_buckets) {
  const size_t conf = C <= 100 ? (size_t)C * C : 2 * (size_t)C;
  return FAV_HIST_HDR + 3 * (size_t)n_bins + 6 * (size_t)n_buckets + conf;
}

static int launch_k34(fav_handle h, const float* d_logits, const int32_t* d_labels, int n, int T, int C, float tau,
                      int n_bins, int n_buckets, int64_t* d_hist, float* d_conf, float* d_entropy, float* d_mi,
                      int32_t* d_pred, uint8_t* d_flag, const float* i_conf, const float* i_H, const float* i_mi,
                      const int32_t* i_pred, void* stream) {
  FAV_REQUIRE(h, "null handle");
  FAV_DEVICE(h);
  FAV_REQUIRE(n >= 0 && C >= 2 && C <= 1024, "C must be in [2,1024] (got %d), n >= 0 (got %d)", C, n);
  FAV_REQUIRE(!d_logits || T >= 1, "T must be >= 1 (got %d)", T);
  if (n == 0) return FAV_OK;
  HistGeom g;
  g.C = C; g.n_bins = n_bins; g.n_buckets = n_buckets; g.tau = tau; g.inv_lnC = float(1.0 / log(double(C)));
  size_t smem = 0;
  if (d_hist) {
    FAV_REQUIRE(d_labels, "histogram accumulation needs labels");
    FAV_REQUIRE(n_bins >= 1 && n_bins <= 1024 && n_buckets >= 1 && n_buckets <= 8192, "bad n_bins/n_buckets");
    smem = (2 * (size_t)n_bins + 6 * (size_t)n_buckets) * 4;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(d_hist);
  if (d_logits && d_hist && C == 10 && n <= 32768 && T >= 5 && (reinterpret_cast<uintptr_t>(d_logits) & 15) == 0 &&
      2 * (size_t)(K34D_THREADS / 8) * T * C * 4 <= 200 * 1024) {
    // small launch (a sweep step): 16 samples per 128-thread CTA, counters straight to global atomics
    const int S = K34D_THREADS / 8;
    const size_t sm = 2 * (size_t)S * T * C * 4;
    if (sm > 48 * 1024) FAV_CUDA_OK(cudaFuncSetAttribute(k34_small_kernel<8, 10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    long long nb = (n + S - 1) / S;
    const long long cap = (long long)h->num_sms * 8;
    if (nb > cap) nb = cap;
    k34_small_kernel<8, 10, true><<<int(nb), K34D_THREADS, sm, st>>>(d_logits, d_labels, n, T, g, hist, 0, d_conf, d_entropy, d_mi, d_pred, d_flag);
    h->launches++;
    FAV_CUDA_OK(cudaGetLastError());
    return FAV_OK;
  }
  if (d_logits && C <= 16 && (reinterpret_cast<uintptr_t>(d_logits) & 15) == 0) {
    // thread-group-per-sample path: the largest tile (256 / TPS samples) whose two buffers fit beside the histogram
    const size_t hist_bytes = d_hist ? ((((smem + 7) & ~size_t(7)) + (size_t)n_bins * 8 + (size_t)C * C * 4 + 15) & ~size_t(15)) : 0;
    const size_t budget = 220 * 1024;
    int tps = 0;
    int first = 1;                                     // no more lanes per sample than passes (T = 1: thread per sample), at most 8 by choice
    while (first < T && first < 8) first *= 2;
    for (int cand = first; cand <= 32; cand *= 2)
      if (hist_bytes + 2 * (size_t)(K34S_THREADS / cand) * T * C * 4 <= budget) { tps = cand; break; }
    if (tps) {
      const size_t sm = hist_bytes + 2 * (size_t)(K34S_THREADS / tps) * T * C * 4;
      const int S = K34S_THREADS / tps;
      long long nb = (n + S - 1) / S;
      if (nb > h->num_sms) nb = h->num_sms;
#define FAV_K34S(TPS, CT)                                                                                               \
  do {                                                                                                                  \
    FAV_CUDA_OK(cudaFuncSetAttribute(k34_small_kernel<TPS, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));  \
    k34_small_kernel<TPS, CT><<<int(nb), K34S_THREADS, sm, st>>>(d_logits, d_labels, n, T, g, hist, int(hist_bytes), d_conf,      \
                                                        d_entropy, d_mi, d_pred, d_flag);                               \
  } while (0)
      if (C == 10) {
        switch (tps) {
          case 1: FAV_K34S(1, 10); break;  case 2: FAV_K34S(2, 10); break;  case 4: FAV_K34S(4, 10); break;
          case 8: FAV_K34S(8, 10); break;  case 16: FAV_K34S(16, 10); break; default: FAV_K34S(32, 10); break;
        }
      } else {
        switch (tps) {
          case 1: FAV_K34S(1, 0); break;  case 2: FAV_K34S(2, 0); break;  case 4: FAV_K34S(4, 0); break;
          case 8: FAV_K34S(8, 0); break;  case 16: FAV_K34S(16, 0); break; default: FAV_K34S(32, 0); break;
        }
      }
#undef FAV_K34S
      h->launches++;
      FAV_CUDA_OK(cudaGetLastError());
      return FAV_OK;
    }
  }
  const int nc = C <= 16 ? 0 : (C <= 32 ? 1 : (C <= 128 ? 4 : 32));
  const int warps = 8;
  long long blocks = d_logits ? (n + warps - 1) / warps : (n + 255) / 256;
  const long long cap = (long long)h->num_sms * (smem > 64 * 1024 ? 2 : 4);
  if (blocks > cap) blocks = cap;
#define FAV_K34(NC)                                                                                           \
  do {                                                                                                        \
    if (smem > 48 * 1024)                                                                                     \
      FAV_CUDA_OK(cudaFuncSetAttribute(k34_kernel<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    k34_kernel<NC, true><<<int(blocks), 256, smem, st>>>(d_logits, d_labels, n, T, g, hist, d_conf, d_entropy, d_mi,  \
                                                         d_pred, d_flag, i_conf, i_H, i_mi, i_pred);          \
  } while (0)
#define FAV_K34_1(NC)                                                                                         \
  do {                                                                                                        \
    if (smem > 48 * 1024)                                                                                     \
      FAV_CUDA_OK(cudaFuncSetAttribute(k34_kernel<NC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    k34_kernel<NC, true, true><<<int(blocks), 256, smem, st>>>(d_logits, d_labels, n, T, g, hist, d_conf, d_entropy, d_mi,  \
                                                               d_pred, d_flag, i_conf, i_H, i_mi, i_pred);    \
  } while (0)
  if (d_logits && T == 1 && nc == 4) FAV_K34_1(4);
  else if (d_logits && T == 1 && nc == 32) FAV_K34_1(32);
  else if (nc == 0) FAV_K34(0); else if (nc == 1) FAV_K34(1); else if (nc == 4) FAV_K34(4); else FAV_K34(32);
#undef FAV_K34_1
#undef FAV_K34
  h->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}

extern "C" int fav_epilogue(fav_handle h, const float* d_logits, const int32_t* d_labels, int n, int T, int C, float tau,
                            float* d_conf, float* d_entropy, float* d_mi, int32_t* d_pred, uint8_t* d_flag,
                            void* stream) {
  if (n == 0 && h) return FAV_OK;
  FAV_REQUIRE(d_logits, "fav_epilogue: logits required");
  FAV_REQUIRE(!d_flag || d_labels, "fav_epilogue: failure flags need labels");
  return launch_k34(h, d_logits, d_labels, n, T, C, tau, 0, 0, nullptr, d_conf, d_entropy, d_mi, d_pred, d_flag,
                    nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int fav_accumulate(fav_handle h, const float* d_conf, const float* d_entropy, const float* d_mi,
                              const int32_t* d_pred, const int32_t* d_labels, int n, int C, float tau, int n_bins,
                              int n_buckets, int64_t* d_hist, void* stream) {
  if (n == 0 && h) return FAV_OK;
  FAV_REQUIRE(d_conf && d_entropy && d_mi && d_pred && d_labels && d_hist, "fav_accumulate: null pointer");
  return launch_k34(h, nullptr, d_labels, n, 1, C, tau, n_bins, n_buckets, d_hist, nullptr, nullptr, nullptr, nullptr,
                    nullptr, d_conf, d_entropy, d_mi, d_pred, stream);
}

extern "C" int fav_epilogue_accumulate(fav_handle h, const float* d_logits, const int32_t* d_labels, int n, int T,
                                       int C, float tau, int n_bins, int n_buckets, int64_t* d_hist, float* d_conf,
                                       float* d_entropy, float* d_mi, int32_t* d_pred, uint8_t* d_flag, void* stream) {
  if (n == 0 && h) return FAV_OK;
  FAV_REQUIRE(d_logits && d_labels && d_hist, "fav_epilogue_accumulate: null pointer");
  return launch_k34(h, d_logits, d_labels, n, T, C, tau, n_bins, n_buckets, d_hist, d_conf, d_entropy, d_mi, d_pred,
                    d_flag, nullptr, nullptr, nullptr, nullptr, stream);
}
