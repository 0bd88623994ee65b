// trust.cu -- f4: batched replay of the reference's TrustEngine over many independent tick sequences.
//
// Replaces (reference): the per-tick Python loop of the batch replay, platform/backend/main.py:340-352, i.e.
// TrustEngine.update (platform/backend/trust_engine.py:139-243) with _update_policy (:68-87) and
// _update_contradiction_detector (:89-137).  Sequential in time, data-parallel across sequences: one thread per
// sequence, float64 throughout, every update written with explicit round-to-nearest intrinsics in the reference's
// operation order (no FMA contraction), so reliability / anomaly_integral / trust_velocity / recovery_debt /
// recovery_coeff are bit-identical to the Python engine (tests/golden/trust_replay.json, produced by the real class).
// Oracle twin: oracle/trust.py.  Layout: tick-major [L][S] so that the threads of a warp touch consecutive addresses.
#include <cmath>
#include "common.cuh"

namespace fav {

__device__ __forceinline__ int trust_policy(double rel, double vel) {
  if (rel >= 0.7 && vel < -0.15) return 1;      // VISION_DECLINING
  if (rel >= 0.7) return 0;                     // VISION_ALLOWED
  if (rel >= 0.3) return 2;                     // VISION_DEGRADED
  return 3;                                     // VISION_BLOCKED
}

constexpr int TRUST_BUF = 60;                   // rolling (status, score) window, trust_engine.py:57

__global__ void __launch_bounds__(128) k_trust_replay(const int8_t* __restrict__ status, const double* __restrict__ score,
                                                      const double* __restrict__ dts, double dt_const, int S, int L,
                                                      double* __restrict__ o_state, uint8_t* __restrict__ o_policy,
                                                      uint8_t* __restrict__ o_contra, int32_t* __restrict__ o_count,
                                                      double* __restrict__ o_final) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double rel = 1.0, integ = 0.0, vel = 0.0, debt = 0.0, coeff = 0.10, prev_rel = 1.0;
  int cur = -1, policy = 0, contra = 0, count = 0;
  double buf_sc[TRUST_BUF];
  int8_t buf_st[TRUST_BUF];
  int head = 0, fill = 0;                       // ring buffer: oldest entry at head
  for (int i = 0; i < L; ++i) {
    const size_t at = (size_t)i * S + s;
    int st = status[at];
    if (unsigned(st) > 3u) st = 3;               // unknown codes are treated as VISION_CORRUPTED (the host wrapper rejects them)
    const double sc = score[at];
    const bool has_sc = !isnan(sc);
    const double d = dts ? dts[i] : dt_const;
    if (cur < 0) {                              // first call (:153-158)
      cur = st;
      policy = trust_policy(rel, vel);
    } else if (st != cur) {                     // status change: timing reset only (:161-170)
      const int prev = cur;
      cur = st;
      if (st != 0 && prev == 0) integ = 0.0;
      policy = trust_policy(rel, vel);
    } else {
      if (st == 0) {                            // VISION_OK: recovery, debt drain, bounded ML penalty (:178-197)
        debt = fmax(0.0, __dsub_rn(debt, __dmul_rn(0.10, d)));
        coeff = fmax(0.03, __dsub_rn(0.10, __dmul_rn(0.008, debt)));
        rel = __dadd_rn(rel, __dmul_rn(coeff, d));
        if (has_sc) {
          integ = __dadd_rn(integ, __dmul_rn(sc, d));
          integ = __dsub_rn(integ, __dmul_rn(__dmul_rn(0.5, integ), d));
          integ = fmax(0.0, integ);
          rel = __dsub_rn(rel, __dmul_rn(__dmul_rn(0.15, integ), d));
        }
      } else {                                  // explicit failures dominate (:199-228)
        const double rate = st == 1 ? 0.30 : (st == 2 ? 0.60 : 1.00);
        const double debt_rate = fmax(0.0, __dsub_rn(0.7, rel));
        debt = fmin(10.0, __dadd_rn(debt, __dmul_rn(debt_rate, d)));
        rel = __dsub_rn(rel, __dmul_rn(rate, d));
        integ = 0.0;
      }
      rel = fmax(0.0, fmin(1.0, rel));
      const double raw_v = __ddiv_rn(__dsub_rn(rel, prev_rel), fmax(d, 0.001));
      vel = __dadd_rn(__dmul_rn(0.12, raw_v), __dmul_rn(1.0 - 0.12, vel));
      prev_rel = rel;
      // contradiction detector (:89-137): z-score of the current reading against the same-status scores of the window
      if (!has_sc) {
        contra = 0;
      } else {
        if (fill < TRUST_BUF) {
          const int pos = (head + fill) % TRUST_BUF;
          buf_sc[pos] = sc; buf_st[pos] = int8_t(st);
          ++fill;
        } else {
          buf_sc[head] = sc; buf_st[head] = int8_t(st);
          head = (head + 1) % TRUST_BUF;
        }
        int n_same = 0;
        double sum = 0.0, comp = 0.0;           // Neumaier-compensated: the reference sums exactly (statistics.mean)
        for (int k = 0; k < fill; ++k) {
          const int pos = (head + k) % TRUST_BUF;
          if (buf_st[pos] != st) continue;
          const double x = buf_sc[pos], t = __dadd_rn(sum, x);
          comp = __dadd_rn(comp, fabs(sum) >= fabs(x) ? __dadd_rn(__dsub_rn(sum, t), x) : __dadd_rn(__dsub_rn(x, t), sum));
          sum = t;
          ++n_same;
        }
        if (fill < 30 || n_same < 10) {
          contra = 0;
        } else {
          const double mean = __ddiv_rn(__dadd_rn(sum, comp), double(n_same));
          double ss = 0.0, c2 = 0.0;
          for (int k = 0; k < fill; ++k) {
            const int pos = (head + k) % TRUST_BUF;
            if (buf_st[pos] != st) continue;
            const double dv = __dsub_rn(buf_sc[pos], mean), x = __dmul_rn(dv, dv), t = __dadd_rn(ss, x);
            c2 = __dadd_rn(c2, fabs(ss) >= fabs(x) ? __dadd_rn(__dsub_rn(ss, t), x) : __dadd_rn(__dsub_rn(x, t), ss));
            ss = t;
          }
          const double var = __ddiv_rn(__dadd_rn(ss, c2), double(n_same - 1));
          const double sd = fmax(sqrt(var), 0.001);
          if (st == 0 && __ddiv_rn(__dsub_rn(sc, mean), sd) > 3.0) {
            if (!contra) ++count;
            contra = 1;
          } else {
            contra = 0;
          }
        }
      }
      policy = trust_policy(rel, vel);
    }
    if (o_state) {
      double* o = o_state + at * 5;
      o[0] = rel; o[1] = integ; o[2] = vel; o[3] = debt; o[4] = coeff;
    }
    if (o_policy) o_policy[at] = uint8_t(policy);
    if (o_contra) o_contra[at] = uint8_t(contra);
    if (o_count) o_count[at] = count;
  }
  if (o_final) {
    double* o = o_final + (size_t)s * 8;
    o[0] = rel; o[1] = integ; o[2] = vel; o[3] = debt; o[4] = coeff; o[5] = double(policy); o[6] = double(contra); o[7] = double(count);
  }
}

}  // namespace fav

using namespace fav;

extern "C" int fav_trust_replay(fav_handle h, const int8_t* d_status, const double* d_score, const double* d_dt, double dt_const,
                                int n_seq, int n_ticks, double* d_state, uint8_t* d_policy, uint8_t* d_contra, int32_t* d_count,
                                double* d_final, void* stream) {
  FAV_REQUIRE(h, "null handle");
  FAV_DEVICE(h);
  FAV_REQUIRE(n_seq >= 0 && n_ticks >= 0, "fav_trust_replay: bad shape %d x %d", n_seq, n_ticks);
  if (n_seq == 0 || n_ticks == 0) return FAV_OK;
  FAV_REQUIRE(d_status && d_score, "fav_trust_replay: null pointer");
  FAV_REQUIRE(d_dt || dt_const > 0.0, "fav_trust_replay: dt must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  k_trust_replay<<<(n_seq + 127) / 128, 128, 0, st>>>(d_status, d_score, d_dt, dt_const, n_seq, n_ticks, d_state, d_policy, d_contra,
                                                     d_count, d_final);
  h->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}
