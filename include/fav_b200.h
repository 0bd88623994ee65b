/*
 * fav_b200.h -- C ABI of the B200-native corruption-sweep evaluation path.
 *
 * Drop-in boundary.  The reference (Indra-jith/failure-aware-vision) has no
 * FFI: it composes plain Python objects in platform/backend/main.py:110-118 and duck-types
 *     analyzer.analyze_frame(frame) -> dict            (main.py:160, signal_analyzer.py:47-143)
 *     anomaly.compute_anomaly(noise, brightness, st)   (main.py:141-143, anomaly_simulator.py:34-77)
 *     vision.set_noise/set_brightness/set_mode         (main.py:269-282, vision_simulator.py:25-36)
 * Each entry point below names the reference call whose work it takes over (or "none" where
 * the reference only states the intent, README.md:2,15-24).  The Python mirror of those
 * objects lives in failure-aware-vision_b200/ and calls this library through ctypes
 * (binding shown in INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer named d_* is a DEVICE pointer owned by the
 *     caller; the library owns only its weight arena + workspace inside fav_handle.
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns; no hidden
 *     synchronisation unless stated.
 *   - return 0 on success, negative FAV_E_* on error; message via fav_last_error()
 *     (thread-local).  There is no CPU fallback: without a usable sm_100 device fav_init fails.
 *   - images are uint8 NHWC [n,h,w,3]; activations are bf16 NHWC; logits are fp32 [n,T,C].
 */
#ifndef FAV_B200_H
#define FAV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAV_ABI_VERSION 2

#define FAV_OK 0
#define FAV_E_ARG (-1)      /* bad argument */
#define FAV_E_CUDA (-2)     /* CUDA runtime / driver error */
#define FAV_E_DEVICE (-3)   /* not an sm_100 device */
#define FAV_E_STATE (-4)    /* call order (e.g. forward before load_weights) */
#define FAV_E_UNSUPPORTED (-5)

typedef struct fav_ctx* fav_handle;

/* ---- corruption ids (order of Hendrycks & Dietterich's 15; 0 = clean) --------------------- */
enum {
  FAV_CLEAN = 0, FAV_GAUSSIAN_NOISE = 1, FAV_SHOT_NOISE = 2, FAV_IMPULSE_NOISE = 3,
  FAV_DEFOCUS_BLUR = 4, FAV_GLASS_BLUR = 5, FAV_MOTION_BLUR = 6, FAV_ZOOM_BLUR = 7,
  FAV_SNOW = 8, FAV_FROST = 9, FAV_FOG = 10, FAV_BRIGHTNESS = 11, FAV_CONTRAST = 12,
  FAV_ELASTIC = 13, FAV_PIXELATE = 14, FAV_JPEG = 15
};

/* flags for fav_corrupt_normalize */
#define FAV_SRC_BGR 1u        /* source frames are BGR (reference frames: signal_analyzer.py:51,62) */
#define FAV_OUT_F32 2u        /* write fp32 instead of bf16 (parity tooling) */
#define FAV_NO_NORMALIZE 4u   /* write the corrupted [0,1] value, skip (x-mean)/std */

/* model ids for fav_load_weights */
#define FAV_RESNET18 18
#define FAV_RESNET50 50

/* ---- lifetime ------------------------------------------------------------------------------ */
int fav_abi_version(void);
const char* fav_last_error(void);
/* replaces: object construction at main.py:110-118 (one handle per connection / per process). */
int fav_init(int device, fav_handle* out);
/* replaces: reset() at signal_analyzer.py:41-45 / main.py:285-290 (drops workspace, keeps weights). */
int fav_reset(fav_handle h);
int fav_destroy(fav_handle h);

/* ---- K1: corrupt + normalize ---------------------------------------------------------------
 * replaces: vision_simulator.py:25-36 (set_mode / set_noise / set_brightness knobs; here the 15 x 5
 * grid of Hendrycks & Dietterich) and the display-only JS effects (frontend/js/app.js:789-799,
 * 834-851); the reference has no server-side generator.
 *
 * d_src  uint8 [n,h,w,3]; d_dst bf16 (or fp32 with FAV_OUT_F32) [n,h,w,3], RGB order.
 * Self-sufficient (the signature of SURVEY.md 8b): the library derives every per-corruption constant
 * and table (Poisson inverse-CDF thresholds, stencil taps, resampling ranges, libjpeg quantisation
 * tables, Pillow BOX coefficients, elastic smoothing matrices) from (corruption, severity, h, w,
 * profile) on the host, caches it in the handle and owns the device scratch of the multi-pass
 * corruptions.  The first call for a new (corruption, severity, h, w) builds and uploads its table
 * (host work + a synchronous copy); every later call only enqueues kernels on `stream`.
 * Profile: CIFAR-10-C constants for frames up to 64 px, ImageNet-C above, or forced by flags.
 * All randomness is Philox4x32-10 keyed by (seed, first_image + image index, position):
 *   results do not depend on batch size or GPU count. */
#define FAV_PROFILE_CIFAR 0x10u
#define FAV_PROFILE_IMAGENET 0x20u
int fav_corrupt_normalize(fav_handle h, const uint8_t* d_src, void* d_dst, int n, int height,
                          int width, int corruption, int severity, uint64_t seed,
                          uint64_t first_image, const float mean[3], const float std[3],
                          unsigned flags, void* stream);
/* Table-taking form (tests / tooling): the caller supplies the constants (fparams / iparams), the
 * device table and the device scratch; fav_corrupt_params produces them on the host. */
int fav_corrupt_normalize_ex(fav_handle h, const uint8_t* d_src, void* d_dst, int n, int height,
                             int width, int corruption, int severity, const float* fparams,
                             int n_fparams, const int32_t* iparams, int n_iparams,
                             const void* d_table, size_t table_bytes, void* d_scratch,
                             size_t scratch_bytes, uint64_t seed, uint64_t first_image,
                             const float mean[3], const float std[3], unsigned flags, void* stream);
/* bytes of d_scratch fav_corrupt_normalize_ex needs for (corruption, n, h, w). */
size_t fav_corrupt_scratch_bytes(int corruption, int n, int height, int width);
/* Host only (no device needed): the constants and the table image of one cell.  In: capacities in
 * *n_fparams, *n_iparams, *table_bytes; out: the sizes.  Returns 1 when the buffers were filled, 0 when
 * only the sizes were reported (call again with room), negative on error.  flags: FAV_PROFILE_*. */
int fav_corrupt_params(int corruption, int severity, int height, int width, unsigned flags,
                       float* fparams, int* n_fparams, int32_t* iparams, int* n_iparams, void* table,
                       size_t* table_bytes);
/* Host only: the severity constants of (profile 0 CIFAR-10-C / 1 ImageNet-C, corruption, severity)
 * -> out[0..count); returns count or a negative error. */
int fav_corruption_constants(int profile, int corruption, int severity, double* out, int cap);

/* ---- K2: classifier forward ---------------------------------------------------------------
 * replaces: nothing in the reference ("image classification", README.md:19; torchvision named in
 * requirements.txt:2); fills the ML-score slot of anomaly_simulator.py:34-77.
 *
 * blob: host memory, format FAVW1 written by failure-aware-vision_b200/weights.py (BN already
 * folded, weights bf16 [Cout][R][S][Cin], bias fp32).  Copies to the device arena; synchronous. */
int fav_load_weights(fav_handle h, const void* blob, size_t nbytes, int model_id, int num_classes,
                     int in_h, int in_w);
/* max images per fav_forward_mc call for a given T (sizes the workspace; allocates). */
int fav_reserve(fav_handle h, int max_images, int T);
/* d_x bf16 [n,h,w,3] -> d_logits fp32 [n,T,C].  T MC-dropout passes in one batched launch
 * sequence; masks are generated in the conv epilogues from Philox(seed, first_image+i, t, layer), one byte per
 * activation: p_drop is quantised to round(256 p) / 256 and kept values are scaled by the exact inverse of the realised
 * keep probability (oracle/model.py states the contract).  T == 1 disables dropout. */
int fav_forward_mc(fav_handle h, const void* d_x, float* d_logits, int n, int T, float p_drop,
                   uint64_t seed, uint64_t first_image, void* stream);
/* single convolution (unit-test / tooling entry): y = act(conv(x, w) + bias [+ res]).
 * x bf16 NHWC [p,h,w,cin]; w bf16 [cout][r][s][cin]; y bf16 (or fp32 if out_f32) NHWC.
 * a_mode: -1 auto; low byte 0 TMA / 1 vector gather / 2 scalar gather; bits 8-9 force 128- (1) or 256-pixel (2) CTA tiles. */
int fav_conv2d(fav_handle h, const void* d_x, const void* d_w, const float* d_bias,
               const void* d_res, void* d_y, int p, int height, int width, int cin, int cout,
               int r, int s, int stride, int pad, int relu, int out_f32, int a_mode,
               void* stream);

/* ---- K3: uncertainty epilogue ---------------------------------------------------------------
 * replaces: nothing (README.md:2 "uncertainty estimation"; failure definition README.md:22-24).
 * d_logits fp32 [n,T,C]; d_labels int32 [n] or NULL; outputs may individually be NULL. */
int fav_epilogue(fav_handle h, const float* d_logits, const int32_t* d_labels, int n, int T, int C,
                 float tau, float* d_conf, float* d_entropy, float* d_mi, int32_t* d_pred,
                 uint8_t* d_flag, void* stream);

/* ---- K4: calibration / detection aggregates --------------------------------------------------
 * replaces: nothing (nearest: failure_attributor.py:93-108 summary counters).
 * Arena of int64 words, layout below; counts and Q32 fixed-point sums only, so a plain
 * integer sum over ranks (one NCCL all-reduce) gives bit-identical metrics on 1/2/4/8 GPUs. */
#define FAV_HIST_HDR 8
#define FAV_HIST_N 0
#define FAV_HIST_NCORRECT 1
#define FAV_HIST_NFLAG 2
#define FAV_HIST_SUM_CONF 3
#define FAV_HIST_SUM_H 4
#define FAV_HIST_SUM_MI 5
#define FAV_HIST_NINVALID 6   /* samples whose label (or supplied prediction) is outside [0, C): counted here, excluded elsewhere */
size_t fav_hist_words(int C, int n_bins, int n_buckets);
int fav_accumulate(fav_handle h, const float* d_conf, const float* d_entropy, const float* d_mi,
                   const int32_t* d_pred, const int32_t* d_labels, int n, int C, float tau,
                   int n_bins, int n_buckets, int64_t* d_hist, void* stream);
/* fused K3+K4 (what the sweep uses): logits -> arena, per-sample outputs optional. */
int fav_epilogue_accumulate(fav_handle h, const float* d_logits, const int32_t* d_labels, int n,
                            int T, int C, float tau, int n_bins, int n_buckets, int64_t* d_hist,
                            float* d_conf, float* d_entropy, float* d_mi, int32_t* d_pred,
                            uint8_t* d_flag, void* stream);

/* ---- synthetic inputs on the device (bench / tests; Philox streams "images", "labels") ------ */
int fav_synth_images(fav_handle h, uint8_t* d_dst, int n, int height, int width, uint64_t seed,
                     uint64_t first_image, void* stream);
int fav_synth_labels(fav_handle h, int32_t* d_dst, int n, int C, uint64_t seed,
                     uint64_t first_image, void* stream);

/* ---- f1: fused SignalAnalyzer statistics -----------------------------------------------------
 * replaces: SignalAnalyzer.analyze_frame arithmetic, signal_analyzer.py:62-105
 * (cvtColor BGR2GRAY, Laplacian(CV_64F) sum / sum of squares, mean, absdiff vs previous gray,
 * 256-bin histogram).  d_frame BGR u8 [h,w,3]; d_prev_gray u8 [h,w] (in: previous, out: current;
 * first_frame != 0 skips the diff); d_out int64[4 + 256] = {sum_lap, sum_lap_sq, sum_gray,
 * sum_absdiff, hist[256]} -- all integers, bit-exact vs OpenCV. */
int fav_frame_stats(fav_handle h, const uint8_t* d_frame, uint8_t* d_prev_gray, int height,
                    int width, int first_frame, int64_t* d_out, void* stream);

/* ---- f4: batched TrustEngine replay -------------------------------------------------------------
 * replaces: the per-tick Python loop of the batch replay, platform/backend/main.py:340-352, i.e.
 * TrustEngine.update (trust_engine.py:139-243) incl. _update_policy (:68-87) and the contradiction
 * detector (:89-137), for n_seq independent sequences of n_ticks ticks.  Tick-major device arrays:
 * d_status int8 [n_ticks][n_seq] (0 OK, 1 FROZEN, 2 BLANK, 3 CORRUPTED), d_score f64 (NaN = None),
 * d_dt f64 [n_ticks] or null (then dt_const).  Outputs (each may be null): d_state f64
 * [n_ticks][n_seq][5] = {reliability, anomaly_integral, trust_velocity, recovery_debt,
 * recovery_coeff} unrounded, d_policy u8 (0 ALLOWED, 1 DECLINING, 2 DEGRADED, 3 BLOCKED), d_contra u8,
 * d_count i32, d_final f64 [n_seq][8] (the five floats, policy, contradiction, count after the last
 * tick).  float64, bit-identical to the Python engine (tests/golden/trust_replay.json). */
int fav_trust_replay(fav_handle h, const int8_t* d_status, const double* d_score, const double* d_dt,
                     double dt_const, int n_seq, int n_ticks, double* d_state, uint8_t* d_policy,
                     uint8_t* d_contra, int32_t* d_count, double* d_final, void* stream);

/* ---- multi-GPU: the path's one exchange -------------------------------------------
 * One process per GPU; work items are sharded with no data-path collective; at sweep end one integer all-reduce makes
 * the per-cell histogram arenas global (bit-identical metrics on 1/2/4/8 GPUs).  replaces: nothing -- the reference
 * is single-process (main.py:109-118).  fav_comm_unique_id: rank 0 creates the 128-byte NCCL id and the host language
 * broadcasts it (torch.distributed / MPI / a file); fav_comm_init: every rank joins; fav_allreduce: in-place
 * ncclAllReduce(sum, int64) of `count` words on `stream` (a no-op on a single rank). */
int fav_comm_unique_id(void* out128);
int fav_comm_init(fav_handle h, const void* id128, int rank, int world_size);
int fav_allreduce(fav_handle h, int64_t* d_hist, size_t count, void* stream);

/* per-handle switches.  "splitk" = 1: convolutions that fill less than half the GPU split their K loop over several CTAs
 * with a deterministic in-kernel fix-up.  Experimental and off by default: the sweep's results must not depend on how
 * many rows a launch has, and on the batch-1 gate (main.py:160) the split launches measured slower than the unsplit ones.
 * "k1_legacy" = 1: the round-1 corruption kernels; "k1_list_stencil" = 1: defocus_blur through the tap-list loop instead of
 * the register-tiled dense loop.  Both are A/B switches for measurements and tests: the results are bit-identical. */
int fav_set_option(fav_handle h, const char* name, int value);

/* counters for bench.py's gpu_launches claim */
uint64_t fav_launch_count(fav_handle h);
/* in-situ timing of the tensor-core conv launches (bench.py roofline): while enabled, every conv launch is
 * bracketed by CUDA events on its stream; fav_conv_timing_read synchronises, returns the summed device time
 * and the number of launches since the last read, and clears the record. */
int fav_conv_timing_enable(fav_handle h, int on);
int fav_conv_timing_read(fav_handle h, float* total_ms, int* n_launches);
/* same, but per launch in launch order: ms[i] and the launch's algorithmic GFLOP (2*M*K*N, padded taps counted). */
int fav_conv_timing_read_all(fav_handle h, float* ms, float* gflop, int cap, int* n_launches);
/* per-launch algorithmic GB (each operand read once, result written once) of the launches recorded since the last
 * read -- the HBM side of the per-layer roofline; call before fav_conv_timing_read* (which clear the record). */
int fav_conv_timing_read_bytes(fav_handle h, float* gbyte, int cap, int* n_launches);
/* tuning aid: per-launch role wait counters (cycles summed over CTAs, 8 per launch; layout in api.cu). */
int fav_conv_stats_read(fav_handle h, uint64_t* out, int cap_launches, int* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* FAV_B200_H */
