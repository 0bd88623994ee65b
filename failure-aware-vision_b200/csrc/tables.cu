// tables.cu -- host side of K1: the 15 x 5 severity constants and every table the corruption kernels consume
// (Poisson inverse-CDF thresholds, stencil tap lists, resampling ranges, libjpeg quantisation tables, Pillow BOX
// coefficients, folded elastic smoothing matrices), built in C++ so that fav_corrupt_normalize is callable from C alone.
//
// Replaces (reference): the knobs of platform/backend/vision_simulator.py:25-36 (set_mode / set_noise / set_brightness);
// the reference has no severity grid.  Definitions: SURVEY.md Appendix A.2 (Hendrycks & Dietterich); the CPU oracle
// (oracle/corruptions.py, oracle/jpeg.py) states the same definitions independently in numpy / scipy / cv2 / PIL and the
// tests compare the two (tests/test_host.py: tables; tests/test_gpu_parity.py: pixels).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>

#include "common.cuh"
#include "tables.h"

namespace fav {

// ---------------------------------------------------------------------------------------------- severity constants
// [profile][corruption id - 1][severity - 1][up to 7 values]; profile 0 = CIFAR-10-C, 1 = ImageNet-C
static const double SEV[2][15][5][7] = {
    {   // cifar
        {{.04}, {.06}, {.08}, {.09}, {.10}},                                                     // gaussian_noise
        {{500}, {250}, {100}, {75}, {50}},                                                       // shot_noise
        {{.01}, {.02}, {.03}, {.05}, {.07}},                                                     // impulse_noise
        {{.3, .4}, {.4, .5}, {.5, .6}, {1, .2}, {1.5, .1}},                                      // defocus_blur
        {{.05, 1, 1}, {.25, 1, 1}, {.4, 1, 1}, {.25, 1, 2}, {.4, 1, 2}},                         // glass_blur
        {{10, 1}, {10, 1.5}, {10, 2}, {10, 2.5}, {12, 3}},                                       // motion_blur
        {{1.06, .01}, {1.11, .01}, {1.16, .01}, {1.21, .01}, {1.26, .01}},                       // zoom_blur
        {{.1, .2, 1, .6, 8, 3, .95}, {.1, .2, 1, .5, 10, 4, .9}, {.15, .3, 1.75, .55, 10, 4, .9},
         {.25, .3, 2.25, .6, 12, 6, .85}, {.3, .3, 1.25, .65, 14, 12, .8}},                      // snow
        {{1, .2}, {1, .3}, {.9, .4}, {.85, .4}, {.75, .45}},                                     // frost
        {{.2, 3}, {.5, 3}, {.75, 2.5}, {1, 2}, {1.5, 1.75}},                                     // fog
        {{.05}, {.1}, {.15}, {.2}, {.3}},                                                        // brightness
        {{.75}, {.5}, {.4}, {.3}, {.15}},                                                        // contrast
        {{0, 0, .08}, {.05, .2, .07}, {.08, .06, .06}, {.1, .04, .05}, {.1, .03, .03}},          // elastic_transform
        {{.95}, {.9}, {.85}, {.75}, {.65}},                                                      // pixelate
        {{80}, {65}, {58}, {50}, {40}},                                                          // jpeg_compression
    },
    {   // imagenet
        {{.08}, {.12}, {.18}, {.26}, {.38}},
        {{60}, {25}, {12}, {5}, {3}},
        {{.03}, {.06}, {.09}, {.17}, {.27}},
        {{3, .1}, {4, .5}, {6, .5}, {8, .5}, {10, .5}},
        {{.7, 1, 2}, {.9, 2, 1}, {1, 2, 3}, {1.1, 3, 2}, {1.5, 4, 2}},
        {{10, 3}, {15, 5}, {15, 8}, {15, 12}, {20, 15}},
        {{1.11, .01}, {1.16, .01}, {1.21, .02}, {1.26, .02}, {1.33, .03}},
        {{.1, .3, 3, .5, 10, 4, .8}, {.2, .3, 2, .5, 12, 4, .7}, {.55, .3, 4, .9, 12, 8, .7},
         {.55, .3, 4.5, .85, 12, 8, .65}, {.55, .3, 2.5, .85, 12, 12, .55}},
        {{1, .4}, {.8, .6}, {.7, .7}, {.65, .7}, {.6, .75}},
        {{1.5, 2}, {2., 2}, {2.5, 1.7}, {2.5, 1.5}, {3., 1.4}},
        {{.1}, {.2}, {.3}, {.4}, {.5}},
        {{.4}, {.3}, {.2}, {.1}, {.05}},
        {{2., .7, .1}, {2., .08, .2}, {.05, .01, .02}, {.07, .01, .02}, {.12, .01, .02}},
        {{.6}, {.5}, {.4}, {.3}, {.25}},
        {{25}, {18}, {15}, {10}, {7}},
    },
};
static const int SEV_COUNT[15] = {1, 1, 1, 2, 3, 2, 2, 7, 2, 2, 1, 1, 3, 1, 1};

int profile_of(unsigned flags, int height, int width) {
  if (flags & FAV_PROFILE_CIFAR) return 0;
  if (flags & FAV_PROFILE_IMAGENET) return 1;
  return std::max(height, width) <= 64 ? 0 : 1;          // CIFAR-10-C constants for frames up to 64 px, ImageNet-C above
}

// ---------------------------------------------------------------------------------------------- small helpers
template <typename T>
static void append(std::vector<uint8_t>& buf, const T* p, size_t count) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
  buf.insert(buf.end(), b, b + count * sizeof(T));
}

// shot noise: kmin int32[256] | thr uint32[256][width] | jump uint16[256][256]
// thr[v][j] = floor(CDF_Poisson(v c / 255)(kmin[v] + j) * 2^32) clipped to 2^32 - 1; a 32-bit draw u maps to
// k = kmin[v] + #{j : thr[v][j] <= u}; jump[v][b] = #{j : thr[v][j] < b << 24} (where the probe for top byte b starts)
static void poisson_table(double c, int* width_out, size_t* guide_off, std::vector<uint8_t>& tab) {
  double lam[256];
  long long kmin[256];
  int width = 0;
  long long kmax_all = 0;
  for (int v = 0; v < 256; ++v) {
    lam[v] = double(v) * (c / 255.0);
    const double sd = std::sqrt(lam[v]);
    kmin[v] = std::max(0.0, std::floor(lam[v] - 7.5 * sd - 4));
    width = std::max(width, int(std::ceil(lam[v] + 7.5 * sd + 12) - double(kmin[v])));
    kmax_all = std::max(kmax_all, kmin[v]);
  }
  width = (width + 1 + 3) / 4 * 4;
  std::vector<double> lg(size_t(kmax_all) + width + 2);                         // ln k!
  double acc = 0.0;
  for (size_t k = 0; k < lg.size(); ++k) { acc += std::log(std::max(double(k), 1.0)); lg[k] = acc; }
  std::vector<int32_t> kmin32(256);
  std::vector<uint32_t> thr(size_t(256) * width);
  std::vector<uint16_t> jump(size_t(256) * 256);
  for (int v = 0; v < 256; ++v) {
    kmin32[v] = int32_t(kmin[v]);
    uint32_t* row = thr.data() + size_t(v) * width;
    if (lam[v] == 0.0) {
      for (int j = 0; j < width; ++j) row[j] = 4294967295u;
    } else {
      const double ll = std::log(lam[v]);
      double below = 0.0;                                                      // mass under the window (< 1e-13)
      for (long long k = 0; k < kmin[v]; ++k) below += std::exp(-lam[v] + double(k) * ll - lg[size_t(k)]);
      double cdf = 0.0;
      for (int j = 0; j < width; ++j) {
        const long long k = kmin[v] + j;
        cdf += std::exp(-lam[v] + double(k) * ll - lg[size_t(k)]);
        const double t = std::floor(std::min(cdf + below, 1.0) * 4294967296.0);
        row[j] = t >= 4294967295.0 ? 4294967295u : uint32_t(t);
      }
    }
    for (int b = 0; b < 256; ++b) {
      const uint64_t edge = uint64_t(b) << 24;
      int cnt = 0;
      while (cnt < width && uint64_t(row[cnt]) < edge) ++cnt;
      jump[size_t(v) * 256 + b] = uint16_t(cnt);
    }
  }
  append(tab, kmin32.data(), 256);
  append(tab, thr.data(), thr.size());
  append(tab, jump.data(), jump.size());
  *width_out = width;
  // guide[v][b] (k1_shot_smem), 10 + 6 bits.  b >= 1: k of the draw b << 24 (= kmin + jump[b]) and the number of thresholds
  // inside the cell [b << 24, (b + 1) << 24), saturated at 63 -- the kernel counts UP from the cell's start.  b == 0 (the
  // lower tail: dozens of tiny thresholds, almost all of them below any given draw): k of the cell's END (= kmin + jump[1])
  // and the number of non-zero thresholds inside -- the kernel counts DOWN from the top, a couple of steps on average
  // instead of walking the whole tail.  Only when every k fits 10 bits.
  *guide_off = 0;
  bool fits = true;
  std::vector<uint16_t> guide(size_t(256) * 256);
  for (int v = 0; v < 256 && fits; ++v)
    for (int b = 0; b < 256; ++b) {
      const int j0 = jump[size_t(v) * 256 + b];
      const int j1 = b < 255 ? int(jump[size_t(v) * 256 + b + 1]) : width;
      long long k = kmin[v] + j0;
      int cnt = j1 - j0;
      if (b == 0) {
        int nz = 0;
        while (nz < j1 && thr[size_t(v) * width + nz] == 0u) ++nz;
        k = kmin[v] + j1;
        cnt = j1 - nz;
      }
      if (k > 1023) { fits = false; break; }
      guide[size_t(v) * 256 + b] = uint16_t(k) | uint16_t(std::min(cnt, 63) << 10);
    }
  if (fits) {
    while (tab.size() % 16) tab.push_back(0);
    *guide_off = tab.size();
    append(tab, guide.data(), guide.size());
  }
}

static std::vector<double> gaussian_1d(int ksize, double sigma) {
  std::vector<double> k(ksize);
  double s = 0.0;
  for (int i = 0; i < ksize; ++i) {
    const double x = double(i) - (ksize - 1) / 2.0;
    k[i] = std::exp(-(x * x) / (2.0 * sigma * sigma));
    s += k[i];
  }
  for (double& v : k) v /= s;
  return k;
}
static int reflect101(int i, int n) {
  i = std::abs(i);
  return i >= n ? 2 * (n - 1) - i : i;
}

struct Tap { int dy, dx; float w; };
// tap-list table: n_entries x [ int32 ntaps, pad[3], max_taps x {int16 dy | int16 dx << 16, float w} ]; geometry -> iparams
static void pack_taps(const std::vector<std::vector<Tap>>& entries, int border, std::vector<int32_t>& ip, std::vector<uint8_t>& tab) {
  size_t max_taps = 1;
  int dy_min = 0, dy_max = 0, dx_min = 0, dx_max = 0;
  for (const auto& e : entries) {
    max_taps = std::max(max_taps, e.size());
    for (const Tap& t : e) {
      dy_min = std::min(dy_min, t.dy); dy_max = std::max(dy_max, t.dy);
      dx_min = std::min(dx_min, t.dx); dx_max = std::max(dx_max, t.dx);
    }
  }
  const size_t rec = 16 + 8 * max_taps;
  const size_t base = tab.size();
  tab.resize(base + entries.size() * rec, 0);
  for (size_t i = 0; i < entries.size(); ++i) {
    uint8_t* o = tab.data() + base + i * rec;
    const int32_t nt = int32_t(entries[i].size());
    memcpy(o, &nt, 4);
    for (size_t t = 0; t < entries[i].size(); ++t) {
      const Tap& tp = entries[i][t];
      const uint32_t pos = uint32_t(uint16_t(int16_t(tp.dy))) | (uint32_t(uint16_t(int16_t(tp.dx))) << 16);
      memcpy(o + 16 + 8 * t, &pos, 4);
      memcpy(o + 16 + 8 * t + 4, &tp.w, 4);
    }
  }
  ip = {int32_t(entries.size()), int32_t(max_taps), border, dy_min, dy_max, dx_min, dx_max};
}

// defocus: aliased disk smoothed by a separable Gaussian with reflect-101 borders (make_imagenet_c.disk: cv2.GaussianBlur)
static std::vector<Tap> defocus_taps(double radius, double alias_blur) {
  int lo, hi, ks;
  if (radius <= 8) { lo = -8; hi = 8; ks = 3; } else { lo = -int(radius); hi = int(radius); ks = 5; }
  const int n = hi - lo + 1;
  std::vector<double> disk(size_t(n) * n), tmp(size_t(n) * n), out(size_t(n) * n);
  double s = 0.0;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      const double X = lo + x, Y = lo + y;
      disk[size_t(y) * n + x] = (X * X + Y * Y) <= radius * radius ? 1.0 : 0.0;
      s += disk[size_t(y) * n + x];
    }
  for (double& v : disk) v /= s;
  const std::vector<double> g = gaussian_1d(ks, alias_blur);
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      double a = 0.0;
      for (int t = 0; t < ks; ++t) a += disk[size_t(y) * n + reflect101(x + t - ks / 2, n)] * g[t];
      tmp[size_t(y) * n + x] = a;
    }
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      double a = 0.0;
      for (int t = 0; t < ks; ++t) a += tmp[size_t(reflect101(y + t - ks / 2, n)) * n + x] * g[t];
      out[size_t(y) * n + x] = a;
    }
  std::vector<Tap> taps;
  const int r = n / 2;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x)
      if (out[size_t(y) * n + x] != 0.0) taps.push_back({y - r, x - r, float(out[size_t(y) * n + x])});
  return taps;
}

// shift-and-add motion blur line (imagecorruptions formulation): taps whose shift leaves the frame end the line
static std::vector<Tap> motion_taps(int radius, double sigma, int angle_deg, int h, int w) {
  const int width = 2 * radius + 1;
  std::vector<double> k(width);
  double s = 0.0;
  for (int i = 0; i < width; ++i) { k[i] = std::exp(-(double(i) * double(i)) / (2.0 * sigma * sigma)); s += k[i]; }
  const double a = double(angle_deg) * (3.14159265358979323846 / 180.0);
  const double py = width * std::sin(a), px = width * std::cos(a);
  const double hyp = std::hypot(py, px);
  std::vector<Tap> taps;
  for (int i = 0; i < width; ++i) {
    const int sy = -int(std::ceil(i * py / hyp - 0.5));
    const int sx = -int(std::ceil(i * px / hyp - 0.5));
    if (std::abs(sy) >= h || std::abs(sx) >= w) break;
    taps.push_back({-sy, -sx, float(k[i] / s)});
  }
  return taps;
}

// clipped_zoom geometry of one axis: per output index {i0 | i1 << 16, frac}
static void zoom_axis(int n, double z, uint32_t* out2) {
  const int nc = int(std::ceil(n / z));
  const int top = (n - nc) / 2;
  const int no = int(std::nearbyint(nc * z));                   // Python round(): half to even
  const int trim = (no - n) / 2;                                // Python floor division of a non-negative value
  const float scale = no > 1 ? float(double(nc - 1) / double(no - 1)) : 0.0f;
  for (int o = 0; o < n; ++o) {
    const float src = (float(o) + float(trim)) * scale;
    int i0 = int(std::floor(src));
    i0 = std::min(std::max(i0, 0), nc - 1);
    const int i1 = std::min(i0 + 1, nc - 1);
    const float fr = src - float(i0);
    out2[2 * o] = uint32_t(i0 + top) | (uint32_t(i1 + top) << 16);
    memcpy(&out2[2 * o + 1], &fr, 4);
  }
}
static std::vector<double> zoom_factors(double zmax, double step) {
  // numpy.arange(1, zmax, step): ceil((zmax - 1) / step) values 1 + i * step
  // (numpy fills with start + i * delta, delta = (start + step) - start -- not exactly `step`)
  std::vector<double> zs;
  const int cnt = int(std::ceil((zmax - 1.0) / step));
  const volatile double next = 1.0 + step;
  const double delta = next - 1.0;
  for (int i = 0; i < cnt; ++i) zs.push_back(1.0 + i * delta);
  return zs;
}

// Pillow's BOX resampling of one axis (src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc):
// per output index {first source index, tap count, ksize fixed-point coefficients (22 fractional bits)}
static int pil_box_axis(int in_size, int out_size, std::vector<int32_t>& rec) {
  const double scale = double(in_size) / out_size;
  const double filterscale = std::max(scale, 1.0);
  const double support = 0.5 * filterscale;
  const int ksize = int(std::ceil(support)) * 2 + 1;
  rec.assign(size_t(out_size) * (2 + ksize), 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale, ss = 1.0 / filterscale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double a = (x + xmin - center + 0.5) * ss;
      k[x] = (a > -0.5 && a <= 0.5) ? 1.0 : 0.0;
      ww += k[x];
    }
    int32_t* o = rec.data() + size_t(xx) * (2 + ksize);
    o[0] = xmin; o[1] = xmax;
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? k[x] / ww : k[x];
      o[2 + x] = int32_t(v * 4194304.0 + (v < 0 ? -0.5 : 0.5));
    }
  }
  return ksize;
}

// libjpeg quantisation tables (jcparam.c jpeg_set_quality, force_baseline): lum[64], chr[64] in natural order
static const int JPEG_LUM[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                                 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const int JPEG_CHR[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                                 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// elastic smoothing folded into one weight per (destination, source) pixel: M[d][s] = sum of the fp32 taps k[t] whose
// source pixel reflect_sym(d + t - r, n) is s (float64 sums)
static void elastic_fold(const std::vector<float>& k, int r, int n, std::vector<double>& m) {
  m.assign(size_t(n) * n, 0.0);
  for (int d = 0; d < n; ++d)
    for (int t = 0; t <= 2 * r; ++t) {
      int src = (d + t - r) % (2 * n);
      if (src < 0) src += 2 * n;
      if (src >= n) src = 2 * n - 1 - src;
      m[size_t(d) * n + src] += double(k[t]);
    }
}

// ---------------------------------------------------------------------------------------------- the one entry point
int corrupt_params(int corruption, int severity, int h, int w, int profile, K1Params& out) {
  out.fp.clear(); out.ip.clear(); out.table.clear();
  if (corruption == FAV_CLEAN) return FAV_OK;
  if (corruption < 1 || corruption > 15 || severity < 1 || severity > 5 || (profile != 0 && profile != 1)) return FAV_E_ARG;
  const double* c = SEV[profile][corruption - 1][severity - 1];
  switch (corruption) {
    case FAV_GAUSSIAN_NOISE: case FAV_BRIGHTNESS: case FAV_CONTRAST:
      out.fp = {float(c[0])};
      break;
    case FAV_SHOT_NOISE: {
      int width = 0;
      size_t guide_off = 0;
      poisson_table(c[0], &width, &guide_off, out.table);
      out.fp = {float(c[0])};
      out.ip = {width, int32_t(guide_off)};
      break;
    }
    case FAV_IMPULSE_NOISE: {
      const uint32_t tp = uint32_t(std::floor(c[0] / 2 * 4294967296.0)), ts = uint32_t(std::floor(c[0] * 4294967296.0));
      out.ip = {int32_t(tp), int32_t(ts)};
      break;
    }
    case FAV_FOG:
      out.fp = {float(c[0]), float(c[1])};
      break;
    case FAV_FROST:                                              // procedural texture: c0, c1, plasma decay, icy tint
      out.fp = {float(c[0]), float(c[1]), 2.0f, 0.85f, 0.92f, 1.0f};
      break;
    case FAV_DEFOCUS_BLUR:
      pack_taps({defocus_taps(c[0], c[1])}, 0 /* reflect-101 (cv2.filter2D default) */, out.ip, out.table);
      break;
    case FAV_MOTION_BLUR: {
      std::vector<std::vector<Tap>> e;
      for (int a = 0; a < 91; ++a) e.push_back(motion_taps(int(c[0]), c[1], a - 45, h, w));   // integer degrees -45..45
      pack_taps(e, 1 /* clamp */, out.ip, out.table);
      break;
    }
    case FAV_ZOOM_BLUR: {
      const std::vector<double> zs = zoom_factors(c[0], c[1]);
      std::vector<uint32_t> t(zs.size() * size_t(h + w) * 2);
      for (size_t i = 0; i < zs.size(); ++i) {
        zoom_axis(h, zs[i], t.data() + i * size_t(h + w) * 2);
        zoom_axis(w, zs[i], t.data() + i * size_t(h + w) * 2 + size_t(h) * 2);
      }
      append(out.table, t.data(), t.size());
      out.ip = {int32_t(zs.size())};
      break;
    }
    case FAV_PIXELATE: {
      // PIL: x.resize((int(w c), int(h c)), BOX).resize((w, h), BOX); horizontal pass first, u8 between the passes
      const int sw = std::max(1, int(w * c[0])), sh = std::max(1, int(h * c[0]));
      std::vector<int32_t> hx, vy, ux, uy;
      const int kh = pil_box_axis(w, sw, hx), kv = pil_box_axis(h, sh, vy);
      const int ku = pil_box_axis(sw, w, ux), kw = pil_box_axis(sh, h, uy);
      // BOX up-sampling always has exactly one tap of weight 1 (support 0.5): store the source index only
      std::vector<int32_t> upx(w), upy(h);
      for (int x = 0; x < w; ++x) { if (ux[size_t(x) * (2 + ku) + 1] != 1) return FAV_E_UNSUPPORTED; upx[x] = ux[size_t(x) * (2 + ku)]; }
      for (int y = 0; y < h; ++y) { if (uy[size_t(y) * (2 + kw) + 1] != 1) return FAV_E_UNSUPPORTED; upy[y] = uy[size_t(y) * (2 + kw)]; }
      out.ip = {sw, sh, kh, kv};
      append(out.table, hx.data(), hx.size());
      append(out.table, vy.data(), vy.size());
      append(out.table, upx.data(), upx.size());
      append(out.table, upy.data(), upy.size());
      break;
    }
    case FAV_JPEG: {
      const int q = std::min(std::max(int(c[0]), 1), 100);
      const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
      std::vector<int32_t> t(128);
      for (int i = 0; i < 64; ++i) {
        t[i] = std::min(std::max((JPEG_LUM[i] * scale + 50) / 100, 1), 255);
        t[64 + i] = std::min(std::max((JPEG_CHR[i] * scale + 50) / 100, 1), 255);
      }
      append(out.table, t.data(), t.size());
      out.ip = {q};
      break;
    }
    case FAV_GLASS_BLUR: {
      const double sigma = c[0];
      const int r = int(4.0 * sigma + 0.5);
      std::vector<double> k(2 * r + 1);
      double s = 0.0;
      for (int i = 0; i <= 2 * r; ++i) { const double x = (i - r) / sigma; k[i] = std::exp(-0.5 * x * x); s += k[i]; }
      std::vector<int32_t> q(2 * r + 1);
      std::vector<float> kf(2 * r + 1);
      long long qs = 0;
      for (int i = 0; i <= 2 * r; ++i) { k[i] /= s; q[i] = int32_t(std::nearbyint(k[i] * 65536.0)); qs += q[i]; kf[i] = float(k[i]); }
      q[r] += int32_t(65536 - qs);
      append(out.table, q.data(), q.size());
      append(out.table, kf.data(), kf.size());
      out.ip = {int32_t(c[1]), int32_t(c[2]), r};
      break;
    }
    case FAV_ELASTIC: {
      // ImageNet-C scales its constants by the literal 244 on 224-pixel images, CIFAR-10-C by the image size 32
      const double S = profile == 1 ? 244.0 * std::min(h, w) / 224.0 : double(std::min(h, w));
      const double alpha = c[0] * S, sigma = c[1] * S, mag = c[2] * S;
      int r = int(3.0 * sigma + 0.5);
      std::vector<float> k;
      if (sigma <= 1e-6) {
        r = 0; k = {1.0f};
      } else {
        std::vector<double> kd(2 * r + 1);
        double s = 0.0;
        for (int i = 0; i <= 2 * r; ++i) { const double x = (i - r) / sigma; kd[i] = std::exp(-0.5 * x * x); s += kd[i]; }
        for (double v : kd) k.push_back(float(v / s));
      }
      out.fp = {float(alpha), float(mag), float(h / 2), float(w / 2), float(std::min(h, w) / 3)};
      if (2 * r + 1 > std::min(h, w) / 4) {
        // a long kernel wraps around the reflected row: fold it into [w][w] transposed ([src x][dst x]) then [h][h] ([dst y][src y])
        std::vector<double> mw, mh;
        elastic_fold(k, r, w, mw);
        elastic_fold(k, r, h, mh);
        std::vector<float> f(size_t(w) * w + size_t(h) * h);
        for (int s2 = 0; s2 < w; ++s2)
          for (int d = 0; d < w; ++d) f[size_t(s2) * w + d] = float(mw[size_t(d) * w + s2]);
        for (size_t i = 0; i < size_t(h) * h; ++i) f[size_t(w) * w + i] = float(mh[i]);
        append(out.table, f.data(), f.size());
        out.ip = {r, 1};
      } else {
        append(out.table, k.data(), k.size());
        out.ip = {r, 0};
      }
      break;
    }
    case FAV_SNOW: {
      std::vector<std::vector<Tap>> e;
      for (int a = 0; a < 91; ++a) e.push_back(motion_taps(int(c[4]), c[5], a - 135, h, w));  // integer degrees -135..-45
      pack_taps(e, 1, out.ip, out.table);
      while (out.table.size() % 16) out.table.push_back(0);
      const int taps_bytes = int(out.table.size());
      std::vector<uint32_t> z(size_t(h + w) * 2);
      zoom_axis(h, c[2], z.data());
      zoom_axis(w, c[2], z.data() + size_t(h) * 2);
      append(out.table, z.data(), z.size());
      out.ip.push_back(taps_bytes);
      const float blend = float(c[6]);
      out.fp = {float(c[0]), float(c[1]), float(c[3]), blend, 1.0f - blend, float(1.0 / (65536.0 * std::sqrt(8.0 / 12.0)))};
      break;
    }
    default:
      return FAV_E_UNSUPPORTED;
  }
  return FAV_OK;
}

// per-handle cache: built once per (corruption, severity, h, w, profile), table uploaded to the device
struct K1Cache {
  std::map<std::tuple<int, int, int, int, int>, K1Entry> entries;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

int k1_lookup(Ctx* ctx, int corruption, int severity, int h, int w, int profile, cudaStream_t st, const K1Entry** out) {
  if (!ctx->k1_cache) ctx->k1_cache = new K1Cache();
  K1Cache* cache = static_cast<K1Cache*>(ctx->k1_cache);
  const auto key = std::make_tuple(corruption, severity, h, w, profile);
  auto it = cache->entries.find(key);
  if (it == cache->entries.end()) {
    K1Entry e;
    const int rc = corrupt_params(corruption, severity, h, w, profile, e.p);
    if (rc != FAV_OK) { set_error("no host table for corruption %d severity %d at %dx%d", corruption, severity, h, w); return rc; }
    if (!e.p.table.empty()) {
      FAV_CUDA_OK(cudaMalloc(&e.d_table, e.p.table.size()));
      // synchronous upload (first use of a cell only): the host vector may be reallocated by later insertions
      FAV_CUDA_OK(cudaMemcpy(e.d_table, e.p.table.data(), e.p.table.size(), cudaMemcpyHostToDevice));
    }
    (void)st;
    it = cache->entries.emplace(key, std::move(e)).first;
  }
  *out = &it->second;
  return FAV_OK;
}

int k1_scratch(Ctx* ctx, size_t bytes, void** out) {
  if (!ctx->k1_cache) ctx->k1_cache = new K1Cache();
  K1Cache* cache = static_cast<K1Cache*>(ctx->k1_cache);
  if (bytes > cache->scratch_bytes) {
    if (cache->scratch) FAV_CUDA_OK(cudaFree(cache->scratch));          // cudaFree synchronises: earlier users are done
    cache->scratch = nullptr; cache->scratch_bytes = 0;
    const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    FAV_CUDA_OK(cudaMalloc(&cache->scratch, want));
    cache->scratch_bytes = want;
  }
  *out = cache->scratch;
  return FAV_OK;
}

int k1_scratch_release(Ctx* ctx) {
  K1Cache* cache = static_cast<K1Cache*>(ctx->k1_cache);
  if (cache && cache->scratch) {
    FAV_CUDA_OK(cudaFree(cache->scratch));
    cache->scratch = nullptr;
    cache->scratch_bytes = 0;
  }
  return FAV_OK;
}

void k1_cache_destroy(Ctx* ctx) {
  K1Cache* cache = static_cast<K1Cache*>(ctx->k1_cache);
  if (!cache) return;
  for (auto& kv : cache->entries)
    if (kv.second.d_table) cudaFree(kv.second.d_table);
  if (cache->scratch) cudaFree(cache->scratch);
  delete cache;
  ctx->k1_cache = nullptr;
}

}  // namespace fav

using namespace fav;

extern "C" int fav_corruption_constants(int profile, int corruption, int severity, double* out, int cap) {
  FAV_REQUIRE(out && (profile == 0 || profile == 1) && corruption >= 1 && corruption <= 15 && severity >= 1 && severity <= 5,
              "fav_corruption_constants: bad arguments");
  const int cnt = SEV_COUNT[corruption - 1];
  FAV_REQUIRE(cap >= cnt, "fav_corruption_constants: need room for %d values", cnt);
  for (int i = 0; i < cnt; ++i) out[i] = SEV[profile][corruption - 1][severity - 1][i];
  return cnt;
}

extern "C" int fav_corrupt_params(int corruption, int severity, int height, int width, unsigned flags, float* fparams,
                                  int* n_fparams, int32_t* iparams, int* n_iparams, void* table, size_t* table_bytes) {
  FAV_REQUIRE(n_fparams && n_iparams && table_bytes, "fav_corrupt_params: null count pointer");
  FAV_REQUIRE(height > 0 && width > 0, "fav_corrupt_params: bad frame size");
  K1Params p;
  const int rc = corrupt_params(corruption, severity, height, width, profile_of(flags, height, width), p);
  if (rc != FAV_OK) { set_error("fav_corrupt_params: corruption %d severity %d is not defined", corruption, severity); return rc; }
  const bool fits = int(p.fp.size()) <= *n_fparams && int(p.ip.size()) <= *n_iparams && p.table.size() <= *table_bytes;
  *n_fparams = int(p.fp.size()); *n_iparams = int(p.ip.size()); *table_bytes = p.table.size();
  if (!fits) return FAV_OK;                          // sizes reported; call again with room
  if (fparams && !p.fp.empty()) memcpy(fparams, p.fp.data(), p.fp.size() * 4);
  if (iparams && !p.ip.empty()) memcpy(iparams, p.ip.data(), p.ip.size() * 4);
  if (table && !p.table.empty()) memcpy(table, p.table.data(), p.table.size());
  return 1;                                          // filled
}
