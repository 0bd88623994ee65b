"""Corruption / severity configuration and the host-side tables the K1 kernels consume.

Reference surface being extended: ``VisionSimulator.set_noise / set_brightness / set_mode``
(platform/backend/vision_simulator.py:25-36) -- two sliders and four modes.  Here the config is
the 15-corruption x 5-severity grid of Hendrycks & Dietterich.  This
module is numpy-only product code; it never imports ``oracle`` (tests compare the two).
"""
import math
from dataclasses import dataclass

import numpy as np

CORRUPTIONS = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur",
    "motion_blur", "zoom_blur", "snow", "frost", "fog", "brightness", "contrast",
    "elastic_transform", "pixelate", "jpeg_compression",
)
CORRUPTION_ID = {name: i + 1 for i, name in enumerate(CORRUPTIONS)}
# corruptions with a device kernel in this build; the sweep reports the rest as unavailable
IMPLEMENTED = CORRUPTIONS      # all 15 have device kernels

SEVERITY = {
    "imagenet": {
        "gaussian_noise": [.08, .12, .18, .26, .38],
        "shot_noise": [60, 25, 12, 5, 3],
        "impulse_noise": [.03, .06, .09, .17, .27],
        "defocus_blur": [(3, .1), (4, .5), (6, .5), (8, .5), (10, .5)],
        "motion_blur": [(10, 3), (15, 5), (15, 8), (15, 12), (20, 15)],
        "zoom_blur": [(1.11, .01), (1.16, .01), (1.21, .02), (1.26, .02), (1.33, .03)],
        "fog": [(1.5, 2), (2., 2), (2.5, 1.7), (2.5, 1.5), (3., 1.4)],
        "brightness": [.1, .2, .3, .4, .5],
        "contrast": [.4, .3, .2, .1, .05],
        "pixelate": [.6, .5, .4, .3, .25],
        "jpeg_compression": [25, 18, 15, 10, 7],
        "glass_blur": [(.7, 1, 2), (.9, 2, 1), (1, 2, 3), (1.1, 3, 2), (1.5, 4, 2)],
        "frost": [(1, .4), (.8, .6), (.7, .7), (.65, .7), (.6, .75)],
        "snow": [(.1, .3, 3, .5, 10, 4, .8), (.2, .3, 2, .5, 12, 4, .7), (.55, .3, 4, .9, 12, 8, .7),
                 (.55, .3, 4.5, .85, 12, 8, .65), (.55, .3, 2.5, .85, 12, 12, .55)],
        "elastic_transform": [(2., .7, .1), (2., .08, .2), (.05, .01, .02), (.07, .01, .02), (.12, .01, .02)],
    },
    "cifar": {
        "gaussian_noise": [.04, .06, .08, .09, .10],
        "shot_noise": [500, 250, 100, 75, 50],
        "impulse_noise": [.01, .02, .03, .05, .07],
        "defocus_blur": [(.3, .4), (.4, .5), (.5, .6), (1, .2), (1.5, .1)],
        "motion_blur": [(10, 1), (10, 1.5), (10, 2), (10, 2.5), (12, 3)],
        "zoom_blur": [(1.06, .01), (1.11, .01), (1.16, .01), (1.21, .01), (1.26, .01)],
        "fog": [(.2, 3), (.5, 3), (.75, 2.5), (1, 2), (1.5, 1.75)],
        "brightness": [.05, .1, .15, .2, .3],
        "contrast": [.75, .5, .4, .3, .15],
        "pixelate": [.95, .9, .85, .75, .65],
        "jpeg_compression": [80, 65, 58, 50, 40],
        "glass_blur": [(.05, 1, 1), (.25, 1, 1), (.4, 1, 1), (.25, 1, 2), (.4, 1, 2)],
        "frost": [(1, .2), (1, .3), (.9, .4), (.85, .4), (.75, .45)],
        "snow": [(.1, .2, 1, .6, 8, 3, .95), (.1, .2, 1, .5, 10, 4, .9), (.15, .3, 1.75, .55, 10, 4, .9),
                 (.25, .3, 2.25, .6, 12, 6, .85), (.3, .3, 1.25, .65, 14, 12, .8)],
        "elastic_transform": [(0, 0, .08), (.05, .2, .07), (.08, .06, .06), (.1, .04, .05), (.1, .03, .03)],
    },
}

MEAN_STD = {
    "imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "cifar": ((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616)),
}


def profile_for(h, w):
    """CIFAR-10-C constants for frames up to 64 px, ImageNet-C constants above."""
    return "cifar" if max(h, w) <= 64 else "imagenet"


@dataclass(frozen=True)
class CorruptionConfig:
    """One cell of the sweep grid.  ``name`` None / 'clean' with severity 0 is the clean cell.

    Mirrors the reference's permissive style (vision_simulator.py:27 silently ignores unknown
    modes) only in spirit: here an unknown name or severity raises, because a silently-ignored
    cell would corrupt a benchmark table."""
    name: str = None
    severity: int = 0

    def __post_init__(self):
        if self.name in (None, "clean", "none"):
            object.__setattr__(self, "name", None)
            object.__setattr__(self, "severity", 0)
            return
        if self.name not in CORRUPTION_ID:
            raise ValueError(f"unknown corruption '{self.name}'; expected one of {CORRUPTIONS}")
        if not 1 <= int(self.severity) <= 5:
            raise ValueError("severity must be in 1..5")
        object.__setattr__(self, "severity", int(self.severity))

    @property
    def id(self):
        return 0 if self.name is None else CORRUPTION_ID[self.name]

    def to_dict(self):
        return {"corruption": self.name or "clean", "severity": self.severity}


# ------------------------------------------------------------------------------- host tables
def poisson_table(c):
    """kmin int32[256], width, thr uint32[256,width]: 32-bit inverse-CDF thresholds of
    Poisson(v/255*c).  Built from the pmf recurrence in float64 (window +-7.5 sigma)."""
    lam = np.arange(256, dtype=np.float64) * (float(c) / 255.0)
    sd = np.sqrt(lam)
    kmin = np.maximum(0, np.floor(lam - 7.5 * sd - 4)).astype(np.int64)
    width = int(np.max(np.ceil(lam + 7.5 * sd + 12) - kmin)) + 1
    width = (width + 3) // 4 * 4
    thr = np.empty((256, width), dtype=np.uint32)
    lg = np.cumsum(np.log(np.maximum(np.arange(0, int(kmin.max()) + width + 2, dtype=np.float64), 1.0)))  # ln k!
    for v in range(256):
        ks = kmin[v] + np.arange(width)
        if lam[v] == 0.0:
            cdf = np.ones(width)
        else:
            logp = -lam[v] + ks * math.log(lam[v]) - lg[ks]
            cdf = np.cumsum(np.exp(logp))
            if kmin[v] > 0:                       # mass below the window (< 1e-13), from the lower tail
                kb = np.arange(0, kmin[v])
                cdf = cdf + np.exp(-lam[v] + kb * math.log(lam[v]) - lg[kb]).sum()
        thr[v] = np.minimum(np.floor(np.minimum(cdf, 1.0) * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    return kmin.astype(np.int32), width, thr


def _gaussian_1d(ksize, sigma):
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return k / k.sum()


def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def disk_kernel(radius, alias_blur):
    """Aliased disk smoothed by a separable Gaussian with reflect-101 borders."""
    if radius <= 8:
        L = np.arange(-8, 9)
        ks = 3
    else:
        L = np.arange(-int(radius), int(radius) + 1)
        ks = 5
    X, Y = np.meshgrid(L, L)
    disk = ((X * X + Y * Y) <= radius * radius).astype(np.float64)
    disk /= disk.sum()
    g = _gaussian_1d(ks, alias_blur)
    n = disk.shape[0]
    idx = _reflect101(np.arange(n)[:, None] + (np.arange(ks) - ks // 2)[None, :], n)      # [n, ks]
    tmp = (disk[:, idx] * g[None, None, :]).sum(-1)           # along x
    out = (tmp[idx, :] * g[None, :, None]).sum(1)             # along y
    return out


def _pack_taps(entries):
    """entries: list of (dys, dxs, ws) -> (iparams geometry, uint8 table)."""
    max_taps = max(1, max(len(e[2]) for e in entries))
    rec = 16 + 8 * max_taps
    buf = np.zeros(len(entries) * rec, dtype=np.uint8)
    all_dy = [d for e in entries for d in e[0]] + [0]
    all_dx = [d for e in entries for d in e[1]] + [0]
    for i, (dys, dxs, ws) in enumerate(entries):
        o = i * rec
        buf[o:o + 4] = np.array([len(ws)], dtype=np.int32).view(np.uint8)
        if len(ws):
            t = np.zeros((len(ws), 2), dtype=np.uint32)
            t[:, 0] = (np.asarray(dys, np.int16).view(np.uint16).astype(np.uint32)
                       | (np.asarray(dxs, np.int16).view(np.uint16).astype(np.uint32) << 16))
            t[:, 1] = np.asarray(ws, np.float32).view(np.uint32)
            buf[o + 16:o + 16 + 8 * len(ws)] = t.view(np.uint8).ravel()
    geom = [len(entries), max_taps, 0, min(all_dy), max(all_dy), min(all_dx), max(all_dx)]
    return geom, buf


def defocus_table(radius, alias_blur):
    k = disk_kernel(radius, alias_blur)
    r = k.shape[0] // 2
    iy, ix = np.nonzero(k)
    geom, buf = _pack_taps([((iy - r).tolist(), (ix - r).tolist(), k[iy, ix].astype(np.float32).tolist())])
    geom[2] = 0           # reflect-101 border (cv2.filter2D default)
    return geom, buf


MOTION_ANGLES = 91        # integer degrees -45..45, chosen per image by a Philox draw
SNOW_ANGLES = 91          # integer degrees -135..-45 for the snow layer's motion blur


def motion_taps(radius, sigma, angle_deg, h, w):
    width = 2 * radius + 1
    k = np.exp(-(np.arange(width, dtype=np.float64) ** 2) / (2.0 * sigma * sigma))
    k = (k / k.sum()).astype(np.float32)
    a = math.radians(angle_deg)
    py, pxx = width * math.sin(a), width * math.cos(a)
    hyp = math.hypot(py, pxx)
    dys, dxs, ws = [], [], []
    for i in range(width):
        sy = -math.ceil(i * py / hyp - 0.5)
        sx = -math.ceil(i * pxx / hyp - 0.5)
        if abs(sy) >= h or abs(sx) >= w:          # shift left the frame: the rest of the line is dropped
            break
        dys.append(-sy), dxs.append(-sx), ws.append(float(k[i]))
    return dys, dxs, ws


def motion_table(radius, sigma, h, w):
    geom, buf = _pack_taps([motion_taps(radius, sigma, a - 45, h, w) for a in range(MOTION_ANGLES)])
    geom[2] = 1           # clamp border (edge replication)
    return geom, buf


def zoom_factors(spec):
    zmax, step = spec
    return [float(z) for z in np.arange(1.0, zmax, step)]


def _zoom_axis(n, z):
    nc = int(math.ceil(n / z))
    top = (n - nc) // 2
    no = int(round(nc * z))
    trim = (no - n) // 2
    o = np.arange(n, dtype=np.float32) + np.float32(trim)
    scale = np.float32((nc - 1) / (no - 1)) if no > 1 else np.float32(0)
    src = (o * scale).astype(np.float32)
    i0 = np.clip(np.floor(src).astype(np.int32), 0, nc - 1)
    i1 = np.minimum(i0 + 1, nc - 1)
    fr = (src - i0.astype(np.float32)).astype(np.float32)
    return i0 + top, i1 + top, fr


def zoom_table(spec, h, w):
    zs = zoom_factors(spec)
    tab = np.zeros((len(zs), h + w, 2), dtype=np.uint32)
    for i, z in enumerate(zs):
        for off, n in ((0, h), (h, w)):
            i0, i1, fr = _zoom_axis(n, z)
            tab[i, off:off + n, 0] = i0.astype(np.uint32) | (i1.astype(np.uint32) << 16)
            tab[i, off:off + n, 1] = fr.view(np.uint32)
    return len(zs), tab.view(np.uint8).ravel()


def _pixelate_axis(n, c):
    small = max(1, int(n * c))
    scale = n / small
    lo, hi = np.empty(small, np.int64), np.empty(small, np.int64)
    for j in range(small):
        ctr = (j + 0.5) * scale
        a, b = max(int(ctr - 0.5 * scale + 0.5), 0), min(int(ctr + 0.5 * scale + 0.5), n)
        lo[j], hi[j] = a, max(b, a + 1)
    up = np.minimum(((np.arange(n) + 0.5) * small / n).astype(np.int64), small - 1)
    return lo[up], hi[up]


def pixelate_table(c, h, w):
    tab = np.zeros(h + w, dtype=np.uint32)
    for off, n in ((0, h), (h, w)):
        lo, hi = _pixelate_axis(n, c)
        tab[off:off + n] = lo.astype(np.uint32) | (hi.astype(np.uint32) << 16)
    return tab.view(np.uint8)


# ---- jpeg: Annex-K tables scaled by the libjpeg quality rule, 13-bit fixed-point orthonormal DCT matrix
_JPEG_LUM = [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
             80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
             95, 98, 112, 100, 103, 99]
_JPEG_CHR = [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99,
             99, 99] + [99] * 32


def jpeg_table(quality):
    """int32[192]: luminance table, chrominance table (row-major v,u), DCT matrix T[u][x]."""
    q = int(min(max(quality, 1), 100))
    scale = 5000 // q if q < 50 else 200 - 2 * q
    tabs = [np.clip((np.asarray(b, dtype=np.int64) * scale + 50) // 100, 1, 255) for b in (_JPEG_LUM, _JPEG_CHR)]
    u = np.arange(8, dtype=np.float64)[:, None]
    x = np.arange(8, dtype=np.float64)[None, :]
    t = np.rint(8192.0 * np.where(u == 0, math.sqrt(0.125), 0.5) * np.cos((2 * x + 1) * u * math.pi / 16.0))
    return np.concatenate([tabs[0], tabs[1], t.ravel()]).astype(np.int32)


FROST_TINT = (0.85, 0.92, 1.0)
FROST_DECAY = 2.0


def elastic_fold(k, r, n):
    """M[d, s] = sum of the fp32 taps k[t] whose source pixel reflect_sym(d + t - r, n) is s (float64 sums)."""
    t = np.arange(2 * r + 1)
    m = np.zeros((n, n), dtype=np.float64)
    for d in range(n):
        src = np.mod(d + t - r, 2 * n)
        src = np.where(src >= n, 2 * n - 1 - src, src)
        np.add.at(m[d], src, k.astype(np.float64))
    return m


def glass_table(sigma):
    """radius, uint8 table = int32 fixed-point taps (sum 65536) followed by fp32 taps."""
    r = int(4.0 * float(sigma) + 0.5)
    xs = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (xs / float(sigma)) ** 2)
    k = k / k.sum()
    q = np.rint(k * 65536.0).astype(np.int64)
    q[r] += 65536 - q.sum()
    return r, np.concatenate([q.astype(np.int32).view(np.uint8), k.astype(np.float32).view(np.uint8)])


def kernel_params(cfg: CorruptionConfig, h, w, profile=None):
    """-> (fparams list, iparams list, table uint8 ndarray or None) for fav_corrupt_normalize."""
    if cfg.name is None:
        return [], [], None
    c = SEVERITY[profile or profile_for(h, w)][cfg.name][cfg.severity - 1]
    n = cfg.name
    if n == "gaussian_noise":
        return [float(c)], [], None
    if n == "shot_noise":
        kmin, width, thr = poisson_table(c)
        # jump[v][b] = #{j : thr[v][j] < b << 24}: where the linear probe for a draw with top byte b starts
        edges = (np.arange(256, dtype=np.uint64) << np.uint64(24))
        jump = (thr[:, None, :].astype(np.uint64) < edges[None, :, None]).sum(-1).astype(np.uint16)
        tab = np.concatenate([kmin.view(np.uint8), thr.view(np.uint8).ravel(), jump.view(np.uint8).ravel()])
        return [float(c)], [width], tab
    if n == "impulse_noise":
        tp, ts = int(math.floor(c / 2 * 2.0 ** 32)), int(math.floor(c * 2.0 ** 32))
        return [], [_as_i32(tp), _as_i32(ts)], None
    if n in ("brightness", "contrast"):
        return [float(c)], [], None
    if n == "fog":
        return [float(c[0]), float(c[1])], [], None
    if n == "defocus_blur":
        geom, buf = defocus_table(*c)
        return [], geom, buf
    if n == "motion_blur":
        geom, buf = motion_table(c[0], c[1], h, w)
        return [], geom, buf
    if n == "zoom_blur":
        nz, buf = zoom_table(c, h, w)
        return [], [nz], buf
    if n == "pixelate":
        return [], [], pixelate_table(c, h, w)
    if n == "jpeg_compression":
        return [], [int(c)], jpeg_table(c).view(np.uint8)
    if n == "frost":
        return [float(c[0]), float(c[1]), FROST_DECAY] + [float(t) for t in FROST_TINT], [], None
    if n == "glass_blur":
        r, tab = glass_table(c[0])
        return [], [int(c[1]), int(c[2]), r], tab
    if n == "elastic_transform":
        S = min(h, w)
        alpha, sigma, mag = float(c[0]) * S, float(c[1]) * S, float(c[2]) * S
        r = int(3.0 * sigma + 0.5)
        if sigma <= 1e-6:
            r, k = 0, np.ones(1, dtype=np.float32)
        else:
            xs = np.arange(-r, r + 1, dtype=np.float64)
            k = np.exp(-0.5 * (xs / sigma) ** 2)
            k = (k / k.sum()).astype(np.float32)
        fp = [alpha, mag, float(h // 2), float(w // 2), float(min(h, w) // 3)]
        if 2 * r + 1 > min(h, w) // 4:
            # a long kernel (224-pixel rows: 941 taps at severity 1, wrapping around the reflected row; 109 at severity 2): fold it
            # into one weight per (destination, source) pixel -- [w][w] transposed ([src x][dst x]) then [h][h] ([dst y][src y])
            return fp, [r, 1], np.concatenate([elastic_fold(k, r, w).T.ravel(), elastic_fold(k, r, h).ravel()]).astype(np.float32).view(np.uint8)
        return fp, [r, 0], k.view(np.uint8)
    if n == "snow":
        loc, scale, zoom, thresh, mb_r, mb_s, blend = c
        geom, taps = _pack_taps([motion_taps(int(mb_r), float(mb_s), a - 135, h, w) for a in range(SNOW_ANGLES)])
        geom[2] = 1                                    # clamp border
        ztab = np.zeros((h + w, 2), dtype=np.uint32)
        for off, size in ((0, h), (h, w)):
            i0, i1, fr = _zoom_axis(size, float(zoom))
            ztab[off:off + size, 0] = i0.astype(np.uint32) | (i1.astype(np.uint32) << 16)
            ztab[off:off + size, 1] = fr.view(np.uint32)
        pad = (-len(taps)) % 16
        tab = np.concatenate([taps, np.zeros(pad, np.uint8), ztab.view(np.uint8).ravel()])
        irwin_hall = float(np.float32(1.0 / (65536.0 * math.sqrt(8.0 / 12.0))))
        return ([float(loc), float(scale), float(thresh), float(blend), float(np.float32(1 - blend)), irwin_hall],
                geom + [len(taps) + pad], tab)
    raise AssertionError(n)


def _as_i32(u):
    """uint32 value -> the int32 with the same bits (C ABI passes int32 iparams)."""
    u &= 0xFFFFFFFF
    return u - (1 << 32) if u >= (1 << 31) else u
