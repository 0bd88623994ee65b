"""Image corruption generators -- CPU oracle (numpy / cv2 / scipy).  TEST INFRASTRUCTURE ONLY.

The reference has no server-side corruption generator (only display-only JS effects,
platform/frontend/js/app.js:789-799, :834-851); its "corruption config" is two sliders and
four modes (platform/backend/vision_simulator.py:25-36).  These generators restate the
Hendrycks & Dietterich (ICLR'19) definitions as recorded in SURVEY.md Appendix A.2, with all
randomness re-expressed on the counter-based Philox stream of oracle/philox.py so that a GPU
kernel can reproduce every draw (PARITY UNPINNED by the reference; this file is the definition).

Input : uint8 [N,H,W,3] (RGB order), severity 1..5.
Output: float32 [N,H,W,3] in [0,1].
"""
import math
import numpy as np

from . import philox as px

CORRUPTIONS = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur",
    "motion_blur", "zoom_blur", "snow", "frost", "fog", "brightness", "contrast",
    "elastic_transform", "pixelate", "jpeg_compression",
)
CORRUPTION_ID = {name: i + 1 for i, name in enumerate(CORRUPTIONS)}   # 0 = clean

# per-severity constants: [profile][name][severity-1]
CONSTANTS = {
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
        "snow": [(.1, .3, 3, .5, 10, 4, .8), (.2, .3, 2, .5, 12, 4, .7), (.55, .3, 4, .9, 12, 8, .7),
                 (.55, .3, 4.5, .85, 12, 8, .65), (.55, .3, 2.5, .85, 12, 12, .55)],
        "frost": [(1, .4), (.8, .6), (.7, .7), (.65, .7), (.6, .75)],
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
        "snow": [(.1, .2, 1, .6, 8, 3, .95), (.1, .2, 1, .5, 10, 4, .9), (.15, .3, 1.75, .55, 10, 4, .9),
                 (.25, .3, 2.25, .6, 12, 6, .85), (.3, .3, 1.25, .65, 14, 12, .8)],
        "frost": [(1, .2), (1, .3), (.9, .4), (.85, .4), (.75, .45)],
        "elastic_transform": [(0, 0, .08), (.05, .2, .07), (.08, .06, .06), (.1, .04, .05), (.1, .03, .03)],
    },
}


def profile_for(h, w):
    """CIFAR-10-C constants for small frames, ImageNet-C constants otherwise."""
    return "cifar" if max(h, w) <= 64 else "imagenet"


def _to_float(x_u8):
    return (x_u8.astype(np.float32) / np.float32(255.0)).astype(np.float32)


def _stream(name, severity, kind=px.KIND_CORRUPT):
    return px.stream_id(kind, CORRUPTION_ID[name], severity)


def _elem_draws(n, per, name, severity, seed, first_image, sub=0):
    """One u32 per element: [n, per] uint32. Element e uses word e%4 of Philox call e//4."""
    nch = (per + 3) // 4
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    ch = np.arange(nch, dtype=np.uint64)[None, :]
    xs = px.philox4x32_10(ch, img, sub, _stream(name, severity), seed)
    return np.stack(xs, axis=-1).reshape(n, nch * 4)[:, :per]


# ----------------------------------------------------------------------------- noise family
def gaussian_noise(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    c = np.float32(CONSTANTS[profile or profile_for(h, w)]["gaussian_noise"][severity - 1])
    per = h * w * 3
    nch = (per + 3) // 4
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    ch = np.arange(nch, dtype=np.uint64)[None, :]
    x0, x1, x2, x3 = px.philox4x32_10(ch, img, 0, _stream("gaussian_noise", severity), seed)
    z0, z1 = px.box_muller(x0, x1)
    z2, z3 = px.box_muller(x2, x3)
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(n, nch * 4)[:, :per].reshape(x_u8.shape)
    return np.clip(_to_float(x_u8) + c * z, 0, 1).astype(np.float32)


def poisson_table(c):
    """Inverse-CDF tables for Poisson(lambda = v/255*c), v = 0..255.

    Returns (kmin[256] int32, width int, thr[256, width] uint32) with
    thr[v, j] = floor(CDF(kmin[v] + j) * 2^32) clipped to 2^32-1; a draw u (uint32) maps to
    k = kmin[v] + #{j : thr[v, j] <= u}.  Integer compares only -> bit-exact on any device.
    """
    from scipy.stats import poisson
    lam = np.arange(256, dtype=np.float64) / 255.0 * float(c)
    sd = np.sqrt(lam)
    kmin = np.maximum(0, np.floor(lam - 7.5 * sd - 4)).astype(np.int64)
    width = int(np.max(np.ceil(lam + 7.5 * sd + 12) - kmin)) + 1
    width = (width + 3) // 4 * 4
    ks = kmin[:, None] + np.arange(width)[None, :]
    cdf = poisson.cdf(ks, lam[:, None])
    thr = np.minimum(np.floor(cdf * 2.0 ** 32), 2.0 ** 32 - 1).astype(np.uint64).astype(np.uint32)
    return kmin.astype(np.int32), width, thr


def shot_noise(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    c = CONSTANTS[profile or profile_for(h, w)]["shot_noise"][severity - 1]
    kmin, width, thr = poisson_table(c)
    per = h * w * 3
    u = _elem_draws(n, per, "shot_noise", severity, seed, first_image).reshape(x_u8.shape)
    rows = thr[x_u8]                                              # [n,h,w,3,width]
    k = kmin[x_u8] + (rows <= u[..., None]).sum(-1).astype(np.int32)
    return np.clip(k.astype(np.float32) / np.float32(c), 0, 1).astype(np.float32)


def impulse_thresholds(c):
    return int(math.floor(c / 2 * 2.0 ** 32)), int(math.floor(c * 2.0 ** 32))


def impulse_noise(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    c = CONSTANTS[profile or profile_for(h, w)]["impulse_noise"][severity - 1]
    tp, ts = impulse_thresholds(c)
    u = _elem_draws(n, h * w * 3, "impulse_noise", severity, seed, first_image).reshape(x_u8.shape)
    x = _to_float(x_u8)
    x = np.where(u < np.uint32(ts), np.float32(1.0), x)
    x = np.where(u < np.uint32(tp), np.float32(0.0), x)
    return x.astype(np.float32)


# ----------------------------------------------------------------------------- tap stencils
def disk_kernel(radius, alias_blur):
    """Aliased disk blurred by a small Gaussian (cv2), as in make_imagenet_c.disk()."""
    import cv2
    if radius <= 8:
        L = np.arange(-8, 8 + 1)
        ksize = (3, 3)
    else:
        L = np.arange(-int(radius), int(radius) + 1)
        ksize = (5, 5)
    X, Y = np.meshgrid(L, L)
    disk = np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.float32)
    disk /= disk.sum()
    return cv2.GaussianBlur(disk, ksize=ksize, sigmaX=alias_blur)


def _reflect101(i, n):
    i = np.abs(i)
    i = np.where(i >= n, 2 * (n - 1) - i, i)
    # very wide kernels on tiny images can overshoot twice
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def _apply_taps(x, dys, dxs, ws, border):
    """x float32 [N,H,W,3]; out = sum_i w_i * x[border(y+dy_i), border(x+dx_i)], fp32 in tap order."""
    n, h, w, _ = x.shape
    ys = np.arange(h)
    xs = np.arange(w)
    out = np.zeros_like(x)
    for dy, dx, wt in zip(dys, dxs, ws):
        if border == "reflect101":
            yy, xx = _reflect101(ys + dy, h), _reflect101(xs + dx, w)
        else:
            yy, xx = np.clip(ys + dy, 0, h - 1), np.clip(xs + dx, 0, w - 1)
        out += np.float32(wt) * x[:, yy][:, :, xx]
    return out


def defocus_taps(radius, alias_blur):
    k = disk_kernel(radius, alias_blur)
    r = k.shape[0] // 2
    dys, dxs, ws = [], [], []
    for iy in range(k.shape[0]):
        for ix in range(k.shape[1]):
            if k[iy, ix] != 0:
                # cv2.filter2D is correlation: out(y,x) = sum k(iy,ix) * src(y+iy-r, x+ix-r)
                dys.append(iy - r), dxs.append(ix - r), ws.append(np.float32(k[iy, ix]))
    return dys, dxs, ws


def defocus_blur(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    radius, alias = CONSTANTS[profile or profile_for(h, w)]["defocus_blur"][severity - 1]
    dys, dxs, ws = defocus_taps(radius, alias)
    return np.clip(_apply_taps(_to_float(x_u8), dys, dxs, ws, "reflect101"), 0, 1).astype(np.float32)


MOTION_ANGLES = 91          # integer degrees -45..45


def motion_taps(radius, sigma, angle_deg):
    """Shift-and-add motion blur taps (imagecorruptions-package formulation, no Wand)."""
    width = radius * 2 + 1
    k = np.exp(-(np.arange(width, dtype=np.float64) ** 2) / (2.0 * sigma ** 2))
    k = (k / k.sum()).astype(np.float32)
    a = math.radians(angle_deg)
    p0, p1 = width * math.sin(a), width * math.cos(a)
    hyp = math.hypot(p0, p1)
    dys, dxs = [], []
    for i in range(width):
        dy = -math.ceil((i * p0) / hyp - 0.5)
        dx = -math.ceil((i * p1) / hyp - 0.5)
        # shifted image S(y,x) = X(clamp(y-dy), clamp(x-dx))  ->  tap offset = (-dy, -dx)
        dys.append(-dy), dxs.append(-dx)
    return dys, dxs, list(k)


def motion_angle_index(n, severity, seed, first_image):
    img = np.arange(first_image, first_image + n, dtype=np.uint64)
    x0, _, _, _ = px.philox4x32_10(0, img, 0, _stream("motion_blur", severity, px.KIND_AUX), seed)
    return (x0 % np.uint32(MOTION_ANGLES)).astype(np.int32)


def motion_blur(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    radius, sigma = CONSTANTS[profile or profile_for(h, w)]["motion_blur"][severity - 1]
    aidx = motion_angle_index(n, severity, seed, first_image)
    x = _to_float(x_u8)
    out = np.empty_like(x)
    for i in range(n):
        dys, dxs, ws = motion_taps(radius, sigma, int(aidx[i]) - 45)
        # taps whose shift leaves the frame are dropped from that point on (reference 'break')
        keep = len(ws)
        for j, (dy, dx) in enumerate(zip(dys, dxs)):
            if abs(dy) >= h or abs(dx) >= w:
                keep = j
                break
        out[i] = _apply_taps(x[i:i + 1], dys[:keep], dxs[:keep], ws[:keep], "clamp")[0]
    return np.clip(out, 0, 1).astype(np.float32)


def zoom_factors(spec):
    zmax, step = spec
    return [float(z) for z in np.arange(1.0, zmax, step)]


def zoom_geometry(h, z):
    """clipped_zoom geometry: crop hc rows at `top`, bilinear-zoom to ho rows, trim `trim` rows."""
    hc = int(math.ceil(h / z))
    top = (h - hc) // 2
    ho = int(round(hc * z))
    trim = (ho - h) // 2
    return hc, top, ho, trim


def _zoom_sample_axis(h, z):
    """For each output index o in [0,h): (i0, i1, frac) into the uncropped axis."""
    hc, top, ho, trim = zoom_geometry(h, z)
    o = np.arange(h, dtype=np.float32) + np.float32(trim)
    scale = np.float32((hc - 1) / (ho - 1)) if ho > 1 else np.float32(0)
    src = (o * scale).astype(np.float32)
    i0 = np.floor(src).astype(np.int32)
    i0 = np.clip(i0, 0, hc - 1)
    i1 = np.minimum(i0 + 1, hc - 1)
    fr = (src - i0.astype(np.float32)).astype(np.float32)
    return i0 + top, i1 + top, fr


def zoom_blur(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    zs = zoom_factors(CONSTANTS[profile or profile_for(h, w)]["zoom_blur"][severity - 1])
    x = _to_float(x_u8)
    acc = x.copy()
    for z in zs:
        y0, y1, fy = _zoom_sample_axis(h, z)
        x0, x1, fx = _zoom_sample_axis(w, z)
        fy_ = fy[None, :, None, None]
        fx_ = fx[None, None, :, None]
        top = x[:, y0][:, :, x0] * (1 - fx_) + x[:, y0][:, :, x1] * fx_
        bot = x[:, y1][:, :, x0] * (1 - fx_) + x[:, y1][:, :, x1] * fx_
        acc += (top * (1 - fy_) + bot * fy_).astype(np.float32)
    out = acc / np.float32(len(zs) + 1)
    return np.clip(out, 0, 1).astype(np.float32)


# ----------------------------------------------------------------------------- colour / stats
def brightness(x_u8, severity, seed=0, first_image=0, profile=None):
    """HSV value shift: V <- clip(V+c)  ==  RGB * min(V+c,1)/V  (V = max RGB; V = 0 -> grey c)."""
    n, h, w, _ = x_u8.shape
    c = np.float32(CONSTANTS[profile or profile_for(h, w)]["brightness"][severity - 1])
    x = _to_float(x_u8)
    v = x.max(axis=-1, keepdims=True)
    v2 = np.minimum(v + c, np.float32(1.0))
    safe = np.where(v > 0, v, np.float32(1.0))
    out = np.where(v > 0, x * (v2 / safe), v2)
    return np.clip(out, 0, 1).astype(np.float32)


def contrast(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    c = np.float32(CONSTANTS[profile or profile_for(h, w)]["contrast"][severity - 1])
    s = x_u8.astype(np.int64).sum(axis=(1, 2), keepdims=True)               # exact integer sums
    mu = (s.astype(np.float32) / np.float32(255.0 * h * w)).astype(np.float32)
    x = _to_float(x_u8)
    return np.clip((x - mu) * c + mu, 0, 1).astype(np.float32)


def fog_mapsize(h, w):
    m = 1
    while m < max(h, w):
        m *= 2
    return m


def plasma_fractal(n, mapsize, wibbledecay, severity, seed=0, first_image=0, name="fog"):
    """Diamond-square plasma on a torus, fp32, noise from Philox keyed by the written cell."""
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    cell = np.arange(mapsize * mapsize, dtype=np.uint64)[None, :]
    x0, _, _, _ = px.philox4x32_10(cell, img, 0, _stream(name, severity), seed)
    u = px.u32_to_uniform(x0).reshape(n, mapsize, mapsize)
    noise = (np.float32(2.0) * u - np.float32(1.0)).astype(np.float32)       # U(-1,1)
    M = np.zeros((n, mapsize, mapsize), dtype=np.float32)
    step = mapsize
    wib = np.float32(100.0)
    dec = np.float32(wibbledecay)
    while step >= 2:
        hf = step // 2
        w2 = np.float32(wib * wib)
        ul = M[:, 0:mapsize:step, 0:mapsize:step]
        sq = (ul + np.roll(ul, -1, axis=1)) + (np.roll(ul, -1, axis=2) + np.roll(np.roll(ul, -1, axis=1), -1, axis=2))
        M[:, hf:mapsize:step, hf:mapsize:step] = sq * np.float32(0.25) + w2 * noise[:, hf:mapsize:step, hf:mapsize:step]
        dr = M[:, hf:mapsize:step, hf:mapsize:step]
        ul = M[:, 0:mapsize:step, 0:mapsize:step]
        lt = (dr + np.roll(dr, 1, axis=1)) + (ul + np.roll(ul, -1, axis=2))
        M[:, 0:mapsize:step, hf:mapsize:step] = lt * np.float32(0.25) + w2 * noise[:, 0:mapsize:step, hf:mapsize:step]
        tt = (dr + np.roll(dr, 1, axis=2)) + (ul + np.roll(ul, -1, axis=1))
        M[:, hf:mapsize:step, 0:mapsize:step] = tt * np.float32(0.25) + w2 * noise[:, hf:mapsize:step, 0:mapsize:step]
        step //= 2
        wib = np.float32(wib / dec)
    mn = M.min(axis=(1, 2), keepdims=True)
    M = M - mn
    mx = M.max(axis=(1, 2), keepdims=True)
    return (M / mx).astype(np.float32)


def fog(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    c0, c1 = CONSTANTS[profile or profile_for(h, w)]["fog"][severity - 1]
    c0 = np.float32(c0)
    pl = plasma_fractal(n, fog_mapsize(h, w), c1, severity, seed, first_image)[:, :h, :w, None]
    x = _to_float(x_u8)
    mx = x.max(axis=(1, 2, 3), keepdims=True)
    out = (x + c0 * pl) * (mx / (mx + c0))
    return np.clip(out, 0, 1).astype(np.float32)


PIL_PRECISION_BITS = 32 - 8 - 2


def pil_box_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BOX filter (src/libImaging/Resample.c): per output
    index (first source index, tap count) and the 22-bit fixed-point coefficients.  box_filter(x) = 1 for -0.5 < x <= 0.5."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 0.5 * filterscale
    bounds, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [1.0 if -0.5 < (x + xmin - center + 0.5) * ss <= 0.5 else 0.0 for x in range(xmax)]
        ww = sum(k)
        k = [v / ww if ww != 0.0 else v for v in k]
        bounds.append((xmin, xmax))
        kk.append([int(v * (1 << PIL_PRECISION_BITS) + (0.5 if v >= 0 else -0.5)) for v in k])
    return bounds, kk


def pil_box_resample_axis(img_u8, axis, out_size):
    """ImagingResampleHorizontal_8bpc / Vertical_8bpc: ss = 2^21 + sum(pixel * k); clip8(ss >> 22)."""
    img = np.moveaxis(img_u8, axis, 0).astype(np.int64)
    bounds, kk = pil_box_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.int64)
    for i, ((xmin, cnt), k) in enumerate(zip(bounds, kk)):
        ss = np.full(img.shape[1:], 1 << (PIL_PRECISION_BITS - 1), dtype=np.int64)
        for t in range(cnt):
            ss += img[xmin + t] * k[t]
        out[i] = np.clip(ss >> PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def pil_box_resize(img_u8, out_w, out_h):
    """Image.resize((out_w, out_h), BOX) on [..., H, W, 3] uint8: horizontal pass first, uint8 between the passes."""
    hax, wax = img_u8.ndim - 3, img_u8.ndim - 2
    t = pil_box_resample_axis(img_u8, wax, out_w) if out_w != img_u8.shape[wax] else img_u8
    return pil_box_resample_axis(t, hax, out_h) if out_h != img_u8.shape[hax] else t


def pixelate(x_u8, severity, seed=0, first_image=0, profile=None):
    """make_imagenet_c.pixelate: x.resize((int(w c), int(h c)), BOX).resize((w, h), BOX), restated after Pillow's
    Resample.c; tests/test_oracle.py pins it byte for byte to PIL itself."""
    n, h, w, _ = x_u8.shape
    c = CONSTANTS[profile or profile_for(h, w)]["pixelate"][severity - 1]
    small = pil_box_resize(x_u8, max(1, int(w * c)), max(1, int(h * c)))
    return _to_float(pil_box_resize(small, w, h))


def pixelate_pil(x_u8, c):
    """The real thing: Pillow, image by image (anchor for the restatement above)."""
    from PIL import Image
    out = np.empty_like(x_u8)
    h, w = x_u8.shape[1:3]
    for i in range(x_u8.shape[0]):
        im = Image.fromarray(x_u8[i])
        out[i] = np.array(im.resize((max(1, int(w * c)), max(1, int(h * c))), Image.BOX).resize((w, h), Image.BOX))
    return out


# ----------------------------------------------------------------------------- f2 corruptions
def jpeg_compression(x_u8, severity, seed=0, first_image=0, profile=None):
    """PIL save(quality=c) + reload, restated after libjpeg (oracle/jpeg.py; pinned byte for byte to Pillow)."""
    from . import jpeg as J
    n, h, w, _ = x_u8.shape
    q = CONSTANTS[profile or profile_for(h, w)]["jpeg_compression"][severity - 1]
    return _to_float(J.jpeg_roundtrip_u8(x_u8, q))


FROST_TINT = (0.85, 0.92, 1.0)
FROST_DECAY = 2.0


def frost(x_u8, severity, seed=0, first_image=0, profile=None):
    """Procedural substitute (documented deviation: the six frost photographs of ImageNet-C are not available):
    frost texture = diamond-square plasma (decay 2.0) through a contrast curve, tinted icy blue;
    out = clip(c0 * x + c1 * frost)."""
    n, h, w, _ = x_u8.shape
    c0, c1 = CONSTANTS[profile or profile_for(h, w)]["frost"][severity - 1]
    c0, c1 = np.float32(c0), np.float32(c1)
    pl = plasma_fractal(n, fog_mapsize(h, w), FROST_DECAY, severity, seed, first_image, name="frost")[:, :h, :w, None]
    f = np.clip(np.float32(1.35) * pl - np.float32(0.1), 0, 1).astype(np.float32)
    tex = (f * np.asarray(FROST_TINT, dtype=np.float32)).astype(np.float32)
    return np.clip(c0 * _to_float(x_u8) + c1 * tex, 0, 1).astype(np.float32)


def gaussian_taps(sigma):
    """scipy / skimage style 1-D Gaussian: radius int(4 sigma + 0.5), normalised."""
    r = int(4.0 * float(sigma) + 0.5)
    xs = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (xs / float(sigma)) ** 2)
    return r, k / k.sum()


def gaussian_taps_q16(sigma):
    """Fixed-point taps (sum exactly 65536) for the exact integer first blur of glass_blur."""
    r, k = gaussian_taps(sigma)
    q = np.rint(k * 65536.0).astype(np.int64)
    q[r] += 65536 - q.sum()
    return r, q


def glass_swaps(n, h, w, delta, iters, severity, seed, first_image):
    """(dy, dx) int arrays [n, iters * (h - 2 delta) * (w - 2 delta)] in scan order (h, w descending)."""
    steps = iters * (h - 2 * delta) * (w - 2 * delta)
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    j = np.arange(steps, dtype=np.uint64)[None, :]
    x0, x1, _, _ = px.philox4x32_10(j, img, 0, _stream("glass_blur", severity), seed)
    m = np.uint32(2 * delta)
    return (x1 % m).astype(np.int64) - delta, (x0 % m).astype(np.int64) - delta


def glass_blur(x_u8, severity, seed=0, first_image=0, profile=None):
    """gaussian(sigma) -> uint8 -> `iters` passes of local pixel swaps within +-delta (scan order, sequential) ->
    gaussian(sigma).  First blur is exact fixed-point integer (so the bytes being swapped are bit-exact on any device),
    second blur is fp32; borders clamp ('nearest')."""
    n, h, w, _ = x_u8.shape
    sigma, delta, iters = CONSTANTS[profile or profile_for(h, w)]["glass_blur"][severity - 1]
    r, q = gaussian_taps_q16(sigma)
    xi = x_u8.astype(np.int64)
    idx = lambda size: np.clip(np.arange(size)[:, None] + np.arange(-r, r + 1)[None, :], 0, size - 1)
    a1 = ((xi[:, :, idx(w)] * q[None, None, None, :, None]).sum(3) + 128) >> 8                  # along x, 8 fractional bits
    a2 = (a1[:, idx(h)] * q[None, None, :, None, None]).sum(2)                                  # along y
    b = np.clip((a2 + (1 << 23)) >> 24, 0, 255).astype(np.uint8)
    dy, dx = glass_swaps(n, h, w, delta, iters, severity, seed, first_image)
    for i in range(n):
        j = 0
        img = b[i]
        for _ in range(iters):
            for hh in range(h - delta, delta, -1):
                for ww in range(w - delta, delta, -1):
                    h2, w2 = hh + int(dy[i, j]), ww + int(dx[i, j])
                    j += 1
                    t = img[hh, ww].copy(); img[hh, ww] = img[h2, w2]; img[h2, w2] = t
    _, k = gaussian_taps(sigma)
    k = k.astype(np.float32)
    y = _to_float(b)
    y1 = np.zeros_like(y)
    for t in range(2 * r + 1):
        y1 += k[t] * y[:, :, np.clip(np.arange(w) + t - r, 0, w - 1)]
    y2 = np.zeros_like(y)
    for t in range(2 * r + 1):
        y2 += k[t] * y1[:, np.clip(np.arange(h) + t - r, 0, h - 1)]
    return np.clip(y2, 0, 1).astype(np.float32)


SNOW_ANGLES = 91            # integer degrees -135..-45


def snow_field(n, h, w, loc, scale, severity, seed, first_image):
    """Bit-exact stand-in for N(loc, scale^2): 8-term Irwin-Hall on the eight 16-bit lanes of one Philox call per pixel
    (documented deviation; integer sum -> identical on CPU and GPU, so the hard threshold below cannot flip)."""
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    pix = np.arange(h * w, dtype=np.uint64)[None, :]
    xs = px.philox4x32_10(pix, img, 0, _stream("snow", severity), seed)
    ssum = px.u16_lanes(*xs).astype(np.int64).sum(-1) - 4 * 65535                    # exact integer, |.| <= 262140
    z = (ssum.astype(np.float32) * np.float32(1.0 / (65536.0 * math.sqrt(8.0 / 12.0)))).astype(np.float32)
    return (np.float32(loc) + np.float32(scale) * z).astype(np.float32).reshape(n, h, w)


def snow(x_u8, severity, seed=0, first_image=0, profile=None):
    n, h, w, _ = x_u8.shape
    loc, scale, zoom, thresh, mb_r, mb_s, blend = CONSTANTS[profile or profile_for(h, w)]["snow"][severity - 1]
    L = snow_field(n, h, w, loc, scale, severity, seed, first_image)
    # clipped zoom (bilinear, same geometry and op order as zoom_blur), then hard threshold and clip
    y0, y1, fy = _zoom_sample_axis(h, float(zoom))
    x0, x1, fx = _zoom_sample_axis(w, float(zoom))
    fy_, fx_ = fy[None, :, None], fx[None, None, :]
    top = L[:, y0][:, :, x0] * (1 - fx_) + L[:, y0][:, :, x1] * fx_
    bot = L[:, y1][:, :, x0] * (1 - fx_) + L[:, y1][:, :, x1] * fx_
    L = (top * (1 - fy_) + bot * fy_).astype(np.float32)
    L = np.where(L < np.float32(thresh), np.float32(0), L)
    L = np.clip(L, 0, 1).astype(np.float32)
    # motion blur of the layer, one integer angle in [-135, -45] per image
    img = np.arange(first_image, first_image + n, dtype=np.uint64)
    a0, _, _, _ = px.philox4x32_10(0, img, 0, _stream("snow", severity, px.KIND_AUX), seed)
    aidx = (a0 % np.uint32(SNOW_ANGLES)).astype(np.int64)
    B = np.empty_like(L)
    for i in range(n):
        dys, dxs, ws = motion_taps(int(mb_r), float(mb_s), int(aidx[i]) - 135)
        keep = len(ws)
        for j, (dy, dx) in enumerate(zip(dys, dxs)):
            if abs(dy) >= h or abs(dx) >= w:
                keep = j
                break
        B[i] = _apply_taps(L[i:i + 1, :, :, None], dys[:keep], dxs[:keep], ws[:keep], "clamp")[0, :, :, 0]
    x = _to_float(x_u8)
    gray = (np.float32(0.299) * x[..., 0] + np.float32(0.587) * x[..., 1] + np.float32(0.114) * x[..., 2]).astype(np.float32)
    lift = (gray * np.float32(1.5) + np.float32(0.5))[..., None]
    x = np.float32(blend) * x + np.float32(1 - blend) * np.maximum(x, lift)
    return np.clip(x + B[..., None] + B[:, ::-1, ::-1, None], 0, 1).astype(np.float32)


def _reflect_sym(i, n):
    """scipy 'reflect' (half-sample symmetric): ... c b a | a b c ... valid for any integer i."""
    i = np.mod(i, 2 * n)
    return np.where(i >= n, 2 * n - 1 - i, i)


def elastic_gauss_taps(sigma):
    r = int(3.0 * float(sigma) + 0.5)
    if sigma <= 1e-6:
        return 0, np.ones(1, dtype=np.float32)
    xs = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (xs / float(sigma)) ** 2)
    return r, (k / k.sum()).astype(np.float32)


def elastic_params(h, w, c, profile=None):
    """alpha, sigma, affine magnitude.  make_imagenet_c scales its constants by the literal 244 on 224-pixel images
    (c = [(244 * 2, 244 * 0.7, 244 * 0.1), ...]); make_cifar_c by IMSIZE = 32."""
    S = 244.0 * min(h, w) / 224.0 if (profile or profile_for(h, w)) == "imagenet" else float(min(h, w))
    return float(c[0]) * S, float(c[1]) * S, float(c[2]) * S          # alpha, sigma, affine magnitude


def elastic_affine(n, h, w, mag, severity, seed, first_image):
    """Per image the INVERSE map dst -> src as (p0 [2], q0 [2], A [2,2]): src = p0 + A (dst - q0), all fp32.
    pts1 -> pts2 = pts1 + U(-mag, mag) as in make_imagenet_c.elastic_transform."""
    img = np.arange(first_image, first_image + n, dtype=np.uint64)
    a = px.philox4x32_10(0, img, 0, _stream("elastic_transform", severity, px.KIND_AUX), seed)
    b = px.philox4x32_10(1, img, 0, _stream("elastic_transform", severity, px.KIND_AUX), seed)
    u = np.stack([px.u32_to_uniform(v) for v in (a[0], a[1], a[2], a[3], b[0], b[1])], -1)          # [n, 6]
    jit = ((np.float32(2.0) * u - np.float32(1.0)) * np.float32(mag)).astype(np.float32).reshape(n, 3, 2)
    c0, c1, sq = np.float32(h // 2), np.float32(w // 2), np.float32(min(h, w) // 3)
    p = np.array([[c0 + sq, c1 + sq], [c0 + sq, c1 - sq], [c0 - sq, c1 - sq]], dtype=np.float32)        # (x, y) points
    q = (p[None] + jit).astype(np.float32)
    e1, e2 = q[:, 1] - q[:, 0], q[:, 2] - q[:, 0]                        # columns of Q
    d1, d2 = p[1] - p[0], p[2] - p[0]                                    # columns of P
    det = (e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]).astype(np.float32)
    inv = np.float32(1.0) / det
    # Q^-1 = 1/det [[e2y, -e2x], [-e1y, e1x]];  A = P Q^-1
    qi00, qi01 = e2[:, 1] * inv, -e2[:, 0] * inv
    qi10, qi11 = -e1[:, 1] * inv, e1[:, 0] * inv
    A = np.empty((n, 2, 2), dtype=np.float32)
    A[:, 0, 0] = d1[0] * qi00 + d2[0] * qi10
    A[:, 0, 1] = d1[0] * qi01 + d2[0] * qi11
    A[:, 1, 0] = d1[1] * qi00 + d2[1] * qi10
    A[:, 1, 1] = d1[1] * qi01 + d2[1] * qi11
    return p[0], q[:, 0], A


def elastic_transform(x_u8, severity, seed=0, first_image=0, profile=None):
    """Random affine warp (bilinear, reflect-101) followed by a Gaussian-smoothed random displacement field sampled
    bilinearly with symmetric reflection; own float formulation of make_imagenet_c.elastic_transform."""
    n, h, w, _ = x_u8.shape
    alpha, sigma, mag = elastic_params(h, w, CONSTANTS[profile or profile_for(h, w)]["elastic_transform"][severity - 1], profile)
    x = _to_float(x_u8)
    p0, q0, A = elastic_affine(n, h, w, mag, severity, seed, first_image)
    X, Y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    warped = np.empty_like(x)
    for i in range(n):
        dxs, dys = X - q0[i, 0], Y - q0[i, 1]
        sx = (p0[0] + (A[i, 0, 0] * dxs + A[i, 0, 1] * dys)).astype(np.float32)
        sy = (p0[1] + (A[i, 1, 0] * dxs + A[i, 1, 1] * dys)).astype(np.float32)
        x0, y0 = np.floor(sx), np.floor(sy)
        fx, fy = (sx - x0)[..., None], (sy - y0)[..., None]
        x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
        r101 = lambda v, m: _reflect101(np.clip(v, -(m - 1), 2 * (m - 1)), m)          # far-off coordinates clamp to one reflection
        xa, xb, ya, yb = r101(x0i, w), r101(x0i + 1, w), r101(y0i, h), r101(y0i + 1, h)
        im = x[i]
        top = im[ya, xa] * (1 - fx) + im[ya, xb] * fx
        bot = im[yb, xa] * (1 - fx) + im[yb, xb] * fx
        warped[i] = (top * (1 - fy) + bot * fy).astype(np.float32)
    # displacement fields
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    pix = np.arange(h * w, dtype=np.uint64)[None, :]
    r0, r1, _, _ = px.philox4x32_10(pix, img, 0, _stream("elastic_transform", severity), seed)
    U = np.stack([px.u32_to_uniform(r0), px.u32_to_uniform(r1)], 1).reshape(n, 2, h, w)
    U = (np.float32(2.0) * U - np.float32(1.0)).astype(np.float32)
    r, k = elastic_gauss_taps(sigma)
    P1 = np.zeros_like(U)
    for t in range(2 * r + 1):
        P1 += k[t] * U[:, :, :, _reflect_sym(np.arange(w) + t - r, w)]
    D = np.zeros_like(U)
    for t in range(2 * r + 1):
        D += k[t] * P1[:, :, _reflect_sym(np.arange(h) + t - r, h)]
    D = (D * np.float32(alpha)).astype(np.float32)
    out = np.empty_like(x)
    for i in range(n):
        sx, sy = (X + D[i, 0]).astype(np.float32), (Y + D[i, 1]).astype(np.float32)
        x0, y0 = np.floor(sx), np.floor(sy)
        fx, fy = (sx - x0)[..., None], (sy - y0)[..., None]
        x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
        xa, xb, ya, yb = _reflect_sym(x0i, w), _reflect_sym(x0i + 1, w), _reflect_sym(y0i, h), _reflect_sym(y0i + 1, h)
        im = warped[i]
        top = im[ya, xa] * (1 - fx) + im[ya, xb] * fx
        bot = im[yb, xa] * (1 - fx) + im[yb, xb] * fx
        out[i] = (top * (1 - fy) + bot * fy).astype(np.float32)
    return np.clip(out, 0, 1).astype(np.float32)


GENERATORS = {
    "gaussian_noise": gaussian_noise, "shot_noise": shot_noise, "impulse_noise": impulse_noise,
    "defocus_blur": defocus_blur, "motion_blur": motion_blur, "zoom_blur": zoom_blur,
    "brightness": brightness, "contrast": contrast, "fog": fog, "pixelate": pixelate,
    "jpeg_compression": jpeg_compression, "frost": frost, "glass_blur": glass_blur, "snow": snow,
    "elastic_transform": elastic_transform,
}


def corrupt(x_u8, name, severity, seed=0, first_image=0, profile=None):
    """float32 [N,H,W,3] in [0,1].  name None / 'clean' -> x/255."""
    if name in (None, "clean", "none"):
        return _to_float(x_u8)
    if name not in GENERATORS:
        raise NotImplementedError(f"corruption '{name}' has no oracle yet")
    if not 1 <= int(severity) <= 5:
        raise ValueError("severity must be 1..5")
    return GENERATORS[name](x_u8, int(severity), seed=seed, first_image=first_image, profile=profile)


MEAN_STD = {
    "imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "cifar": ((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616)),
}


def normalize(x, mean, std):
    """(x - mean_c) * (1/std_c) in fp32 -- same two ops (sub, mul by fp32 reciprocal) as the kernel."""
    m = np.asarray(mean, dtype=np.float32)
    inv = (np.float32(1.0) / np.asarray(std, dtype=np.float32)).astype(np.float32)
    return ((x - m) * inv).astype(np.float32)


def to_bf16(x):
    """Round-to-nearest-even fp32 -> bf16, returned as fp32 values."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((b >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    out = ((b + r) & np.uint32(0xFFFF0000)).view(np.float32)
    return np.where(np.isnan(x), x, out).astype(np.float32)
