"""jpeg_compression -- CPU oracle: a restatement of the libjpeg baseline round trip.  TEST INFRASTRUCTURE ONLY.

Hendrycks & Dietterich's jpeg_compression is ``PIL.Image.save(buf, 'JPEG', quality=c)`` then reload
(SURVEY.md Appendix A.2).  The arithmetic lives in a third-party dependency that is not under
/root/reference: libjpeg (IJG 6b API; this image ships libjpeg-turbo behind Pillow 12.2, whose SIMD
paths are bit-identical to the C ones).  This file restates the published integer algorithm stage by
stage -- entropy coding is lossless and omitted:

  compress   jccolor.c  rgb_ycc_convert      16-bit fixed-point JFIF colour transform
             jcsample.c h2v2_downsample      4:2:0 chroma, 2x2 mean with the alternating 1,2 bias
             jfdctint.c jpeg_fdct_islow      LL&M 8x8 forward DCT, CONST_BITS 13 / PASS1_BITS 2
             jcdctmgr.c quantize             divisor = 8 * q, round half away from zero
             jcparam.c  jpeg_set_quality     Annex-K tables scaled by the quality rule, baseline clamp
  decompress jidctint.c jpeg_idct_islow      dequantise + inverse DCT + range limit
             jdsample.c h2v2_fancy_upsample  triangle filter (3/4, 1/4) with the 8 / 7 rounding biases
             jdcolor.c  ycc_rgb_convert      fixed-point YCbCr -> RGB

PINNED: tests/test_oracle.py checks this restatement BYTE FOR BYTE against Pillow's real JPEG round trip
(random, smooth and saturated images, every quality used by the severity tables, sizes that are and are
not multiples of the 16x16 MCU).  The CUDA kernel (k1_jpeg) is in turn bit-exact against both.
"""
import numpy as np

LUM = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)
CHR = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, dtype=np.int64).reshape(8, 8)

# jfdctint.c / jidctint.c constants, FIX(x) = round(x * 2^13)
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172
CONST_BITS, PASS1_BITS = 13, 2


def quant_tables(quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE)."""
    q = int(min(max(quality, 1), 100))
    scale = 5000 // q if q < 50 else 200 - 2 * q
    f = lambda base: np.clip((base * scale + 50) // 100, 1, 255)
    return f(LUM), f(CHR)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first):
    """One pass of jpeg_fdct_islow along the last axis (first: rows, PASS1 scaling; else columns)."""
    d0, d1, d2, d3, d4, d5, d6, d7 = (d[..., i] for i in range(8))
    tmp0, tmp7, tmp1, tmp6 = d0 + d7, d0 - d7, d1 + d6, d1 - d6
    tmp2, tmp5, tmp3, tmp4 = d2 + d5, d2 - d5, d3 + d4, d3 - d4
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    out = [None] * 8
    if first:
        out[0] = (tmp10 + tmp11) << PASS1_BITS
        out[4] = (tmp10 - tmp11) << PASS1_BITS
        sh = CONST_BITS - PASS1_BITS
    else:
        out[0] = _descale(tmp10 + tmp11, PASS1_BITS)
        out[4] = _descale(tmp10 - tmp11, PASS1_BITS)
        sh = CONST_BITS + PASS1_BITS
    z1 = (tmp12 + tmp13) * F_0_541196100
    out[2] = _descale(z1 + tmp13 * F_0_765366865, sh)
    out[6] = _descale(z1 + tmp12 * (-F_1_847759065), sh)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * F_1_175875602
    tmp4, tmp5, tmp6, tmp7 = tmp4 * F_0_298631336, tmp5 * F_2_053119869, tmp6 * F_3_072711026, tmp7 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560), z4 * (-F_0_390180644)
    z3, z4 = z3 + z5, z4 + z5
    out[7] = _descale(tmp4 + z1 + z3, sh)
    out[5] = _descale(tmp5 + z2 + z4, sh)
    out[3] = _descale(tmp6 + z2 + z3, sh)
    out[1] = _descale(tmp7 + z1 + z4, sh)
    return np.stack(out, -1)


def fdct_islow(b):
    """int64 [..., 8(y), 8(x)] level-shifted samples -> coefficients scaled by 8, [..., v, u]."""
    t = _fdct_1d(b, True)                                      # rows
    return np.swapaxes(_fdct_1d(np.swapaxes(t, -1, -2), False), -1, -2)   # columns


def _idct_1d(c, first):
    """One pass of jpeg_idct_islow along the last axis (first: columns -> workspace; else rows -> samples - 128)."""
    c0, c1, c2, c3, c4, c5, c6, c7 = (c[..., i] for i in range(8))
    z2, z3 = c2, c6
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    tmp0, tmp1 = (c0 + c4) << CONST_BITS, (c0 - c4) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = c7, c5, c3, c1
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175875602
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298631336, tmp1 * F_2_053119869, tmp2 * F_3_072711026, tmp3 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560), z4 * (-F_0_390180644)
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    out = [tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3]
    return np.stack([_descale(o, sh) for o in out], -1)


def range_limit(v):
    """jdmaster.c prepare_range_limit_table as seen by the IDCT: index (v & 1023) of a table centred on 128."""
    v = v & 1023
    return np.where(v < 128, v + 128, np.where(v < 512, 255, np.where(v < 896, 0, v - 896)))


def idct_islow(c):
    """int64 [..., v, u] dequantised coefficients -> uint8-range samples [..., y, x]."""
    ws = np.swapaxes(_idct_1d(np.swapaxes(c, -1, -2), True), -1, -2)      # columns
    return range_limit(_idct_1d(ws, False))                               # rows


def quantize(F, Q):
    """jcdctmgr.c: divisor 8 q, round half away from zero; then the decoder's dequantisation (coef * q)."""
    d = Q * 8
    q = np.sign(F) * ((np.abs(F) + (d >> 1)) // d)
    return q * Q


def _blocks(p):
    n, ph, pw = p.shape
    return p.reshape(n, ph // 8, 8, pw // 8, 8).transpose(0, 1, 3, 2, 4)


def _unblocks(b):
    n, by, bx = b.shape[:3]
    return b.transpose(0, 1, 3, 2, 4).reshape(n, by * 8, bx * 8)


def _codec_plane(p, Q):
    return _unblocks(idct_islow(quantize(fdct_islow(_blocks(p) - 128), Q)))


def h2v2_downsample(p):
    """jcsample.c h2v2_downsample: bias 1, 2, 1, 2, ... along each output row."""
    s = p[:, 0::2, 0::2] + p[:, 0::2, 1::2] + p[:, 1::2, 0::2] + p[:, 1::2, 1::2]
    bias = 1 + (np.arange(s.shape[2]) & 1)
    return (s + bias[None, None, :]) >> 2


def h2v2_fancy_upsample(c, ch, cw):
    """jdsample.c h2v2_fancy_upsample on the real downsampled extent [ch, cw] (context rows beyond the image replicate
    the first / last real row, jdmainct.c) -> [2 ch, 2 cw]."""
    c = c[:, :ch, :cw]
    up = np.concatenate([c[:, :1], c[:, :-1]], 1)               # row r-1 (row 0 for r = 0)
    dn = np.concatenate([c[:, 1:], c[:, -1:]], 1)               # row r+1 (last row for r = ch-1)
    out = np.empty((c.shape[0], 2 * ch, 2 * cw), dtype=np.int64)
    for v, far in ((0, up), (1, dn)):
        cs = 3 * c + far                                        # column sums
        left = np.concatenate([cs[:, :, :1], cs[:, :, :-1]], 2)
        right = np.concatenate([cs[:, :, 1:], cs[:, :, -1:]], 2)
        out[:, v::2, 0::2] = (3 * cs + left + 8) >> 4           # first column: (4 cs + 8) >> 4, same expression
        out[:, v::2, 1::2] = (3 * cs + right + 7) >> 4          # last column:  (4 cs + 7) >> 4
    return out


def jpeg_roundtrip_u8(x_u8, quality):
    """uint8 [N,H,W,3] RGB -> uint8 [N,H,W,3]: what PIL's save(quality=q) + reload returns."""
    n, h, w, _ = x_u8.shape
    H, W = (h + 15) // 16 * 16, (w + 15) // 16 * 16
    ch, cw = (h + 1) // 2, (w + 1) // 2
    # columns: the right edge replicates in the INPUT of the downsampler (jcsample.c expand_right_edge); rows: an odd last
    # row is doubled in the conversion buffer, then the DOWNSAMPLED planes are padded to the iMCU height by replicating
    # their last real row (jcprepct.c pre_process_data / expand_bottom_edge) -- not the same thing for even heights
    x = np.pad(x_u8, ((0, 0), (0, 2 * ch - h), (0, W - w), (0, 0)), mode="edge").astype(np.int64)
    R, G, B = x[..., 0], x[..., 1], x[..., 2]
    Y = (19595 * R + 38470 * G + 7471 * B + 32768) >> 16
    Cb = (-11059 * R - 21709 * G + 32768 * B + 8388608 + 32767) >> 16
    Cr = (32768 * R - 27439 * G - 5329 * B + 8388608 + 32767) >> 16
    pad_rows = lambda p, rows: np.pad(p, ((0, 0), (0, rows - p.shape[1]), (0, 0)), mode="edge")
    QL, QC = quant_tables(quality)
    Y = _codec_plane(pad_rows(Y, H), QL)
    Cb = _codec_plane(pad_rows(h2v2_downsample(Cb), H // 2), QC)
    Cr = _codec_plane(pad_rows(h2v2_downsample(Cr), H // 2), QC)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    cb = h2v2_fancy_upsample(Cb, ch, cw)[:, :h, :w] - 128
    cr = h2v2_fancy_upsample(Cr, ch, cw)[:, :h, :w] - 128
    Y = Y[:, :h, :w]
    R = Y + ((91881 * cr + 32768) >> 16)
    G = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16)
    B = Y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([R, G, B], -1), 0, 255).astype(np.uint8)


def pil_roundtrip_u8(x_u8, quality):
    """The real thing ('PIL JPEG quality c'): Pillow save + reload, image by image."""
    import io
    from PIL import Image
    out = np.empty_like(x_u8)
    for i in range(x_u8.shape[0]):
        buf = io.BytesIO()
        Image.fromarray(x_u8[i]).save(buf, "JPEG", quality=int(quality))
        out[i] = np.array(Image.open(buf))
    return out
