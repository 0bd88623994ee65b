"""jpeg_compression -- CPU oracle of an integer baseline-JPEG round trip.  TEST INFRASTRUCTURE ONLY.

Hendrycks & Dietterich's jpeg_compression is `PIL.Image.save(quality=c)` then reload.  libjpeg's
exact arithmetic (jfdctint / jidctint / fancy upsampling) is not reproducible from memory, so this
path defines its own *integer* codec with the same structure -- JFIF colour transform (libjpeg's
16-bit fixed-point constants), 4:2:0 chroma (2x2 mean / replication), 8x8 DCT with a 13-bit
fixed-point orthonormal cosine matrix, Annex-K tables scaled by the libjpeg quality rule, round-to-
nearest quantisation -- and everything is integer, so the CUDA kernel is BIT-EXACT against it.
Entropy coding is lossless and therefore omitted.  tests/test_oracle.py checks that the result
stays close to PIL's real JPEG at the same quality (PARITY UNPINNED by the reference).
"""
import numpy as np

LUM = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)
CHR = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, dtype=np.int64).reshape(8, 8)


def quant_tables(quality):
    q = int(min(max(quality, 1), 100))
    scale = 5000 // q if q < 50 else 200 - 2 * q
    f = lambda base: np.clip((base * scale + 50) // 100, 1, 255)
    return f(LUM), f(CHR)


def dct_matrix():
    """T[u][x] = round(2^13 * c(u) * cos((2x+1) u pi / 16)), c(0) = sqrt(1/8), c(u>0) = 1/2."""
    u = np.arange(8)[:, None].astype(np.float64)
    x = np.arange(8)[None, :].astype(np.float64)
    c = np.where(u == 0, np.sqrt(1.0 / 8.0), 0.5)
    return np.rint(8192.0 * c * np.cos((2 * x + 1) * u * np.pi / 16.0)).astype(np.int64)


def _codec_blocks(f, Q):
    """f int64 [..., 8, 8] level-shifted samples -> reconstructed samples (same shape)."""
    T = dct_matrix()
    t1 = (np.einsum("ux,...yx->...yu", T, f) + 512) >> 10                 # rows,   x8
    F = (np.einsum("vy,...yu->...vu", T, t1) + 4096) >> 13                # cols,   8 * F_true
    Q8 = Q * 8
    q = np.sign(F) * ((np.abs(F) + Q8 // 2) // Q8)
    Fd = q * Q                                                            # F_true'
    t = (np.einsum("vy,...vu->...yu", T, Fd) + 1024) >> 11                # cols^T, x4
    return (np.einsum("ux,...yu->...yx", T, t) + 16384) >> 15             # rows^T, x1


def jpeg_roundtrip_u8(x_u8, quality):
    """uint8 [N,H,W,3] RGB -> uint8 [N,H,W,3]."""
    n, h, w, _ = x_u8.shape
    H, W = (h + 15) // 16 * 16, (w + 15) // 16 * 16
    x = np.pad(x_u8, ((0, 0), (0, H - h), (0, W - w), (0, 0)), mode="edge").astype(np.int64)
    R, G, B = x[..., 0], x[..., 1], x[..., 2]
    Y = (19595 * R + 38470 * G + 7471 * B + 32768) >> 16
    Cb = (-11059 * R - 21709 * G + 32768 * B + 8388608 + 32767) >> 16
    Cr = (32768 * R - 27439 * G - 5329 * B + 8388608 + 32767) >> 16
    sub = lambda p: (p[:, 0::2, 0::2] + p[:, 0::2, 1::2] + p[:, 1::2, 0::2] + p[:, 1::2, 1::2] + 2) >> 2
    Cb, Cr = sub(Cb), sub(Cr)
    QL, QC = quant_tables(quality)

    def plane(p, Q):
        ph, pw = p.shape[1:]
        b = p.reshape(n, ph // 8, 8, pw // 8, 8).transpose(0, 1, 3, 2, 4) - 128
        r = np.clip(_codec_blocks(b, Q) + 128, 0, 255)
        return r.transpose(0, 1, 3, 2, 4).reshape(n, ph, pw)

    Y, Cb, Cr = plane(Y, QL), plane(Cb, QC), plane(Cr, QC)
    up = lambda p: np.repeat(np.repeat(p, 2, axis=1), 2, axis=2)
    cb, cr = up(Cb) - 128, up(Cr) - 128
    R = Y + ((91881 * cr + 32768) >> 16)
    G = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16)
    B = Y + ((116130 * cb + 32768) >> 16)
    out = np.clip(np.stack([R, G, B], -1), 0, 255).astype(np.uint8)
    return out[:, :h, :w]
