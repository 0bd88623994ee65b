"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11) in vectorised numpy.

Oracle side of the RNG contract.  Test infrastructure only.

Contract (shared with failure-aware-vision_b200/csrc/philox.cuh):
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (c0, c1, c2, c3) = (chunk index inside the image, GLOBAL image index,
             sub-draw / MC pass t, stream id)
  so every draw depends only on (seed, global image index, position) and never on
  batch size or on how images are partitioned over GPUs.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# stream ids (c3).  kind << 16 | a << 8 | b
KIND_IMAGES, KIND_LABELS, KIND_CORRUPT, KIND_DROPOUT, KIND_AUX = 1, 2, 3, 4, 5


def stream_id(kind, a=0, b=0):
    return (int(kind) << 16) | (int(a) << 8) | int(b)


def philox4x32_10(c0, c1, c2, c3, seed):
    """All counters broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)))
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for r in range(10):
        if r > 0:
            k0 = (k0 + W0) & 0xFFFFFFFF
            k1 = (k1 + W1) & 0xFFFFFFFF
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u32_to_uniform(x):
    """(0,1] uniform in fp32: fl(fl(x>>8) * 2^-24 + 2^-25) -- same two fp32 ops as the kernel."""
    f = (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return (f + np.float32(2.0 ** -25)).astype(np.float32)


def box_muller(xa, xb):
    """Two u32 arrays -> two fp32 standard normals (cos branch, sin branch)."""
    u1 = u32_to_uniform(xa)
    u2 = u32_to_uniform(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def u16_lanes(x0, x1, x2, x3):
    """4 u32 -> 8 u16 lanes; lane 2i = low half of x_i, lane 2i+1 = high half. Shape (..., 8)."""
    outs = []
    for x in (x0, x1, x2, x3):
        outs.append((x & np.uint32(0xFFFF)).astype(np.uint16))
        outs.append((x >> np.uint32(16)).astype(np.uint16))
    return np.stack(outs, axis=-1)


def synthetic_images(n, h, w, seed=0, first_image=0):
    """uint8 [n,h,w,3] i.i.d. uniform bytes; 16 bytes per Philox call (x0 low byte first)."""
    per = h * w * 3
    nch = (per + 15) // 16
    img = np.arange(first_image, first_image + n, dtype=np.uint64)[:, None]
    ch = np.arange(nch, dtype=np.uint64)[None, :]
    xs = philox4x32_10(ch, img, 0, stream_id(KIND_IMAGES), seed)
    words = np.stack(xs, axis=-1)                      # [n, nch, 4] u32
    by = words.view(np.uint8) if words.dtype.byteorder != '>' else None
    by = np.ascontiguousarray(words).view(np.uint8).reshape(n, nch * 16)[:, :per]
    return by.reshape(n, h, w, 3).copy()


def synthetic_labels(n, num_classes, seed=0, first_image=0):
    img = np.arange(first_image, first_image + n, dtype=np.uint64)
    x0, _, _, _ = philox4x32_10(0, img, 0, stream_id(KIND_LABELS), seed)
    return (x0 % np.uint32(num_classes)).astype(np.int32)
