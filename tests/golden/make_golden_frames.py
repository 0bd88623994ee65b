"""Seeded frame sequences shared by make_golden.py and the tests (no reference import)."""
import numpy as np


def frame_sequence(seed, h, w, n=12):
    """Deterministic BGR frames: noise, smooth gradient scenes, a frozen run, a dark run, a bright run."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    for i in range(n):
        kind = i % 6
        if kind == 0:
            f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            base = ((xx * 255 // max(w - 1, 1)) + (yy * 3) + 17 * i) % 256
            f = np.stack([base, (base * 2) % 256, 255 - base], -1).astype(np.uint8)
            f = np.clip(f.astype(np.int32) + rng.integers(-6, 7, f.shape), 0, 255).astype(np.uint8)
        elif kind == 2:
            f = frames[-1].copy()                                  # frozen
        elif kind == 3:
            f = rng.integers(0, 12, (h, w, 3), dtype=np.uint8)     # dark -> BLANK
        elif kind == 4:
            f = (250 + rng.integers(0, 6, (h, w, 3))).astype(np.uint8)   # over-exposed -> BLANK
        else:
            blk = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3), dtype=np.uint8)
            f = np.kron(blk, np.ones((8, 8, 1), dtype=np.uint8))[:h, :w]
        frames.append(np.ascontiguousarray(f))
    # a genuinely frozen tail to trip VISION_FROZEN (5 consecutive identical frames)
    frames += [frames[1].copy() for _ in range(7)]
    return frames
