"""Tap stencils (defocus_blur, motion_blur): current kernels against the same kernel with the tap-list loop
('k1_list_stencil') and against the round-1 path ('k1_legacy'), CUDA events, inputs > L2.
python tools/k1_stencil_bench.py -> table on stdout"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav
from fav import _lib

PEAK = 6512.3


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for hw, n, ncls in ((32, 65536, 10), (224, 1536, 1000)):
    clf = fav.VisionClassifier("resnet18", ncls, (hw, hw))
    x = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, hw, hw, 3), dtype=torch.bfloat16, device="cuda")
    gb = 9.0 * hw * hw * n / 1e9
    for name in ("defocus_blur", "motion_blur"):
        for sev in (1, 2, 3, 4, 5):
            cfg = fav.CorruptionConfig(name, sev)
            ms = {}
            for opt in (None, b"k1_list_stencil", b"k1_legacy"):
                if opt:
                    _lib.check(clf.lib.fav_set_option(clf.handle.h, opt, 1), "fav_set_option")
                ms[opt] = timeit(lambda: clf.corrupt_normalize(x, cfg, 0, 0, out=out))
                if opt:
                    _lib.check(clf.lib.fav_set_option(clf.handle.h, opt, 0), "fav_set_option")
            pct = lambda t: 100 * gb / t * 1e3 / PEAK
            print(f"{name:13s} s{sev} {hw:3d}x{hw:<3d} n={n:6d}  now {ms[None]:8.3f} ms ({pct(ms[None]):5.1f}% of HBM)   "
                  f"tap list {ms[b'k1_list_stencil']:8.3f} ms ({pct(ms[b'k1_list_stencil']):5.1f}%)   "
                  f"round 1 {ms[b'k1_legacy']:8.3f} ms ({pct(ms[b'k1_legacy']):5.1f}%)")
    del x, out, clf
    torch.cuda.empty_cache()
