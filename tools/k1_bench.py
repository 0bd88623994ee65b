"""HBM roofline of the K1 corruption+normalize kernels and the K3+K4 epilogue (CUDA events, inputs > L2).
python tools/k1_bench.py  -> prints a table and writes gpurun_out/k1_roofline.json"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav
from fav import _lib
from fav.sweep import MetricsAccumulator

PEAK = 6512.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


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


rows = []
for hw, n, ncls in ((32, 65536, 10), (224, 1536, 1000)):
    clf = fav.VisionClassifier("resnet18", ncls, (hw, hw))
    x = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device="cuda")       # 201 / 231 MB > L2
    out = torch.empty((n, hw, hw, 3), dtype=torch.bfloat16, device="cuda")
    names = [None] + list(fav.IMPLEMENTED)
    for name in names:
        for sev in ((0,) if name is None else (1, 5)):
            cfg = fav.CorruptionConfig(name, sev)
            ms = timeit(lambda: clf.corrupt_normalize(x, cfg, 0, 0, out=out))
            gb = 9.0 * hw * hw * n / 1e9
            rows.append({"kernel": f"K1 {name or 'clean'} s{sev}", "hw": hw, "n": n, "ms": ms, "GBps": gb / (ms * 1e-3),
                         "frac_of_measured_hbm": gb / (ms * 1e-3) / PEAK})
            print(f"K1 {str(name or 'clean'):16s} s{sev} {hw:3d}x{hw:<3d} n={n:6d} {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s  {100 * gb / ms * 1e3 / PEAK:5.1f}% of measured HBM")
    del x, out, clf
    torch.cuda.empty_cache()

clf = fav.VisionClassifier("resnet18", 10, (32, 32))
for (n, T, Cc) in ((2_000_000, 20, 10), (200_000, 1, 1000), (20_000, 30, 1000)):
    c2 = clf if Cc == 10 else fav.VisionClassifier("resnet18", Cc, (32, 32))
    logits = torch.randn((n, T, Cc), dtype=torch.float32, device="cuda")
    labels = torch.randint(0, Cc, (n,), dtype=torch.int32, device="cuda")
    acc = MetricsAccumulator(c2, 1)
    ms = timeit(lambda: acc.add_logits(0, logits, labels, 0.9), reps=5)
    gb = n * (T * Cc * 4 + 4) / 1e9
    rows.append({"kernel": f"K3+K4 T={T} C={Cc}", "n": n, "ms": ms, "GBps": gb / (ms * 1e-3), "frac_of_measured_hbm": gb / (ms * 1e-3) / PEAK})
    print(f"K3+K4 fused  n={n} T={T} C={Cc}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s  {100 * gb / ms * 1e3 / PEAK:5.1f}% of measured HBM")
    del logits, labels, acc
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_hbm_gbs": PEAK, "rows": rows}, open("gpurun_out/k1_roofline.json", "w"), indent=1)
