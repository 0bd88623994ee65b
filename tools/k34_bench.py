"""K3 / K3+K4 timing at the sweep shapes (CUDA events; logits > L2).  python tools/k34_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fav
from fav.sweep import MetricsAccumulator

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


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


for (n, T, Cc, scale) in ((2_000_000, 20, 10, 1.0), (2_000_000, 20, 10, 8.0), (1_000_000, 30, 10, 1.0), (4_000_000, 1, 10, 1.0),
                          (200_000, 1, 1000, 1.0), (20_000, 30, 1000, 1.0)):
    clf = fav.VisionClassifier("resnet18", Cc, (32, 32))
    logits = torch.randn((n, T, Cc), dtype=torch.float32, device="cuda") * scale
    labels = torch.randint(0, Cc, (n,), dtype=torch.int32, device="cuda")
    acc = MetricsAccumulator(clf, 1)
    gb = n * (T * Cc * 4 + 4) / 1e9
    ms_f = timeit(lambda: acc.add_logits(0, logits, labels, 0.9), reps=5)
    outs = {k: torch.empty(n, dtype=d, device="cuda") for k, d in (("confidence", torch.float32), ("entropy", torch.float32),
            ("mutual_information", torch.float32), ("pred", torch.int32), ("failure_flag", torch.uint8))}
    ms_e = timeit(lambda: clf.epilogue(logits, labels, 0.9), reps=5)
    print(f"n={n} T={T} C={Cc} scale={scale}: fused K3+K4 {ms_f:7.3f} ms {100 * gb / ms_f * 1e3 / PEAK:5.1f}% of HBM | K3 only (+17 B/sample out) {ms_e:7.3f} ms "
          f"{100 * (gb + n * 17e-9) / ms_e * 1e3 / PEAK:5.1f}%")
    del logits, labels, clf, acc
    torch.cuda.empty_cache()
