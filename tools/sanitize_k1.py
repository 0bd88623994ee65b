"""Small-footprint run of every K1 kernel (both profiles, ragged frames, RGB / BGR, bf16 / fp32), K3+K4 (direct + histogram
variants) and a tiny forward, for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_k1.py
  compute-sanitizer --tool racecheck python tools/sanitize_k1.py k1only
(compute-sanitizer is closed on the round-2 GPU pool -- runs under it left GPUs needing a reset -- so this was not run there; the
script doubles as a no-assert smoke run of every K1 / K3+K4 path: `python tools/sanitize_k1.py`.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fav
from fav.sweep import MetricsAccumulator

k1only = len(sys.argv) > 1 and sys.argv[1] == "k1only"
rng = np.random.default_rng(0)
for (h, w, n, ncls) in ((32, 32, 20, 10), (40, 48, 5, 10), (224, 224, 2, 1000), (120, 160, 2, 1000)):
    clf = fav.VisionClassifier("resnet18", ncls, (h, w))
    x = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
    for name in (None,) + tuple(fav.IMPLEMENTED):
        for sev in ((0,) if name is None else (1, 5)):
            cfg = fav.CorruptionConfig(name, sev)
            for bgr in (False, True):
                for f32 in (False, True):
                    clf.corrupt_normalize(x, cfg, 3, 7, bgr=bgr, out_f32=f32, normalize=not f32)
    torch.cuda.synchronize()
    print("k1 ok", h, w, flush=True)
    if not k1only and h == 32:
        for (ns, T) in ((100, 20), (40000, 7), (33, 1)):
            logits = torch.randn((ns, T, 10), device="cuda") * 3
            labels = torch.randint(-1, 11, (ns,), dtype=torch.int32, device="cuda")       # includes invalid labels
            acc = MetricsAccumulator(clf, 1)
            acc.add_logits(0, logits, labels, 0.5)
        out = clf.uncertainty(x, fav.CorruptionConfig("shot_noise", 2), T=3, labels=torch.zeros(n, dtype=torch.int32), seed=1)
        torch.cuda.synchronize()
        print("k34 + forward ok", flush=True)
print("done")
