"""One launch of every K1 kernel north_star names (+ K3/K4) at the bench's own shapes, for `ncu --set full`:
  ncu --set full --clock-control none --import-source on -k regex:'k1_|k34' -o gpurun_out/k1_full python tools/k1_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav
from fav.sweep import MetricsAccumulator

CELLS = [(None, 0), ("gaussian_noise", 3), ("shot_noise", 1), ("shot_noise", 5), ("impulse_noise", 3), ("defocus_blur", 3),
         ("motion_blur", 3), ("brightness", 3), ("contrast", 3), ("fog", 3)]
for hw, n, ncls in ((32, 4096, 10), (224, 256, 1000)):
    clf = fav.VisionClassifier("resnet18", ncls, (hw, hw))
    x = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, hw, hw, 3), dtype=torch.bfloat16, device="cuda")
    for name, sev in CELLS:
        clf.corrupt_normalize(x, fav.CorruptionConfig(name, sev), 0, 0, out=out)
    torch.cuda.synchronize()
    if hw == 32:
        logits = torch.randn((n, 20, 10), dtype=torch.float32, device="cuda") * 3
        labels = torch.randint(0, 10, (n,), dtype=torch.int32, device="cuda")
        acc = MetricsAccumulator(clf, 1)
        acc.add_logits(0, logits, labels, 0.9)
        torch.cuda.synchronize()
