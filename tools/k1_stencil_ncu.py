"""One launch of the defocus / motion stencil per shape and severity, for `ncu --set full`:
  ncu --set full --clock-control none -k regex:'k1_taps' -o /tmp/ncu/k1_taps python tools/k1_stencil_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav

for hw, n, ncls in ((32, 8192, 10), (224, 256, 1000)):
    clf = fav.VisionClassifier("resnet18", ncls, (hw, hw))
    x = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, hw, hw, 3), dtype=torch.bfloat16, device="cuda")
    for name, sev in (("defocus_blur", 1), ("defocus_blur", 3), ("defocus_blur", 5), ("motion_blur", 3)):
        clf.corrupt_normalize(x, fav.CorruptionConfig(name, sev), 0, 0, out=out)
    torch.cuda.synchronize()
