"""Profiling target: a few C2 sweep steps (ResNet-18, 32x32, T=20).  python tools/step_one.py [block] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fav
from fav.sweep import CorruptionSweep, SweepConfig
block = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sw = CorruptionSweep(SweepConfig(T=20, logit_gain=8.0, block=block, corruptions=("gaussian_noise",), severities=(3,)))
sw.prepare()
x = torch.randint(0, 256, (block, 32, 32, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 10, (block,), dtype=torch.int32, device="cuda")
for _ in range(steps):
    sw.run_item(x, y, (0, 0))
torch.cuda.synchronize()
print("ok")
