"""Profiling target: one sweep step on a ResNet-50 224x224 block.  python tools/r50_one.py [block] [T] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fav
from fav.sweep import CorruptionSweep, SweepConfig
block = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sw = CorruptionSweep(SweepConfig(model="resnet50", num_classes=1000, input_hw=(224, 224), T=T, logit_gain=8.0, block=block,
                                 corruptions=("gaussian_noise",), severities=(3,)))
sw.prepare()
x = torch.randint(0, 256, (block, 224, 224, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 1000, (block,), dtype=torch.int32, device="cuda")
for _ in range(steps):
    sw.run_item(x, y, (0, 0))
torch.cuda.synchronize()
print("ok")
