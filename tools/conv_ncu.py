"""One sweep step for `ncu --set full` of its tensor-core conv launches (C2: 18 launches, C3 / C4: 50):
  ncu --set full --clock-control none -k regex:'conv' -c 18 -o /tmp/conv_full python tools/conv_ncu.py 4096 20
  ncu --set full --clock-control none -k regex:'conv' -c 50 -o /tmp/conv_full_c3 python tools/conv_ncu.py 256 1 resnet50 224"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fav.sweep import CorruptionSweep, SweepConfig

block, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 20
model = sys.argv[3] if len(sys.argv) > 3 else "resnet18"
hw = int(sys.argv[4]) if len(sys.argv) > 4 else 32
ncls = 10 if hw <= 64 else 1000
sw = CorruptionSweep(SweepConfig(model=model, num_classes=ncls, input_hw=(hw, hw), T=T, logit_gain=8.0, block=block,
                                 corruptions=("gaussian_noise",), severities=(3,)))
x = torch.randint(0, 256, (block, hw, hw, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, ncls, (block,), dtype=torch.int32, device="cuda")
sw.run_item(x, y, (0, 0))
torch.cuda.synchronize()
