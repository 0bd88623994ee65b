"""One C2 sweep step (4096 images x T=20) for `ncu --set full` of its 18 tensor-core conv launches:
  ncu --set full --clock-control none --import-source on -k regex:'conv_' -c 18 -o gpurun_out/conv_full python tools/conv_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fav.sweep import CorruptionSweep, SweepConfig

block, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 20
sw = CorruptionSweep(SweepConfig(T=T, logit_gain=8.0, block=block, corruptions=("gaussian_noise",), severities=(3,)))
x = torch.randint(0, 256, (block, 32, 32, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 10, (block,), dtype=torch.int32, device="cuda")
sw.run_item(x, y, (0, 0))
torch.cuda.synchronize()
