"""Whole-sweep throughput of one configuration: every (corruption, severity) cell over one resident image block,
CUDA-event timed (K1 + forward + K3/K4 of all 75 cells).  python tools/sweep_time.py [model] [hw] [block] [T] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav
from fav.sweep import CorruptionSweep, SweepConfig

model = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 224
block = int(sys.argv[3]) if len(sys.argv) > 3 else 256
T = int(sys.argv[4]) if len(sys.argv) > 4 else 1
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ncls = 10 if hw <= 64 else 1000
sw = CorruptionSweep(SweepConfig(model=model, num_classes=ncls, input_hw=(hw, hw), T=T, logit_gain=8.0, block=block))
sw.prepare()
x = torch.randint(0, 256, (block, hw, hw, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, ncls, (block,), dtype=torch.int32, device="cuda")
cells = len(sw.cells)
for ci in range(cells):
    sw.run_item(x, y, (ci, 0))
torch.cuda.synchronize()
per_cell = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(cells + 1)]
tot = 0.0
for _ in range(reps):
    ev[0].record()
    for ci in range(cells):
        sw.run_item(x, y, (ci, 0))
        ev[ci + 1].record()
    torch.cuda.synchronize()
    tot += ev[0].elapsed_time(ev[cells])
    per_cell = [ev[i].elapsed_time(ev[i + 1]) for i in range(cells)]
ms = tot / reps
print(f"{model} {hw}x{hw} T={T} block={block}: {cells} cells in {ms:.1f} ms -> {block * cells / ms * 1e3:.0f} evals/s over the whole sweep")
slow = sorted(range(cells), key=lambda i: -per_cell[i])[:8]
print("slowest cells (ms per block): " + ", ".join(f"{sw.cells[i].name} s{sw.cells[i].severity} {per_cell[i]:.2f}" for i in slow))
print(f"median cell {sorted(per_cell)[cells // 2]:.2f} ms")
