"""Per-launch timing of the conv kernel inside one sweep step (C2 shape): which layers dominate and the
TFLOP/s each one reaches.  GPU only.  python tools/conv_layers.py [block] [T] [model] [hw]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fav
from fav import _lib
from fav.sweep import CorruptionSweep, SweepConfig

block = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model = sys.argv[3] if len(sys.argv) > 3 else "resnet18"
hw = int(sys.argv[4]) if len(sys.argv) > 4 else 32
ncls = 10 if hw <= 64 else 1000
cfg = SweepConfig(model=model, num_classes=ncls, input_hw=(hw, hw), T=T, logit_gain=8.0, block=block,
                  corruptions=("gaussian_noise",), severities=(3,))
sw = CorruptionSweep(cfg)
sw.prepare()
x = torch.randint(0, 256, (block, hw, hw, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, ncls, (block,), dtype=torch.int32, device="cuda")
lib, h = sw.clf.lib, sw.clf.handle.h
for _ in range(3):
    sw.run_item(x, y, (0, 0))
torch.cuda.synchronize()
lib.fav_conv_timing_enable(h, 1)
reps = 5
acc = None
for _ in range(reps):
    sw.run_item(x, y, (0, 0))
    st = (C.c_uint64 * (256 * 8))(); ns = C.c_int()
    _lib.check(lib.fav_conv_stats_read(h, st, 256, C.byref(ns)), "stats")
    stats = [list(st[8 * i: 8 * i + 8]) for i in range(ns.value)]
    gb = (C.c_float * 256)(); nb = C.c_int()
    _lib.check(lib.fav_conv_timing_read_bytes(h, gb, 256, C.byref(nb)), "bytes")
    gbytes = [gb[i] for i in range(nb.value)]
    ms = (C.c_float * 256)(); gf = (C.c_float * 256)(); n = C.c_int()
    _lib.check(lib.fav_conv_timing_read_all(h, ms, gf, 256, C.byref(n)), "read_all")
    row = [(ms[i], gf[i]) for i in range(n.value)]
    acc = row if acc is None else [(a[0] + b[0], a[1]) for a, b in zip(acc, row)]
lib.fav_conv_timing_enable(h, 0)
tot_ms = sum(a[0] for a in acc) / reps
tot_gf = sum(a[1] for a in acc)
PEAK_TF, PEAK_GB = 1422.6, 6512.3          # MEASURED_PEAKS.json: sustained cuBLAS bf16, HBM copy
floor_ms = sum(max(a[1] / PEAK_TF, gbytes[i] * 1e3 / PEAK_GB) for i, a in enumerate(acc))
print(f"{model} {hw}x{hw} block={block} T={T}: {len(acc)} conv launches, {tot_ms:.3f} ms, {tot_gf / tot_ms:.1f} TFLOP/s nominal, "
      f"{sum(gbytes) / tot_ms * 1e3:.0f} GB/s algorithmic; per-layer roofline floor (max of tensor, HBM) {floor_ms:.3f} ms = {100 * floor_ms / tot_ms:.0f}% of measured")
for i, (m, g) in enumerate(acc):
    m /= reps
    s8 = stats[i] if i < len(stats) else [0] * 8
    nc = max(1, s8[7])
    pct = lambda a, b: 100.0 * a / b if b else 0.0
    tfl, hbm = g / PEAK_TF, gbytes[i] * 1e3 / PEAK_GB       # ms
    print(f"  conv[{i:2d}] {m * 1e3:9.1f} us  {g:9.2f} GFLOP  {g / m if m > 0 else 0:8.1f} TFLOP/s  {gbytes[i] / m * 1e3 if m > 0 else 0:7.0f} GB/s  "
          f"{'hbm' if hbm > tfl else 'mma'} {100 * max(tfl, hbm) / m if m > 0 else 0:4.0f}%  {100 * m / tot_ms:5.1f}%"
          f"  | ctas {nc:4d} tma: wait-empty {pct(s8[0], s8[1]):4.0f}%  mma: wait-full {pct(s8[2], s8[4]):4.0f}% wait-tmem {pct(s8[3], s8[4]):4.0f}%"
          f"  epi: wait-acc {pct(s8[5], s8[6]):4.0f}%  cyc/cta {s8[4] / nc:9.0f}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    sw.run_item(x, y, (0, 0))
e1.record(); torch.cuda.synchronize()
print(f"whole step: {e0.elapsed_time(e1) / 10:.3f} ms  -> {block / (e0.elapsed_time(e1) / 10) * 1e3:.0f} evals/s")
