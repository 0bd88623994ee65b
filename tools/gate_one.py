"""Profiling target: a few frames through the streaming gate (eager launches).  python tools/gate_one.py [h w T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fav
h = int(sys.argv[1]) if len(sys.argv) > 1 else 480
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gate = fav.UncertaintyGate(frame_hw=(h, w), T=T, num_classes=1000, logit_gain=2.0, use_graph=False)
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(4)]
for i in range(6):
    r = gate.analyze_frame(frames[i % 4])
print(r["vision_status"], r["anomaly_score"])
