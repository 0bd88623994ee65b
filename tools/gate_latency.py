"""C5: streaming 640x480 BGR frames, batch 1, ResNet-18 + uncertainty gate; p50/p99 latency per frame
(host frame in memory -> state dict on the host).  python tools/gate_latency.py [T]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fav

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
res = {}
for (h, w) in ((240, 320), (480, 640)):
    for TT in sorted({1, T}):
        gate = fav.UncertaintyGate(frame_hw=(h, w), T=TT, num_classes=1000, logit_gain=2.0)
        rng = np.random.default_rng(0)
        frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(16)]
        for i in range(30):
            gate.analyze_frame(frames[i % 16])
        lat = []
        for i in range(300):
            t0 = time.perf_counter()
            r = gate.analyze_frame(frames[i % 16])
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat)
        key = f"{w}x{h}_T{TT}"
        res[key] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean())}
        print(key, res[key], r["vision_status"], r["anomaly_score"])
        # signal-only gate (the reference's own computation, fused on the GPU)
    g2 = fav.UncertaintyGate(frame_hw=(h, w), use_classifier=False, score_source="signal")
    lat = []
    for i in range(330):
        t0 = time.perf_counter()
        g2.analyze_frame(frames[i % 16])
        if i >= 30:
            lat.append((time.perf_counter() - t0) * 1e3)
    res[f"{w}x{h}_signal_only"] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99))}
    print(f"{w}x{h} signal-only", res[f"{w}x{h}_signal_only"])
    try:
        import cv2  # CPU reference arithmetic (same as SignalAnalyzer.analyze_frame's OpenCV calls), for the latency comparison
        lat = []
        prev = None
        for i in range(200):
            f = frames[i % 16]
            t0 = time.perf_counter()
            gray = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
            lv = cv2.Laplacian(gray, cv2.CV_64F).var()
            mb = float(np.mean(gray))
            if prev is not None:
                md = float(np.mean(cv2.absdiff(prev, gray)))
            prev = gray.copy()
            hist = cv2.calcHist([gray], [0], None, [256], [0, 256]).flatten()
            lat.append((time.perf_counter() - t0) * 1e3)
        res[f"{w}x{h}_cpu_opencv_stats"] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99))}
        print(f"{w}x{h} cpu opencv", res[f"{w}x{h}_cpu_opencv_stats"])
    except ImportError:
        pass
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gate_latency.json", "w"), indent=1)
