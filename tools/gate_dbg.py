import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fav
h, w = 480, 640
gate = fav.UncertaintyGate(frame_hw=(h, w), T=1, num_classes=1000, logit_gain=2.0)
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(8)]
for i in range(20):
    gate.analyze_frame(frames[i % 8])
lat = []
for i in range(200):
    t0 = time.perf_counter(); gate.analyze_frame(frames[i % 8]); lat.append((time.perf_counter() - t0) * 1e3)
print("graph active:", gate._graph is not None, "failed:", gate._graph_failed, "p50 ms:", float(np.percentile(lat, 50)))
# device time of one replay
if gate._graph is not None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gate._graph.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph replay device time per frame: %.3f ms" % (e0.elapsed_time(e1) / 20))
