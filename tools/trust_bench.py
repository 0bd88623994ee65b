"""f4: throughput of the batched TrustEngine replay vs the reference-style Python loop (oracle port) on the host.
python tools/trust_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch, fav
from make_golden_trust_inputs import sequences
from oracle import trust as OT

tr = fav.TrustReplay()
for S, L in ((4096, 900), (65536, 900)):
    status, score = sequences(3, min(S, 256), L)
    reps = S // status.shape[0]
    status, score = np.tile(status, (reps, 1)), np.tile(score, (reps, 1))
    tr.run(status[:128], score[:128], 1 / 30)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); res = tr.run(status, score, 1 / 30, trajectory=False); t1 = time.perf_counter()
    t2 = time.perf_counter(); res = tr.run(status, score, 1 / 30, trajectory=True); t3 = time.perf_counter()
    print(f"S={S} L={L}: GPU end-to-end (H2D + kernel + D2H) final-only {S * L / (t1 - t0) / 1e6:8.1f} M ticks/s, "
          f"full trajectory {S * L / (t3 - t2) / 1e6:8.1f} M ticks/s")
n = 64
t0 = time.perf_counter(); OT.replay(status[:n], score[:n], 1 / 30); t1 = time.perf_counter()
print(f"CPU oracle port (pure Python, 1 core): {n * L / (t1 - t0) / 1e6:8.3f} M ticks/s")
