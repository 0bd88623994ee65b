"""Generates tests/golden/trust_replay.json by driving the REAL reference TrustEngine (/root/reference, read-only, only
importable in the build container) over seeded random tick sequences, the way platform/backend/main.py:340-352 replays
them.  Raw engine attributes are stored with repr() precision so the comparison is bit-exact.

  python tests/golden/make_golden_trust.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/platform/backend")
from trust_engine import TrustEngine  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_trust_inputs import sequences  # noqa: E402
STATUS = ("VISION_OK", "VISION_FROZEN", "VISION_BLANK", "VISION_CORRUPTED")
POLICY = ("VISION_ALLOWED", "VISION_DECLINING", "VISION_DEGRADED", "VISION_BLOCKED")


def main():
    cases = []
    for seed, n_seq, n_ticks, dt in ((0, 6, 400, 1.0 / 30.0), (1, 4, 900, 0.033), (2, 3, 200, 0.25)):
        status, score = sequences(seed, n_seq, n_ticks)
        traj = []
        for s in range(n_seq):
            e = TrustEngine()
            rows = []
            for i in range(n_ticks):
                sc = None if np.isnan(score[s, i]) else float(score[s, i])
                st = e.update(STATUS[status[s, i]], sc, dt)
                rows.append([e.reliability, e.anomaly_integral, e.trust_velocity, e.recovery_debt, e.recovery_coeff,
                             POLICY.index(e.policy_state), int(e.contradiction_detected), e.contradiction_count,
                             st["reliability"], st["trust_velocity"]])
            traj.append(rows)
        cases.append({"seed": seed, "n_seq": n_seq, "n_ticks": n_ticks, "dt": dt, "trajectories": traj})
    with open(os.path.join(HERE, "trust_replay.json"), "w") as fh:
        json.dump({"cases": cases}, fh)
    print("wrote trust_replay.json:", sum(c["n_seq"] * c["n_ticks"] for c in cases), "ticks")


if __name__ == "__main__":
    main()
