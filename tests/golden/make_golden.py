"""Generates tests/golden/*.json by running the REAL reference code from /root/reference
(read-only; only importable in the build container, never on the GPU box).

  python tests/golden/make_golden.py

signal_analyzer.json : SignalAnalyzer.analyze_frame (platform/backend/signal_analyzer.py:47-143) on
                       seeded frame sequences (frames are re-generated from the seed by the tests).
trust_engine.json    : the test_trust.py transcript (platform/backend/test_trust.py:1-33) as numbers,
                       plus TrustEngine driven by the analyzer outputs above.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/platform/backend"
sys.path.insert(0, REF)
from signal_analyzer import SignalAnalyzer  # noqa: E402
from trust_engine import TrustEngine  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


from make_golden_frames import frame_sequence  # noqa: E402


def main():
    out = {"cases": []}
    for seed, (h, w) in ((0, (240, 320)), (1, (480, 640)), (2, (96, 130))):
        an = SignalAnalyzer()
        eng = TrustEngine()
        res, states = [], []
        for f in frame_sequence(seed, h, w):
            r = an.analyze_frame(f)
            res.append(r)
            s = eng.update(r["vision_status"], r["anomaly_score"], 1 / 30)
            states.append({"reliability": s["reliability"], "policy_state": s["policy_state"]})
        out["cases"].append({"seed": seed, "h": h, "w": w, "results": res, "trust": states})
    with open(os.path.join(HERE, "signal_analyzer.json"), "w") as fh:
        json.dump(out, fh, indent=1)

    e = TrustEngine()
    dt = 0.033
    tr = []
    s = e.update('VISION_OK', 0.019, dt); tr.append(["OK", s["reliability"], s["policy_state"]])
    for _ in range(50):
        s = e.update('VISION_FROZEN', 0.019, dt)
    tr.append(["FROZEN x50", s["reliability"], s["policy_state"]])
    for _ in range(30):
        s = e.update('VISION_BLANK', None, dt)
    tr.append(["BLANK x30", s["reliability"], s["policy_state"]])
    for _ in range(100):
        s = e.update('VISION_CORRUPTED', None, dt)
    tr.append(["CORRUPT x100", s["reliability"], s["policy_state"]])
    for _ in range(200):
        s = e.update('VISION_OK', 0.019, dt)
    tr.append(["RECOVER x200", s["reliability"], s["policy_state"]])
    with open(os.path.join(HERE, "trust_engine.json"), "w") as fh:
        json.dump({"transcript": tr}, fh, indent=1)
    print("wrote goldens:", [c["results"][0]["anomaly_score"] for c in out["cases"]], tr)


if __name__ == "__main__":
    main()
