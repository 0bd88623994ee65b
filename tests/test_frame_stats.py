"""f1: the reference's real per-frame pixel path.  CPU part: the integer oracle and the
product's host finishing (fav.gate.SignalFinisher) reproduce the reference SignalAnalyzer's outputs
exactly (goldens generated from /root/reference by tests/golden/make_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import frame_stats as FS

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_frames import frame_sequence  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def test_oracle_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    f = np.random.default_rng(5).integers(0, 256, (97, 131, 3), dtype=np.uint8)
    gray = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(FS.gray_bgr(f), gray)
    lap = cv2.Laplacian(gray, cv2.CV_64F)
    assert np.array_equal(FS.laplacian(gray).astype(np.float64), lap)
    st, _ = FS.frame_stats(f)
    assert np.array_equal(st[4:], cv2.calcHist([gray], [0], None, [256], [0, 256]).ravel().astype(np.int64))


def test_finisher_reproduces_reference_signal_analyzer():
    import fav.gate as G
    gold = load("signal_analyzer.json")
    for case in gold["cases"]:
        fin = G.SignalFinisher()
        prev = None
        for f, want in zip(frame_sequence(case["seed"], case["h"], case["w"]), case["results"]):
            st, gray = FS.frame_stats(f, prev)
            prev = gray
            r = fin.finish(st, case["h"] * case["w"])
            assert r["vision_status"] == want["vision_status"]
            assert round(r["signal_score"], 6) == want["anomaly_score"]
            m = want["metrics"]
            assert round(r["blur_score"], 4) == m["blur"] and round(r["brightness_score"], 4) == m["brightness"]
            assert round(r["freeze_score"], 4) == m["freeze"] and round(r["entropy_score"], 4) == m["entropy"]
            assert round(r["laplacian_var"], 2) == m["raw"]["laplacian_var"]
            assert round(r["mean_brightness"], 1) == m["raw"]["mean_brightness"]
            assert round(r["mean_diff"], 2) == m["raw"]["frame_diff"]
            assert round(r["entropy"], 3) == m["raw"]["entropy"]


def test_golden_covers_every_status():
    gold = load("signal_analyzer.json")
    seen = {r["vision_status"] for c in gold["cases"] for r in c["results"]}
    assert seen == {"VISION_OK", "VISION_BLANK", "VISION_FROZEN", "VISION_CORRUPTED"}
