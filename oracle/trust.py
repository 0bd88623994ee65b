"""Batched replay of the reference's TrustEngine -- CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates ``TrustEngine.update`` (platform/backend/trust_engine.py:139-243), ``_update_policy`` (:68-87) and
``_update_contradiction_detector`` (:89-137) for S independent sequences of L ticks, as driven by the reference's batch
replay loop (platform/backend/main.py:340-352).  PARITY PINNED: tests/golden/trust_replay.json holds state trajectories
produced by the REAL reference class on seeded random sequences (tests/golden/make_golden_trust.py); this restatement
reproduces reliability / anomaly_integral / trust_velocity / recovery_debt / recovery_coeff bit for bit (same float64
operations in the same order) and the policy / contradiction outputs exactly on those vectors.  The contradiction
detector's mean / stdev go through Python's exact-rational ``statistics`` module in the reference; here they are plain
float64 two-pass sums, so a z-score within ~1e-12 of the 3.0 threshold could in principle flip (none does on the goldens).
"""
import math

import numpy as np

STATUS = ("VISION_OK", "VISION_FROZEN", "VISION_BLANK", "VISION_CORRUPTED")          # status code = index
POLICY = ("VISION_ALLOWED", "VISION_DECLINING", "VISION_DEGRADED", "VISION_BLOCKED")  # policy code = index
DECAY = (0.0, 0.30, 0.60, 1.00)            # reliability decay per second for codes 1..3 (trust_engine.py:206-228)
RECOVERY_DEBT_MAX, RECOVERY_DEBT_GAIN, RECOVERY_MIN_COEFF, RECOVERY_DEBT_DRAIN = 10.0, 0.008, 0.03, 0.10
ANOMALY_DECAY_GAIN, ANOMALY_LEAK, EMA_ALPHA, BUF = 0.15, 0.5, 0.12, 60
FIELDS = ("reliability", "anomaly_integral", "trust_velocity", "recovery_debt", "recovery_coeff")


def _policy(rel, vel):
    if rel >= 0.7 and vel < -0.15:
        return 1
    if rel >= 0.7:
        return 0
    if rel >= 0.3:
        return 2
    return 3


def replay(status, score, dt):
    """status int [S,L] (codes above); score float64 [S,L] (NaN = None); dt float or float64 [L].
    Returns dict: 'state' float64 [S,L,5] (FIELDS, unrounded), 'policy' uint8 [S,L], 'contradiction' uint8 [S,L],
    'contradiction_count' int32 [S,L]."""
    status = np.asarray(status)
    score = np.asarray(score, dtype=np.float64)
    S, L = status.shape
    dts = np.full(L, dt, dtype=np.float64) if np.isscalar(dt) else np.asarray(dt, dtype=np.float64)
    out = np.zeros((S, L, 5), np.float64)
    pol = np.zeros((S, L), np.uint8)
    con = np.zeros((S, L), np.uint8)
    cnt = np.zeros((S, L), np.int32)
    for s in range(S):
        rel, integ, vel, debt, coeff, prev_rel = 1.0, 0.0, 0.0, 0.0, 0.10, 1.0
        cur, policy, contra, count = -1, 0, 0, 0
        buf = []
        for i in range(L):
            st, d = int(status[s, i]), float(dts[i])
            sc = None if math.isnan(score[s, i]) else float(score[s, i])
            if cur < 0:                                   # first call (:153-158)
                cur = st
                policy = _policy(rel, vel)
            elif st != cur:                               # status change (:161-170)
                prev, cur = cur, st
                if st != 0 and prev == 0:
                    integ = 0.0
                policy = _policy(rel, vel)
            else:
                if st == 0:                               # (:178-197)
                    debt = max(0.0, debt - RECOVERY_DEBT_DRAIN * d)
                    coeff = max(RECOVERY_MIN_COEFF, 0.10 - RECOVERY_DEBT_GAIN * debt)
                    rel += coeff * d
                    if sc is not None:
                        integ += sc * d
                        integ -= ANOMALY_LEAK * integ * d
                        integ = max(0.0, integ)
                        rel -= (ANOMALY_DECAY_GAIN * integ) * d
                else:                                     # (:199-228)
                    debt = min(RECOVERY_DEBT_MAX, debt + max(0.0, 0.7 - rel) * d)
                    rel -= DECAY[st] * d
                    integ = 0.0
                rel = max(0.0, min(1.0, rel))
                vel = EMA_ALPHA * ((rel - prev_rel) / max(d, 0.001)) + (1 - EMA_ALPHA) * vel
                prev_rel = rel
                # contradiction detector (:89-137)
                if sc is None:
                    contra = 0
                else:
                    buf.append((st, sc))
                    if len(buf) > BUF:
                        buf.pop(0)
                    same = [x for t, x in buf if t == st]
                    if len(buf) < 30 or len(same) < 10:
                        contra = 0
                    else:
                        mean = math.fsum(same) / len(same)
                        var = math.fsum((x - mean) ** 2 for x in same) / (len(same) - 1)
                        std = max(math.sqrt(var), 0.001)
                        if st == 0 and (sc - mean) / std > 3.0:
                            if not contra:
                                count += 1
                            contra = 1
                        else:
                            contra = 0
                policy = _policy(rel, vel)
            out[s, i] = (rel, integ, vel, debt, coeff)
            pol[s, i], con[s, i], cnt[s, i] = policy, contra, count
    return {"state": out, "policy": pol, "contradiction": con, "contradiction_count": cnt}


def reference_state(res, s, i):
    """The subset of TrustEngine.get_state() (trust_engine.py:245-263) that the replay determines, with its rounding."""
    rel, integ, vel, debt, coeff = (float(v) for v in res["state"][s, i])
    return {"reliability": round(rel, 6), "policy_state": POLICY[int(res["policy"][s, i])],
            "anomaly_integral": round(integ, 6), "trust_velocity": round(vel, 6), "recovery_debt": round(debt, 4),
            "recovery_coeff": round(coeff, 4), "contradiction_detected": bool(res["contradiction"][s, i]),
            "contradiction_count": int(res["contradiction_count"][s, i])}
