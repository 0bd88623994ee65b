"""Uncertainty scores + failure flag -- CPU oracle (numpy fp32).  TEST INFRASTRUCTURE ONLY.

No reference counterpart (nearest: gray-level histogram entropy, platform/backend/
signal_analyzer.py:100-112 -- a different quantity).  Failure definition from the reference's
README.md:22-24: "Incorrect prediction with high confidence"; the threshold tau is config.
Formulas: SURVEY.md Appendix A.5.
"""
import numpy as np


def softmax(z):
    z = z.astype(np.float32)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m).astype(np.float32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def entropy(p):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, p * np.log(p), np.float32(0)).astype(np.float32)
    return (-t.sum(axis=-1, dtype=np.float32)).astype(np.float32)


def uncertainty(logits, labels=None, tau=0.9):
    """logits float32 [N,T,C] -> dict of per-sample arrays (conf, entropy, mutual_information,
    pred, failure_flag).  pred = argmax of the pass-mean probabilities (lowest index on ties)."""
    p_t = softmax(logits)                                            # [N,T,C]
    T = logits.shape[1]
    pbar = (p_t.sum(axis=1, dtype=np.float32) * np.float32(1.0 / T)).astype(np.float32)
    pred = pbar.argmax(axis=-1).astype(np.int32)
    conf = pbar.max(axis=-1).astype(np.float32)
    H = entropy(pbar)
    Hm = (entropy(p_t).sum(axis=1, dtype=np.float32) * np.float32(1.0 / T)).astype(np.float32)
    mi = np.maximum(H - Hm, np.float32(0)).astype(np.float32)
    out = dict(confidence=conf, entropy=H, mutual_information=mi, pred=pred, pbar=pbar)
    if labels is not None:
        out["failure_flag"] = ((pred != labels) & (conf >= np.float32(tau))).astype(np.uint8)
    return out


def top2_gap(pbar):
    s = np.sort(pbar, axis=-1)
    return (s[:, -1] - s[:, -2]).astype(np.float32)
