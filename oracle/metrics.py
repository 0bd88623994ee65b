"""Calibration / failure-detection aggregates -- CPU oracle.  TEST INFRASTRUCTURE ONLY.

No reference counterpart (failure_attributor.py:93-108 summarises trust excursions, a
different thing).  Definitions: SURVEY.md Appendix A.6 -- 15-bin right-closed ECE (Guo et al.),
bucketed AUROC for failure detection (positive = misclassified), confusion counts.  All
accumulators are integers (counts and Q32 fixed-point sums) so that sums are order-independent
and bit-identical on 1/2/4/8 GPUs.

Arena layout (int64 words; mirrored by include/fav_b200.h FAV_HIST_*):
  [0]                 n
  [1]                 n_correct
  [2]                 n_flag                (wrong with confidence >= tau)
  [3]                 sum_conf_q32
  [4]                 sum_entropy_q32       (H / ln C, clipped to [0,1])
  [5]                 sum_mi_q32            (MI / ln C, clipped to [0,1])
  [6]                 n_invalid             (label outside [0, C): counted here, excluded from everything else)
  [7]                 reserved
  [8 + 3*b + {0,1,2}] ECE bin b: count, sum_conf_q32, n_correct          (B bins)
  [8+3B + (s*K + k)*2 + {0,1}]  AUROC score type s (0: 1-conf, 1: H/lnC, 2: MI/lnC),
                                bucket k: {neg = correct, pos = wrong}   (K buckets)
  [8+3B+6K + ...]     confusion: C*C (label-major) if C <= 100 else per class (total, correct)
"""
import numpy as np

N_BINS = 15
N_BUCKETS = 4096
HDR = 8


def arena_words(num_classes, n_bins=N_BINS, n_buckets=N_BUCKETS):
    conf = num_classes * num_classes if num_classes <= 100 else 2 * num_classes
    return HDR + 3 * n_bins + 6 * n_buckets + conf


def q32(x):
    """round(x * 2^32) for fp32 x in [0,1] -- exact (power-of-two scaling)."""
    return np.rint(x.astype(np.float32).astype(np.float64) * 2.0 ** 32).astype(np.int64)


def ece_bin(conf, n_bins=N_BINS):
    """right-closed bins (b/B, (b+1)/B]; conf == 0 goes to bin 0.  fp32 arithmetic as the kernel."""
    b = np.ceil(conf.astype(np.float32) * np.float32(n_bins)).astype(np.int32) - 1
    return np.clip(b, 0, n_bins - 1)


def bucket(score, n_buckets=N_BUCKETS):
    b = np.floor(score.astype(np.float32) * np.float32(n_buckets)).astype(np.int32)
    return np.clip(b, 0, n_buckets - 1)


def normalised_scores(conf, H, mi, num_classes):
    inv = np.float32(1.0 / np.log(float(num_classes)))
    s0 = np.clip(np.float32(1.0) - conf, 0, 1).astype(np.float32)
    s1 = np.clip(H * inv, 0, 1).astype(np.float32)
    s2 = np.clip(mi * inv, 0, 1).astype(np.float32)
    return s0, s1, s2


def accumulate(arena, conf, H, mi, pred, labels, tau, num_classes, n_bins=N_BINS, n_buckets=N_BUCKETS):
    ok = (labels >= 0) & (labels < num_classes) & (pred >= 0) & (pred < num_classes)
    arena[6] += int((~ok).sum())
    conf, H, mi, pred, labels = conf[ok], H[ok], mi[ok], pred[ok], labels[ok]
    correct = (pred == labels)
    s0, s1, s2 = normalised_scores(conf, H, mi, num_classes)
    arena[0] += len(conf)
    arena[1] += int(correct.sum())
    arena[2] += int((~correct & (conf >= np.float32(tau))).sum())
    arena[3] += int(q32(conf).sum())
    arena[4] += int(q32(s1).sum())
    arena[5] += int(q32(s2).sum())
    b = ece_bin(conf, n_bins)
    np.add.at(arena, HDR + 3 * b, 1)
    np.add.at(arena, HDR + 3 * b + 1, q32(conf))
    np.add.at(arena, HDR + 3 * b + 2, correct.astype(np.int64))
    base = HDR + 3 * n_bins
    for s, sc in enumerate((s0, s1, s2)):
        k = bucket(sc, n_buckets)
        np.add.at(arena, base + (s * n_buckets + k) * 2 + (~correct).astype(np.int64), 1)
    cb = base + 6 * n_buckets
    if num_classes <= 100:
        np.add.at(arena, cb + labels.astype(np.int64) * num_classes + pred, 1)
    else:
        np.add.at(arena, cb + 2 * labels.astype(np.int64), 1)
        np.add.at(arena, cb + 2 * labels.astype(np.int64) + 1, correct.astype(np.int64))
    return arena


def auroc_from_buckets(neg, pos):
    """AUROC = P(score_pos > score_neg) + 0.5 P(tie) from per-bucket counts (float64)."""
    neg = neg.astype(np.float64)
    pos = pos.astype(np.float64)
    P, Nn = pos.sum(), neg.sum()
    if P == 0 or Nn == 0:
        return float("nan")
    below = np.concatenate([[0.0], np.cumsum(neg)[:-1]])
    return float((pos * (below + 0.5 * neg)).sum() / (P * Nn))


def finalize(arena, num_classes, n_bins=N_BINS, n_buckets=N_BUCKETS):
    a = np.asarray(arena, dtype=np.int64)
    n = int(a[0])
    out = dict(n=n, n_invalid=int(a[6]))
    if n == 0:
        return out
    out["accuracy"] = a[1] / n
    out["failure_rate"] = a[2] / n
    out["mean_confidence"] = a[3] / 2.0 ** 32 / n
    out["mean_entropy"] = a[4] / 2.0 ** 32 / n
    out["mean_mutual_information"] = a[5] / 2.0 ** 32 / n
    bins = a[HDR:HDR + 3 * n_bins].reshape(n_bins, 3)
    ece = 0.0
    for cnt, sc, nc in bins:
        if cnt:
            ece += cnt / n * abs(nc / cnt - sc / 2.0 ** 32 / cnt)
    out["ece"] = float(ece)
    base = HDR + 3 * n_bins
    for s, nm in enumerate(("auroc_msp", "auroc_entropy", "auroc_mi")):
        bk = a[base + s * 2 * n_buckets: base + (s + 1) * 2 * n_buckets].reshape(n_buckets, 2)
        out[nm] = auroc_from_buckets(bk[:, 0], bk[:, 1])
    return out
