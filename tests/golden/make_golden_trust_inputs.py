"""Seeded tick sequences for the trust-replay goldens (pure numpy: shared by the generator, which needs the reference,
and by the tests, which must not)."""
import numpy as np


def sequences(seed, n_seq, n_ticks):
    """Piecewise-constant status runs (like the playground's events) with noisy anomaly scores, some None, some outliers."""
    rng = np.random.default_rng(seed)
    status = np.zeros((n_seq, n_ticks), np.int8)
    score = np.zeros((n_seq, n_ticks), np.float64)
    for s in range(n_seq):
        i = 0
        while i < n_ticks:
            run = int(rng.integers(1, 120))
            st = int(rng.choice(4, p=[0.55, 0.15, 0.15, 0.15]))
            status[s, i:i + run] = st
            i += run
        base = rng.uniform(0.0, 0.3)
        score[s] = np.clip(base + 0.02 * rng.standard_normal(n_ticks), 0.0, 1.0)
        spikes = rng.random(n_ticks) < 0.02
        score[s, spikes] = rng.uniform(0.5, 1.0, int(spikes.sum()))
        score[s, rng.random(n_ticks) < 0.05] = np.nan
    return status, score


