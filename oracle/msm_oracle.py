"""oracle/msm_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

NumPy restatement of `_transition_counts` (msmbuilder/msm/core.py:487-602) for
numeric labels: classes = sorted unique non-NaN labels (:544-553), one pair
(y[t], y[t+lag]) per t inside each sequence (:567-587), pairs touching NaN dropped
(:576-580), counts divided by lag_time for the sliding window (:600); for
sliding_window=False the sequences are strided by lag_time and counted at lag 1
(:540-542).

Pinned in tests/test_oracle_msm.py against the reference function itself loaded
verbatim (oracle/ref_loader.load_transition_counts), against the known answers of
the reference's tests/test_transition_counts.py and against
tests/golden/msm_counts.npz.
"""
import numpy as np


def transition_counts(sequences, lag_time=1, sliding_window=True):
    if (not sliding_window) and lag_time > 1:
        return transition_counts([np.asarray(X)[::lag_time] for X in sequences], lag_time=1)
    seqs = [np.asarray(y) for y in sequences]
    for y in seqs:
        if y.ndim != 1:
            raise ValueError("sequences must be a list of 1-D sequences")
    flat = np.concatenate(seqs) if seqs else np.zeros(0, dtype=np.int64)
    classes = np.unique(flat)
    if classes.dtype.kind == "f":
        classes = classes[~np.isnan(classes)]
    n_states = len(classes)
    mapping = dict(zip(classes.tolist(), range(n_states)))
    counts = np.zeros((n_states, n_states), dtype=np.float64)
    for y in seqs:
        a, b = y[:-lag_time], y[lag_time:]
        if y.dtype.kind == "f":
            ok = ~(np.isnan(a) | np.isnan(b))
            a, b = a[ok], b[ok]
        if len(a) == 0:
            continue
        ia = np.searchsorted(classes, a)
        ib = np.searchsorted(classes, b)
        np.add.at(counts, (ia, ib), 1.0)
    counts /= float(lag_time)
    return counts, mapping
