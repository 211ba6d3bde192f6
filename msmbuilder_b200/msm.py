"""Transition counting on device-resident state labels.

``transition_counts`` is ``msmbuilder.msm.core._transition_counts``
(msm/core.py:487-602): the (label_t, label_{t+lag}) histogram every MSM is built
from, the step that directly consumes the labels the assignment kernels produce.
The labels never leave the GPU: a min/max scan and a presence bitmap give the
reference's sorted ``classes`` (core.py:544) without a sort, then one histogram
kernel (warp-merged atomics, shared-memory bins when n_states**2 <= 8192) counts
every pair inside each sequence.  Counts are exact integers, so the result is
bit-identical to the reference's float matrix.

Labels must be integers (NumPy / torch integer arrays, or float arrays whose
finite values are whole numbers, NaN = missing, as produced by
``MSM.partial_transform(mode='fill')``).  Arbitrary Python objects (strings,
None) are a host-side relabelling problem and are not taken here.
"""
from __future__ import absolute_import, print_function, division

import numpy as np

from . import _lib
from .utils import is_tensor

__all__ = ['transition_counts']

_MISSING = np.iinfo(np.int64).min
_MAX_SPAN = 1 << 28


def _as_label_tensor(y):
    """1-D sequence -> (CUDA int32/int64 tensor, is_float) ; NaN -> INT64_MIN."""
    import torch
    if is_tensor(y):
        t = y
    else:
        a = np.asarray(y)
        if a.ndim != 1:
            raise ValueError('sequences must be a list of 1-D label sequences')
        if a.dtype.kind not in 'iufb':
            raise TypeError('device transition counting takes integer (or whole-number '
                            'float, NaN = missing) labels, got dtype %s' % a.dtype)
        if a.dtype.kind == 'b':
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.ndim != 1:
        raise ValueError('sequences must be a list of 1-D label sequences')
    t = t.cuda()
    if t.dtype in (torch.int32, torch.int64):
        return t.contiguous(), False
    if t.dtype.is_floating_point:
        nan = torch.isnan(t)
        whole = torch.where(nan, torch.zeros_like(t), t)
        if not bool((whole == torch.floor(whole)).all().item()):
            raise TypeError('float labels must be whole numbers (NaN = missing)')
        out = whole.to(torch.int64)
        out[nan] = _MISSING
        return out, True
    return t.to(torch.int64).contiguous(), False


def transition_counts(sequences, lag_time=1, sliding_window=True):
    """Count the directed transitions in a collection of label sequences.

    Parameters
    ----------
    sequences : list of 1-D integer arrays (NumPy, or torch on host / GPU)
    lag_time : int
        Index delay of a transition.
    sliding_window : bool
        True: every frame starts a transition and the counts are divided by
        ``lag_time``; False: only frames 0, lag, 2 lag, ... of each sequence do.

    Returns
    -------
    counts : ndarray, shape (n_states, n_states), float64
        ``counts[i, j]``: transitions from state i to state j.
    mapping : dict
        label -> row/column index, labels in ascending order.
    """
    import torch
    from . import _device as dev
    _lib.require_gpu()
    lag_time = int(lag_time)
    if lag_time < 1:
        raise ValueError('lag_time must be a positive integer')
    seqs = list(sequences)
    for s in seqs:
        if not (is_tensor(s) or hasattr(s, '__len__')):
            # np.concatenate of scalars in core.py:544
            raise ValueError('sequences must be a list of sequences')
    strided = (not sliding_window) and lag_time > 1
    parts, any_float = [], False
    for s in seqs:
        t, was_float = _as_label_tensor(s)
        if strided:                      # core.py:540-542: count X[::lag] at lag 1
            t = t[::lag_time].contiguous()
        any_float = any_float or was_float
        parts.append(t)
    lengths = [int(p.numel()) for p in parts]
    n_total = int(sum(lengths))
    if any(p.dtype == torch.int64 for p in parts):
        parts = [p.to(torch.int64) for p in parts]
    labels = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device="cuda")
    label_bytes = labels.element_size()
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)).cuda()

    # classes = np.unique(labels) without NaN (core.py:544-553)
    rng = torch.zeros(2, dtype=torch.int64, device="cuda")
    _lib.call("msmb200_label_range", dev.ptr(labels), n_total, label_bytes, dev.ptr(rng),
              dev.stream_ptr())
    lo, hi = (int(v) for v in rng.cpu().numpy())
    if lo > hi:
        return np.zeros((0, 0)), {}
    span = hi - lo + 1
    if span > _MAX_SPAN:
        raise ValueError('labels span %d values; relabel them to a compact range first' % span)
    flags = torch.empty(span, dtype=torch.uint8, device="cuda")
    _lib.call("msmb200_label_presence", dev.ptr(labels), n_total, label_bytes, lo, span,
              dev.ptr(flags), dev.stream_ptr())
    present = flags.cpu().numpy().astype(bool)
    classes = np.flatnonzero(present).astype(np.int64) + lo
    n_states = len(classes)
    if n_states > 46340:
        raise ValueError('%d states need a %d^2 count matrix; too large' % (n_states, n_states))
    keys = classes.astype(np.float64) if any_float else classes
    mapping = dict(zip(keys.tolist() if any_float else [int(c) for c in classes], range(n_states)))

    identity = (lo == 0 and n_states == span)
    remap = None
    if not identity:
        table = np.full(span, -1, dtype=np.int32)
        table[classes - lo] = np.arange(n_states, dtype=np.int32)
        remap = torch.from_numpy(table).cuda()

    counts = torch.zeros((n_states, n_states), dtype=torch.int64, device="cuda")
    lag = 1 if strided else lag_time
    _lib.call("msmb200_transition_counts", dev.ptr(labels), label_bytes, dev.ptr(offsets),
              len(lengths), n_total, lag, dev.ptr(remap), lo,
              span if remap is not None else 0, n_states, dev.ptr(counts), dev.stream_ptr())
    out = counts.cpu().numpy().astype(np.float64)
    out /= float(lag)                   # core.py:600 (the strided recursion divides by 1)
    return out, mapping
