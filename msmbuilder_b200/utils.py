"""Input validation with the reference's semantics.

check_iter_of_sequences  ~ msmbuilder/utils/validation.py:25-55
array2d                  ~ msmbuilder/utils/validation.py:58-74

Differences, both additive: torch tensors (CPU or CUDA) are accepted wherever an
ndarray is, and because mdtraj may be absent a "trajectory" is anything exposing
``.xyz`` of shape (n_frames, n_atoms, 3) -- or a bare float32 array of that shape.
"""
import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

__all__ = ["check_iter_of_sequences", "array2d", "is_trajectory", "is_tensor"]


def is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def is_trajectory(x):
    return hasattr(x, "xyz") and hasattr(x, "n_atoms") and not isinstance(x, np.ndarray)


def check_iter_of_sequences(sequences, allow_trajectory=False, ndim=2, max_iter=None):
    """Raise ValueError('sequences must be a list of sequences') unless every
    inspected item is an ``ndim``-dimensional array (or an allowed trajectory)."""
    if hasattr(sequences, "shapes") and hasattr(sequences, "to_device"):
        # io.NumpyDirStream: look at the file headers instead of uploading anything
        for shape, _ in sequences.shapes():
            if len(shape) != ndim and not (allow_trajectory and len(shape) == 3):
                raise ValueError('sequences must be a list of sequences')
        return
    ok = True
    for i, X in enumerate(sequences):
        if is_trajectory(X):
            if not allow_trajectory:
                ok = False
                break
        else:
            nd = getattr(X, "ndim", None)
            if nd is None or (nd != ndim and not (allow_trajectory and nd == 3)):
                ok = False
                break
        if max_iter is not None and i >= max_iter:
            break
    if not ok:
        raise ValueError('sequences must be a list of sequences')


def array2d(X, dtype=None, order=None, copy=False, force_all_finite=True):
    """At-least-2-D ndarray view of X; ValueError on NaN / inf (host arrays only --
    device tensors are checked from the column sums the kernel returns)."""
    X_2d = np.asarray(np.atleast_2d(X), dtype=dtype, order=order)
    if force_all_finite and X_2d.dtype.kind == 'f':
        if not np.isfinite(X_2d.sum()) and not np.isfinite(X_2d).all():
            raise ValueError("Input contains NaN, infinity or a value too large "
                             "for %r." % X_2d.dtype)
    if X is X_2d and copy:
        X_2d = np.copy(X_2d, order='K')
    return X_2d
