"""Functional distance API, GPU-backed.

Same five functions, argument order, return types and error behaviour as the
reference's Cython module ``msmbuilder.libdistance``
(msmbuilder/libdistance/libdistance.pyx:82,134,182,229,273):

    assign_nearest(X, Y, metric, X_indices=None) -> (intp[n], float)
    cdist(XA, XB, metric)                         -> float64[na, nb]
    pdist(X, metric, X_indices=None)              -> float64[m(m-1)/2]
    dist(X, y, metric, X_indices=None)            -> float64[n]
    sumdist(X, metric, pair_indices)              -> float

NumPy arrays in, NumPy arrays out (each call uploads its operands; the
estimators in msmbuilder_b200.cluster keep data resident instead and use
msmbuilder_b200._kernels directly).  metric='rmsd' takes trajectories -- objects
with ``.xyz`` of shape (n, n_atoms, 3) -- or bare float32 arrays of that shape.
"""
import numpy as np

from . import _lib
from . import _kernels as K
from . import _device as dev
from .utils import is_trajectory, is_tensor

VECTOR_METRICS = _lib.VECTOR_METRICS
__all__ = ['assign_nearest', 'cdist', 'dist', 'pdist', 'sumdist']


def _metric(metric):
    return metric.decode() if isinstance(metric, bytes) else metric


def _xyz(T):
    """(n, n_atoms, 3) float32 coordinates of a trajectory-like object."""
    a = T.xyz if is_trajectory(T) else T
    if is_tensor(a):
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("rmsd needs coordinates of shape (n_frames, n_atoms, 3)")
        return a
    a = np.asarray(a)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("rmsd needs coordinates of shape (n_frames, n_atoms, 3)")
    return np.ascontiguousarray(a, dtype=np.float32)


def _is_rmsd_input(X):
    return is_trajectory(X) or (getattr(X, "ndim", 0) == 3)


def _centered(T):
    """Device copy of the coordinates, centred, with traces (libdistance.pyx:336-341;
    the reference centres the caller's Trajectory in place, we centre our copy)."""
    import torch
    a = _xyz(T)
    t = (a if is_tensor(a) else torch.from_numpy(a)).to(device="cuda", dtype=torch.float32)
    t = t.contiguous().clone()
    tr = K.rmsd_center(t)
    return t, tr


def _check_vector(metric, *arrays):
    if metric not in VECTOR_METRICS:
        raise ValueError('metric must be one of %s' % ', '.join("'%s'" % s for s in VECTOR_METRICS))
    for a in arrays:
        if not (isinstance(a, np.ndarray) or is_tensor(a)):
            raise TypeError()
    kinds = {str(a.dtype).replace("torch.", "") for a in arrays}
    if kinds == {"float64"} or kinds == {"float32"}:
        return
    raise TypeError('X and y must be both float32 or float64')


def _same_atoms(a, b):
    if a.shape[1] != b.shape[1]:
        raise ValueError("Input trajectories must have same number of atoms. "
                         "found %d and %d." % (a.shape[1], b.shape[1]))


def assign_nearest(X, Y, metric, X_indices=None):
    """For each point in X (or X[X_indices]) the index of the nearest point of Y
    (lowest index on ties) and the sum of those distances."""
    metric = _metric(metric)
    if metric == "rmsd" and _is_rmsd_input(X) and _is_rmsd_input(Y):
        x, xt = _centered(X)
        y, yt = _centered(Y)
        _same_atoms(x, y)
        labels, _, inertia = K.rmsd_assign_nearest(x, xt, y, yt, rows=X_indices)
        return labels.cpu().numpy().astype(np.intp), float(inertia)
    _check_vector(metric, X, Y)
    labels, _, inertia = K.assign_nearest(dev.to_device(X), dev.to_device(Y), metric,
                                          rows=X_indices)
    return labels.cpu().numpy().astype(np.intp), float(inertia)


def cdist(XA, XB, metric):
    """All distances between the rows of XA and the rows of XB."""
    metric = _metric(metric)
    if metric == "rmsd" and _is_rmsd_input(XA) and _is_rmsd_input(XB):
        import torch
        a, at = _centered(XA)
        b, bt = _centered(XB)
        _same_atoms(a, b)
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device="cuda")
        for j in range(int(b.shape[0])):
            out[:, j] = K.rmsd_dist(a, at, b[j], float(bt[j]))
        return out.cpu().numpy()
    _check_vector(metric, XA, XB)
    return K.cdist(dev.to_device(XA), dev.to_device(XB), metric).cpu().numpy()


def pdist(X, metric, X_indices=None):
    """Condensed pairwise distances of X (or of the gathered rows X[X_indices])."""
    metric = _metric(metric)
    if metric == "rmsd" and _is_rmsd_input(X):
        x, xt = _centered(X)
        return K.rmsd_pdist(x, xt, rows=X_indices).cpu().numpy()
    _check_vector(metric, X)
    return K.pdist(dev.to_device(X), metric, rows=X_indices).cpu().numpy()


def dist(X, y, metric, X_indices=None):
    """Distance from every row of X (or X[X_indices]) to the single point y."""
    metric = _metric(metric)
    if metric == "rmsd" and _is_rmsd_input(X):
        x, xt = _centered(X)
        yy, yt = _centered(y if _is_rmsd_input(y) else np.asarray(y)[None])
        _same_atoms(x, yy)
        return K.rmsd_dist(x, xt, yy[0], float(yt[0]), rows=X_indices).cpu().numpy()
    _check_vector(metric, X, y)
    yv = dev.to_device(y, ndim=1).reshape(-1)
    return K.dist(dev.to_device(X), yv, metric, rows=X_indices).cpu().numpy()


def sumdist(X, metric, pair_indices):
    """sum(dist(X[i], X[j]) for (i, j) in pair_indices)."""
    metric = _metric(metric)
    if metric == "rmsd" and _is_rmsd_input(X):
        x, xt = _centered(X)
        pairs = np.asarray(pair_indices, dtype=np.int64).reshape(-1, 2)
        total = 0.0
        for i, j in pairs:
            total += float(K.rmsd_dist(x, xt, x[j], float(xt[j]), rows=[int(i)])[0])
        return total
    if metric not in VECTOR_METRICS:
        raise ValueError('metric must be one of %s' % ', '.join("'%s'" % s for s in VECTOR_METRICS))
    if str(X.dtype).replace("torch.", "") not in ("float32", "float64"):
        raise TypeError('X must be both float32 or float64')
    return float(K.sumdist(dev.to_device(X), metric, pair_indices))
