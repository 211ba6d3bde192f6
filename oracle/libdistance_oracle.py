"""oracle/libdistance_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes face of the two CPU checkers for the libdistance part of the hot path:

* ``impl="port"``       -> oracle/_build/liboracle.so (libdistance_oracle.c, our
                            C restatement),
* ``impl="reference"``  -> oracle/_ref/libref.so (the unmodified reference
                            headers/kmedoids.cc compiled through ref_shim.cc).

The function names and argument meaning mirror the thin Cython dispatch of
msmbuilder/libdistance/libdistance.pyx:82-131 (assign_nearest), :134-179
(cdist), :182-226 (pdist), :229-270 (dist), :273-310 (sumdist) and
msmbuilder/cluster/_kmedoids.pyx:23-117 (kmedoids, contigify_ids), including the
error behaviour (ValueError for an unknown metric, TypeError for mixed dtypes).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

VECTOR_METRICS = ("euclidean", "sqeuclidean", "cityblock", "chebyshev",
                  "canberra", "braycurtis", "hamming", "jaccard")

_c_i64 = ctypes.c_int64
_c_dbl = ctypes.c_double
_p = ctypes.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


class _Lib:
    def __init__(self, impl):
        self.impl = impl
        if impl == "port":
            path = os.path.join(HERE, "_build", "liboracle.so")
            if not os.path.exists(path):
                from . import build_oracle
                build_oracle.build_restatement()
        elif impl == "reference":
            path = os.path.join(HERE, "_ref", "libref.so")
            if not os.path.exists(path):
                from . import build_oracle
                if build_oracle.build_reference() is None:
                    raise FileNotFoundError(
                        "oracle/_ref/libref.so is absent and /root/reference is "
                        "not available to build it")
        else:
            raise ValueError(impl)
        self.lib = ctypes.CDLL(path)
        self.prefix = "oracle_" if impl == "port" else "ref_"

    def pylib(self):
        """Same library through PyDLL (GIL held) for entry points that call back into Python."""
        if getattr(self, "_pylib", None) is None:
            self._pylib = ctypes.PyDLL(self.lib._name)
        return self._pylib

    def fn(self, name, restype):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def metric_arg(self, metric):
        if metric not in VECTOR_METRICS:
            raise ValueError("metric must be one of %s" %
                             ", ".join("'%s'" % s for s in VECTOR_METRICS))
        if self.impl == "port":
            return ctypes.c_int(VECTOR_METRICS.index(metric))
        return ctypes.c_char_p(metric.encode())


_LIBS = {}


def have_reference():
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so")) or \
        os.path.isdir("/root/reference/msmbuilder/libdistance/src")


def get(impl="port"):
    if impl not in _LIBS:
        _LIBS[impl] = _Lib(impl)
    return _LIBS[impl]


def _suffix(*arrays):
    dts = {a.dtype for a in arrays}
    if dts == {np.dtype(np.float32)}:
        return "f32"
    if dts == {np.dtype(np.float64)}:
        return "f64"
    raise TypeError("X and y must be both float32 or float64")


def _c2(a):
    a = np.ascontiguousarray(a)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    return a


def _idx(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def dist(X, y, metric, X_indices=None, impl="port"):
    L = get(impl)
    X = _c2(X)
    y = np.ascontiguousarray(y)
    s = _suffix(X, y)
    rows = _idx(X_indices)
    n_out = len(X) if rows is None else len(rows)
    out = np.zeros(n_out, dtype=np.float64)
    m = L.metric_arg(metric)
    f = L.fn("dist_" + s, None if impl == "reference" else ctypes.c_int)
    if impl == "port":
        f(_ptr(X), _ptr(y), m, _c_i64(X.shape[0]), _c_i64(X.shape[1]),
          _ptr(rows), _c_i64(0 if rows is None else len(rows)), _ptr(out))
    else:
        f(_ptr(X), _ptr(y), m, _c_i64(X.shape[0]), _c_i64(X.shape[1]),
          _ptr(rows), _c_i64(0 if rows is None else len(rows)), _ptr(out))
    return out


def assign_nearest(X, Y, metric, X_indices=None, impl="port"):
    L = get(impl)
    X = _c2(X)
    Y = _c2(Y)
    s = _suffix(X, Y)
    rows = _idx(X_indices)
    n_out = len(X) if rows is None else len(rows)
    assign = np.zeros(n_out, dtype=np.int64)
    f = L.fn("assign_nearest_" + s, _c_dbl)
    inertia = f(_ptr(X), _ptr(Y), L.metric_arg(metric), _ptr(rows),
                _c_i64(X.shape[0]), _c_i64(Y.shape[0]), _c_i64(X.shape[1]),
                _c_i64(0 if rows is None else len(rows)), _ptr(assign))
    return assign.astype(np.intp, copy=False), float(inertia)


def cdist(XA, XB, metric, impl="port"):
    L = get(impl)
    XA = _c2(XA)
    XB = _c2(XB)
    s = _suffix(XA, XB)
    out = np.zeros((XA.shape[0], XB.shape[0]), dtype=np.float64)
    f = L.fn("cdist_" + s, None if impl == "reference" else ctypes.c_int)
    f(_ptr(XA), _ptr(XB), L.metric_arg(metric), _c_i64(XA.shape[0]),
      _c_i64(XB.shape[0]), _c_i64(XA.shape[1]), _ptr(out))
    return out


def pdist(X, metric, X_indices=None, impl="port"):
    L = get(impl)
    X = _c2(X)
    s = _suffix(X)
    rows = _idx(X_indices)
    n = len(X) if rows is None else len(rows)
    out = np.zeros(n * (n - 1) // 2, dtype=np.float64)
    f = L.fn("pdist_" + s, None if impl == "reference" else ctypes.c_int)
    f(_ptr(X), L.metric_arg(metric), _c_i64(X.shape[0]), _c_i64(X.shape[1]),
      _ptr(rows), _c_i64(0 if rows is None else len(rows)), _ptr(out))
    return out


def sumdist(X, metric, pair_indices, impl="port"):
    L = get(impl)
    X = _c2(X)
    s = _suffix(X)
    pairs = np.ascontiguousarray(pair_indices, dtype=np.int64)
    f = L.fn("sumdist_" + s, _c_dbl)
    return float(f(_ptr(X), L.metric_arg(metric), _c_i64(X.shape[0]),
                   _c_i64(X.shape[1]), _ptr(pairs), _c_i64(pairs.shape[0])))


def random_assignment(n_clusters, n_elements, random):
    """kmedoids.cc:314-383: cluster sizes from successive binomials (each cluster
    keeps at least one element), then one shuffle.  Same RandomState calls, same order."""
    cid = np.zeros(n_elements, dtype=np.int64)
    n = n_elements - n_clusters
    k = 0
    for i in range(n_clusters - 1):
        j = int(random.binomial(float(n), 1.0 / (n_clusters - i)))
        n -= j
        j += k + 1
        cid[k:j] = i
        k = j
    cid[k:] = n_clusters - 1
    random.shuffle(cid)
    return cid


def kmedoids(n_clusters, distmatrix, n_pass, clusterid=None, random_state=None,
             impl="port"):
    """_kmedoids.pyx:23-107.  impl='reference' runs the reference C++ for every
    n_pass; impl='port' restates the restart loop (kmedoids.cc:181-250) in Python
    over the C port of one descent."""
    from sklearn.utils import check_random_state
    dm = np.ascontiguousarray(distmatrix, dtype=np.float64)
    n_elements = int(1 + np.sqrt(8 * len(dm) + 1) / 2.0)
    if len(dm) != (n_elements * (n_elements - 1) / 2):
        raise ValueError("len(distmatrix)=%s is not a valid size of a condensed "
                         "distance matrix" % len(dm))
    if n_clusters > n_elements:
        raise ValueError("Number of clusters requested (%d) greater than "
                         "number of elements (%d)" % (n_clusters, n_elements))
    if clusterid is not None and len(clusterid) != n_elements:
        raise ValueError("clusterid must be None or an array of length n_elements")
    if n_pass < 0:
        raise ValueError("n_pass must be greater than or equal to zero.")
    if clusterid is None:
        cid = np.zeros(n_elements, dtype=np.int64)
    else:
        cid = np.array(clusterid, dtype=np.int64, copy=True)
    random = check_random_state(random_state)
    err = _c_dbl(0.0)
    L = get(impl)
    if impl == "reference":
        if n_pass == 0:
            ifound = L.fn("kmedoids_npass0", _c_i64)(
                _c_i64(n_clusters), _c_i64(n_elements), _ptr(dm), _ptr(cid), ctypes.byref(err))
        else:
            f = getattr(L.pylib(), "ref_kmedoids_npass")
            f.restype = _c_i64
            ifound = f(_c_i64(n_clusters), _c_i64(n_elements), _ptr(dm), _c_i64(n_pass),
                       _ptr(cid), ctypes.py_object(random), ctypes.byref(err))
        return cid.astype(np.intp, copy=False), err.value, int(ifound)

    descent = L.fn("kmedoids", ctypes.c_int)
    if n_pass == 0:
        ifound = descent(_c_i64(n_clusters), _c_i64(n_elements), _ptr(dm), _ptr(cid),
                         ctypes.byref(err))
        return cid.astype(np.intp, copy=False), err.value, int(ifound)
    if n_pass == 1:
        # tclusterid IS clusterid (kmedoids.cc:165-166): one descent from a random start
        cid = random_assignment(n_clusters, n_elements, random)
        ifound = descent(_c_i64(n_clusters), _c_i64(n_elements), _ptr(dm), _ptr(cid),
                         ctypes.byref(err))
        return cid.astype(np.intp, copy=False), err.value, int(ifound)
    best = float(np.finfo(np.float64).max)
    ifound = -1
    for _ in range(n_pass):
        t = random_assignment(n_clusters, n_elements, random)
        descent(_c_i64(n_clusters), _c_i64(n_elements), _ptr(dm), _ptr(t), ctypes.byref(err))
        total = 0.0
        for i in range(n_elements):        # same summation order as kmedoids.cc:213-233
            if t[i] != i:
                total += dm[condensed_index(i, int(t[i]), n_elements)]
        # kmedoids.cc:237-249: t now holds centroid ids; compare with the kept solution
        if np.array_equal(cid, t):
            ifound += 1
        elif total < best:
            ifound = 1
            best = total
            cid[:] = t
    return cid.astype(np.intp, copy=False), best, int(ifound)


def contigify_ids(ids, impl="port"):
    """_kmedoids.pyx:110-117: in-place relabel, returns (ids, {old: new})."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    keys = np.zeros(max(len(ids), 1), dtype=np.int64)
    L = get(impl)
    n = L.fn("contigify_ids", _c_i64)(_ptr(ids), _c_i64(len(ids)), _ptr(keys))
    return ids.astype(np.intp, copy=False), {int(keys[r]): r for r in range(int(n))}


def condensed_index(i, j, n):
    f = get("port").fn("condensed_index", _c_i64)
    return int(f(_c_i64(i), _c_i64(j), _c_i64(n)))
