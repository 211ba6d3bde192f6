"""oracle/ref_loader.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Loads the reference's OWN hot-path Python files *verbatim* from /root/reference
under a handful of stub modules, so that the restatements
in oracle/ can be pinned against the real thing in the build container and so
that oracle/gen_golden.py can write fixtures from it.  `import msmbuilder` itself
is impossible here (mdtraj absent; py3.12 / SciPy 1.18 incompatibilities), but:

  * decomposition/tica.py loads unmodified given stub packages and a keyword
    shim scipy.linalg.eigh(eigvals=(lo,hi)) -> subset_by_index=(lo,hi)
    (tica.py:188-189 uses the keyword removed in SciPy 1.14);
  * base.py, utils/validation.py, cluster/base.py, cluster/kcenters.py and
    cluster/minibatchkmedoids.py load unmodified given a stub `mdtraj`
    (class Trajectory) and stand-ins for the two Cython modules
    `msmbuilder.libdistance` / `msmbuilder.cluster._kmedoids`, which here are
    backed by the reference's own C++ compiled through oracle/ref_shim.cc.

Nothing is copied: the files are exec'd from where they lie.  /root/reference
does not exist on the GPU box, so everything here is guarded by `available()`.
"""
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("MSMB_REFERENCE_ROOT", "/root/reference")
_PKG = os.path.join(REFERENCE_ROOT, "msmbuilder")

_loaded = {}


def available():
    return os.path.isfile(os.path.join(_PKG, "decomposition", "tica.py"))


def _stub_package(name):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # mark as package
        sys.modules[name] = mod
    return mod


def _load_file(modname, relpath):
    if modname in sys.modules and getattr(sys.modules[modname], "__ref_verbatim__", False):
        return sys.modules[modname]
    path = os.path.join(_PKG, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)  # '\l' in tica.py:27 docstring
        spec.loader.exec_module(mod)
    mod.__ref_verbatim__ = True
    return mod


def _install_eigh_shim():
    import scipy.linalg
    if getattr(scipy.linalg.eigh, "__msmb_shim__", False):
        return
    real = scipy.linalg.eigh

    def eigh(a, b=None, *args, **kw):
        if "eigvals" in kw:
            ev = kw.pop("eigvals")
            if ev is not None:
                kw["subset_by_index"] = tuple(ev)
        return real(a, b, *args, **kw)

    eigh.__msmb_shim__ = True
    eigh.__wrapped__ = real
    scipy.linalg.eigh = eigh


def _common():
    if "common" in _loaded:
        return
    if not available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    if "mdtraj" not in sys.modules:
        md = types.ModuleType("mdtraj")

        class Trajectory(object):
            pass

        md.Trajectory = Trajectory
        md.__stub__ = True
        sys.modules["mdtraj"] = md
    if "msmbuilder" in sys.modules and not getattr(sys.modules["msmbuilder"], "__stub__", False):
        raise RuntimeError("a real msmbuilder is importable; ref_loader is for when it is not")
    top = _stub_package("msmbuilder")
    top.__stub__ = True
    top.base = _load_file("msmbuilder.base", "base.py")
    validation = _load_file("msmbuilder.utils.validation", os.path.join("utils", "validation.py"))
    utils = _stub_package("msmbuilder.utils")
    utils.check_iter_of_sequences = validation.check_iter_of_sequences
    utils.array2d = validation.array2d
    utils.list_of_1d = validation.list_of_1d
    top.utils = utils
    _loaded["common"] = True


def load_tica():
    """Returns the reference's tICA class (decomposition/tica.py:26), unmodified."""
    if "tica" not in _loaded:
        _common()
        _install_eigh_shim()
        _stub_package("msmbuilder.decomposition")
        mod = _load_file("msmbuilder.decomposition.tica", os.path.join("decomposition", "tica.py"))
        _loaded["tica"] = mod
    return _loaded["tica"].tICA


def _libdistance_standin():
    """`msmbuilder.libdistance` backed by the compiled reference C++ (no RMSD)."""
    from . import libdistance_oracle as lo
    mod = types.ModuleType("msmbuilder.libdistance")

    def _m(metric):
        return metric.decode() if isinstance(metric, bytes) else metric

    mod.assign_nearest = lambda X, Y, metric, X_indices=None: \
        lo.assign_nearest(X, Y, _m(metric), X_indices, impl="reference")
    mod.dist = lambda X, y, metric, X_indices=None: \
        lo.dist(X, y, _m(metric), X_indices, impl="reference")
    mod.pdist = lambda X, metric, X_indices=None: \
        lo.pdist(X, _m(metric), X_indices, impl="reference")
    mod.cdist = lambda XA, XB, metric: lo.cdist(XA, XB, _m(metric), impl="reference")
    mod.sumdist = lambda X, metric, pair_indices: \
        lo.sumdist(X, _m(metric), pair_indices, impl="reference")
    return mod


def _kmedoids_standin():
    from . import libdistance_oracle as lo
    mod = types.ModuleType("msmbuilder.cluster._kmedoids")
    mod.kmedoids = lambda n_clusters, distmatrix, n_pass, clusterid=None, random_state=None: \
        lo.kmedoids(n_clusters, distmatrix, n_pass, clusterid, random_state, impl="reference")
    mod.contigify_ids = lambda ids: lo.contigify_ids(ids, impl="reference")
    return mod


def load_cluster():
    """Returns (KCenters, MiniBatchKMedoids, MultiSequenceClusterMixin): the
    reference classes from cluster/kcenters.py:132, cluster/minibatchkmedoids.py:170
    and cluster/base.py:17, unmodified."""
    if "cluster" not in _loaded:
        _common()
        top = sys.modules["msmbuilder"]
        libd = _libdistance_standin()
        sys.modules["msmbuilder.libdistance"] = libd
        top.libdistance = libd
        base = _load_file("msmbuilder.cluster.base", os.path.join("cluster", "base.py"))
        pkg = _stub_package("msmbuilder.cluster")
        pkg.MultiSequenceClusterMixin = base.MultiSequenceClusterMixin
        km = _kmedoids_standin()
        sys.modules["msmbuilder.cluster._kmedoids"] = km
        pkg._kmedoids = km
        kc = _load_file("msmbuilder.cluster.kcenters", os.path.join("cluster", "kcenters.py"))
        mb = _load_file("msmbuilder.cluster.minibatchkmedoids",
                        os.path.join("cluster", "minibatchkmedoids.py"))
        _loaded["cluster"] = (kc.KCenters, mb.MiniBatchKMedoids, base.MultiSequenceClusterMixin)
    return _loaded["cluster"]


def _numpy_aliases():
    """msm/core.py:560-562 uses np.int / np.float (gone since NumPy 1.24)."""
    import numpy as np
    for name, typ in (("int", int), ("float", float)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if not hasattr(np, "row_stack"):
        np.row_stack = np.vstack


def load_transition_counts():
    """Returns the reference's `_transition_counts` (msm/core.py:487-602), unmodified.
    `msmbuilder.msm._ratematrix` (Cython, not on this path) is an empty stand-in."""
    if "msm_core" not in _loaded:
        _common()
        _numpy_aliases()
        pkg = _stub_package("msmbuilder.msm")
        rm = types.ModuleType("msmbuilder.msm._ratematrix")
        sys.modules["msmbuilder.msm._ratematrix"] = rm
        pkg._ratematrix = rm
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = _load_file("msmbuilder.msm.core", os.path.join("msm", "core.py"))
        _loaded["msm_core"] = mod
    return _loaded["msm_core"]._transition_counts


def load_more_clusterers():
    """Returns (RegularSpatial, KMedoids): cluster/regularspatial.py:106 and
    cluster/kmedoids.py:144, unmodified, over the same stand-ins as load_cluster()."""
    if "cluster_more" not in _loaded:
        load_cluster()
        rs = _load_file("msmbuilder.cluster.regularspatial", os.path.join("cluster", "regularspatial.py"))
        km = _load_file("msmbuilder.cluster.kmedoids", os.path.join("cluster", "kmedoids.py"))
        _loaded["cluster_more"] = (rs.RegularSpatial, km.KMedoids)
    return _loaded["cluster_more"]


def load_agglomerative():
    """Returns the reference's LandmarkAgglomerative (cluster/agglomerative.py:296), unmodified.
    `fastcluster.linkage` (absent here) is SciPy's `linkage`: fastcluster is a drop-in
    reimplementation of that function with the same stepwise-dendrogram output; np.infty
    (agglomerative.py:252, gone in NumPy 2) is aliased to np.inf."""
    if "agglomerative" not in _loaded:
        load_cluster()
        import numpy as np
        import scipy.cluster.hierarchy
        if "fastcluster" not in sys.modules:
            fc = types.ModuleType("fastcluster")
            fc.linkage = scipy.cluster.hierarchy.linkage
            fc.__stub__ = True
            sys.modules["fastcluster"] = fc
        if not hasattr(np, "infty"):
            np.infty = np.inf
        ag = _load_file("msmbuilder.cluster.agglomerative", os.path.join("cluster", "agglomerative.py"))
        _loaded["agglomerative"] = ag.LandmarkAgglomerative
    return _loaded["agglomerative"]
