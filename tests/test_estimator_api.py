"""CPU: drop-in boundary checks that need no device -- constructor signatures,
clone / pickle / get_params, numpydoc Parameters == __init__ (the reference's CLI
derives flags from that: msmbuilder/cmdline.py:334-384), base class
(tests/test_estimator_subclassing.py:52-55), error behaviour."""
import inspect
import os
import pickle
import re

import numpy as np
import pytest
from sklearn.base import BaseEstimator as SkBase, clone

from msmbuilder_b200.base import BaseEstimator
from msmbuilder_b200.cluster import (KCenters, MiniBatchKMedoids, MiniBatchKMeans, RegularSpatial,
                                     KMedoids)
from msmbuilder_b200.decomposition import tICA
from msmbuilder_b200.utils import check_iter_of_sequences, array2d

REF_SIGNATURES = {
    # reference: tica.py:108-109, kcenters.py:74, minibatchkmedoids.py:81-82
    tICA: ["n_components", "lag_time", "shrinkage", "kinetic_mapping", "commute_mapping"],
    KCenters: ["n_clusters", "metric", "random_state"],
    MiniBatchKMedoids: ["n_clusters", "max_iter", "batch_size", "metric", "max_no_improvement",
                        "random_state"],
    # regularspatial.py:65, kmedoids.py:73-74
    RegularSpatial: ["d_min", "metric"],
    KMedoids: ["n_clusters", "n_passes", "metric", "random_state"],
}
REF_DEFAULTS = {
    tICA: dict(n_components=None, lag_time=1, shrinkage=None, kinetic_mapping=False,
               commute_mapping=False),
    KCenters: dict(n_clusters=8, metric='euclidean', random_state=None),
    MiniBatchKMedoids: dict(n_clusters=8, max_iter=5, batch_size=100, metric='euclidean',
                            max_no_improvement=10, random_state=None),
    RegularSpatial: dict(metric='euclidean'),
    KMedoids: dict(n_clusters=8, n_passes=1, metric='euclidean', random_state=None),
}


@pytest.mark.parametrize("cls", [tICA, KCenters, MiniBatchKMedoids, RegularSpatial, KMedoids])
def test_signature_matches_reference(cls):
    params = inspect.signature(cls.__init__).parameters
    names = [p for p in params if p != "self"]
    assert names[:len(REF_SIGNATURES[cls])] == REF_SIGNATURES[cls]
    for k, v in REF_DEFAULTS[cls].items():
        assert params[k].default == v
    for extra in names[len(REF_SIGNATURES[cls]):]:
        assert params[extra].default is not inspect.Parameter.empty


@pytest.mark.parametrize("cls", [tICA, KCenters, MiniBatchKMedoids, MiniBatchKMeans, RegularSpatial,
                                 KMedoids])
def test_clone_pickle_base(cls):
    est = cls(d_min=1.0) if cls is RegularSpatial else cls()
    assert isinstance(est, BaseEstimator) and isinstance(est, SkBase)
    c = clone(est)
    assert c.get_params() == est.get_params()
    pickle.loads(pickle.dumps(est))
    assert isinstance(est.summarize() if cls is MiniBatchKMeans else "x", str)


@pytest.mark.parametrize("cls", [tICA, KCenters, MiniBatchKMedoids, RegularSpatial, KMedoids])
def test_numpydoc_parameters_cover_init(cls):
    doc = cls.__doc__
    sect = doc[doc.index("Parameters"):doc.index("Attributes")]
    documented = re.findall(r"^\s{4}(\w+) : ", sect, flags=re.M)
    names = [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
    assert documented == names


def test_tica_rejects_both_mappings():
    with pytest.raises(ValueError):
        tICA(kinetic_mapping=True, commute_mapping=True)


def test_tica_solve_before_fit():
    t = tICA()
    t._initialize(3)
    with pytest.raises(RuntimeError):
        t.eigenvalues_


def test_validation_helpers():
    check_iter_of_sequences([np.zeros((3, 2)), np.zeros((5, 2))])
    with pytest.raises(ValueError):
        check_iter_of_sequences([np.zeros(3)])
    with pytest.raises(ValueError):
        check_iter_of_sequences([np.zeros((3, 2, 1))])
    check_iter_of_sequences([np.zeros((3, 4, 3))], allow_trajectory=True)
    with pytest.raises(ValueError):
        array2d(np.array([[1.0, np.nan]]))
    assert array2d([1.0, 2.0]).shape == (1, 2)


def test_tica_state_names_and_packed_fold():
    # the private accumulator names are part of the de-facto ABI (tica.py:123-148)
    t = tICA(lag_time=2)
    t._initialize(2)
    for name in ("_outer_0_to_T_lagged", "_sum_0_to_TminusTau", "_sum_tau_to_T", "_sum_0_to_T",
                 "_outer_0_to_TminusTau", "_outer_offset_to_T", "_initialized", "_is_dirty"):
        assert hasattr(t, name)
    from oracle.tica_oracle import TicaOracle
    X = np.random.RandomState(0).randn(50, 2)
    o = TicaOracle(lag_time=2).fit([X])
    t._add_packed(o.packed_moments())
    np.testing.assert_array_equal(t._outer_0_to_T_lagged, o._outer_0_to_T_lagged)
    np.testing.assert_array_equal(t._outer_offset_to_T, o._outer_offset_to_T)
    assert t.n_observations_ == 50 and t.n_sequences_ == 1
    np.testing.assert_allclose(t.eigenvalues_, o.eigenvalues_, atol=1e-14)
    np.testing.assert_allclose(t.means_, o.means_)
    bad = o.packed_moments()
    bad[-3] = np.nan
    with pytest.raises(ValueError):
        t._add_packed(bad)


def test_npy_stream_host_side(tmp_path):
    """io.NumpyDirStream without a GPU: file discovery, header shapes, validation."""
    from msmbuilder_b200.io import NumpyDirStream, save_sequences
    from msmbuilder_b200.utils import check_iter_of_sequences
    from msmbuilder_b200 import _lib
    seqs = [np.zeros((5, 3), np.float32), np.ones((2, 3), np.float64), np.zeros((0, 3), np.float32)]
    path = save_sequences(str(tmp_path / "ds"), seqs)
    open(os.path.join(path, "PROVENANCE.txt"), "w").write("x")        # ignored like dataset.py:329
    s = NumpyDirStream(path)
    assert len(s) == 3 and s.keys() == [0, 1, 2]
    assert [sh for sh, _ in s.shapes()] == [(5, 3), (2, 3), (0, 3)]
    check_iter_of_sequences(s)
    save_sequences(str(tmp_path / "bad"), [np.zeros(4)])
    with pytest.raises(ValueError):
        check_iter_of_sequences(NumpyDirStream(str(tmp_path / "bad")))
    with pytest.raises(ValueError):
        NumpyDirStream(str(tmp_path))                                   # no .npy items
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_lib.Msmb200Error):
            next(iter(s))
