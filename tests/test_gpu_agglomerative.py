"""GPU parity: LandmarkAgglomerative (fit on device pdist + host linkage, predict through
msmb200_cdist + msmb200_pooled_assign) against tests/golden/agglomerative.npz, written by the
reference's own cluster/agglomerative.py over its compiled libdistance (oracle/gen_golden.py
gen_agglomerative); RMSDFeaturizer (featurizer.py:255-321) against the QCP oracle."""
import ast
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cluster_inputs(seed, n_seq, length, D, dtype):          # oracle/gen_golden.py cluster_inputs
    rs = np.random.RandomState(seed)
    centers = rs.randn(7, D) * 3
    return [(centers[rs.randint(0, 7, size=length)] + rs.randn(length, D)).astype(dtype)
            for _ in range(n_seq)]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "agglomerative.npz"))


@pytest.mark.parametrize("ci", range(10))
def test_landmark_agglomerative_matches_reference(gold, ci):
    from msmbuilder_b200.cluster import LandmarkAgglomerative
    assert int(gold["n_cases"]) == 10
    linkage, metric, dtype, wp = ast.literal_eval(str(gold["cases"][ci]))
    kw = dict(ast.literal_eval(str(gold["ag%d_kw" % ci])))
    seqs = _cluster_inputs(40 + ci, 3, 300, 6, np.dtype(dtype))
    new = _cluster_inputs(90 + ci, 2, 250, 6, np.dtype(dtype))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LandmarkAgglomerative(**kw).fit(seqs)
        pred = m.predict(new)
    key = "ag%d_" % ci
    np.testing.assert_array_equal(m.landmarks_, gold[key + "landmarks"])
    np.testing.assert_array_equal(m.landmark_labels_, gold[key + "landmark_labels"])
    np.testing.assert_array_equal(m.cardinality_, gold[key + "cardinality"])
    np.testing.assert_allclose(m.squared_distances_within_cluster_, gold[key + "sqsum"], rtol=1e-13)
    np.testing.assert_allclose(m.cluster_centers_, gold[key + "centers"], rtol=1e-6)
    got = np.concatenate(pred)
    want = gold[key + "pred"]
    assert got.shape == want.shape and len(pred) == 2 and len(pred[0]) == 250
    # min / max pooling is exact; mean / ward pooling sums in a different order than NumPy
    # (1e-16 relative): a label may only differ on a genuine tie, and there are none here
    np.testing.assert_array_equal(got, want)


def test_landmark_agglomerative_all_frames_are_landmarks():
    # n_landmarks=None: every frame is a landmark; predict(X) of the training frames with single
    # linkage returns each frame's own cluster (distance 0 to itself)
    from msmbuilder_b200.cluster import LandmarkAgglomerative
    seqs = _cluster_inputs(3, 2, 120, 4, np.float32)
    m = LandmarkAgglomerative(n_clusters=4, linkage="single").fit(seqs)
    got = np.concatenate(m.predict(seqs))
    np.testing.assert_array_equal(got, m.landmark_labels_)
    assert m.cluster_centers_.shape == (4, 4)
    with pytest.raises(ValueError):
        LandmarkAgglomerative(n_clusters=4, linkage="median").fit(seqs).predict(seqs)


def test_rmsd_featurizer_matches_qcp_oracle():
    from msmbuilder_b200.featurizer import RMSDFeaturizer
    from msmbuilder_b200.synthetic import rmsd_conformations_numpy
    from oracle import rmsd_oracle as ro
    xyz, _ = rmsd_conformations_numpy(300, n_atoms=23, n_templates=5, seed=4)
    ref, _ = rmsd_conformations_numpy(7, n_atoms=23, n_templates=5, seed=5)
    keep = xyz.copy()
    f = RMSDFeaturizer(ref)
    out = f.transform([xyz[:200], xyz[200:]])
    assert [o.shape for o in out] == [(200, 7), (100, 7)] and out[0].dtype == np.float64
    np.testing.assert_array_equal(xyz, keep)                  # inputs are not centred in place
    want = ro.cdist_rmsd(xyz, ref)
    np.testing.assert_allclose(np.concatenate(out), want, rtol=0, atol=1e-5)
    # atom subset == featurizing the sliced coordinates
    idx = np.array([0, 3, 4, 9, 15, 22])
    g = RMSDFeaturizer(ref, atom_indices=idx)
    np.testing.assert_allclose(g.partial_transform(xyz), ro.cdist_rmsd(xyz[:, idx], ref[:, idx]),
                               rtol=0, atol=1e-5)
    with pytest.raises(ValueError):
        RMSDFeaturizer()
