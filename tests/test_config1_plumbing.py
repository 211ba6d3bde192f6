"""BASELINE.json configs[0]: tICA(lag_time=10, n_components=4) on AlanineDipeptide dihedral
features, reference CPU/NumPy path -- "plumbing, no GPU".

The dataset needs mdtraj and a figshare download (msmbuilder/example_datasets/
alanine_dipeptide.py:18-54, featurizer.py:555-657), neither available offline; the seeded stand-in
`synthetic.dihedral_standin_numpy` has its shape (10 trajectories x 9,999 frames x
[sin phi, cos phi, sin psi, cos psi]) and two metastable angular coordinates.  The golden
tests/golden/tica_config1_dihedral.npz was written by the reference's own tica.py loaded verbatim
(oracle/gen_golden.py gen_config1).  Checked here, on the CPU:

  * the oracle restatement (oracle/tica_oracle.py) reproduces the reference's eigenvalues,
    eigenvectors, means, matrices, projections, score and summarize() text on this input;
  * the estimator's host-side algebra (moments -> RBLW shrinkage -> eigenproblem -> properties,
    decomposition/tica.py) does too when it is handed the oracle's sufficient statistics -- the
    only part of the estimator that is not in this test is the device accumulation, which
    tests/test_gpu_tica_at_size.py::test_config1_dihedral_standin_runs_on_the_device_too covers.
"""
import os
import warnings

import numpy as np
import pytest

from msmbuilder_b200.synthetic import dihedral_standin_numpy
from oracle.tica_oracle import TicaOracle


@pytest.fixture(scope="module")
def case(golden_dir):
    g = np.load(os.path.join(golden_dir, "tica_config1_dihedral.npz"))
    seqs = dihedral_standin_numpy()
    assert len(seqs) == 10 and all(s.shape == (9999, 4) for s in seqs)
    return g, seqs


def _same_up_to_sign(a, b, atol):
    sign = np.sign(np.sum(a * b, axis=0))
    np.testing.assert_allclose(a * sign, b, rtol=0, atol=atol)


def test_standin_is_metastable(case):
    g, seqs = case
    # two slow processes (the two double wells), then a gap: what makes the config meaningful
    ev = g["eigenvalues"]
    assert ev[0] > 0.95 and ev[1] > 0.95 and ev[2] < 0.7


def test_oracle_reproduces_the_reference_on_config1(case):
    g, seqs = case
    m = TicaOracle(n_components=4, lag_time=10).fit(seqs)
    assert m.n_observations_ == int(g["n_observations"]) == 99990
    assert m.n_sequences_ == int(g["n_sequences"]) == 10
    np.testing.assert_allclose(m.eigenvalues_, g["eigenvalues"], rtol=0, atol=1e-12)
    _same_up_to_sign(m.eigenvectors_, g["eigenvectors"], 1e-9)
    np.testing.assert_allclose(m.means_, g["means"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.covariance_, g["covariance"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.offset_correlation_, g["offset_correlation"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.timescales_, g["timescales"], rtol=1e-9)
    assert abs(m.shrinkage_ - float(g["shrinkage_"])) < 1e-15


def test_estimator_host_algebra_reproduces_the_reference_on_config1(case):
    g, seqs = case
    from msmbuilder_b200.decomposition import tICA
    ref = TicaOracle(n_components=4, lag_time=10).fit(seqs)
    m = tICA(n_components=4, lag_time=10)
    m._initialize(4)
    m._add_packed(ref.packed_moments())          # the device accumulator's layout (include/msmb200.h)
    assert m.n_observations_ == 99990 and m.n_sequences_ == 10
    np.testing.assert_allclose(m.eigenvalues_, g["eigenvalues"], rtol=0, atol=1e-12)
    _same_up_to_sign(m.eigenvectors_, g["eigenvectors"], 1e-9)
    np.testing.assert_allclose(m.means_, g["means"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.covariance_, g["covariance"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.timescales_, g["timescales"], rtol=1e-9)
    assert m.components_.shape == (4, 4)
    # summarize() is the CLI's output (commands/fit_transform.py:66-102): same text
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert m.summarize() == str(g["summary"])


def test_oracle_projection_and_score_on_config1(case):
    g, seqs = case
    m = TicaOracle(n_components=4, lag_time=10).fit(seqs)
    proj = m.transform([seqs[0][:50]])[0]
    sign = np.sign(np.sum(proj * g["proj50"], axis=0))
    np.testing.assert_allclose(proj * sign, g["proj50"], rtol=0, atol=1e-9)
    assert abs(m.score(seqs[:3]) - float(g["score"])) < 1e-9
    mk = TicaOracle(n_components=2, lag_time=10, kinetic_mapping=True).fit(seqs)
    pk = mk.transform([seqs[0][:50]])[0]
    sign = np.sign(np.sum(pk * g["proj50_kinetic"], axis=0))
    np.testing.assert_allclose(pk * sign, g["proj50_kinetic"], rtol=0, atol=1e-9)
