"""GPU parity of the DEFAULT tICA engine at the BASELINE.json sizes and on awkward data.

The default engine (engine='auto' -> tcgen05, fp16 split of the centred, power-of-two scaled frames)
is approximate where the reference is float64 (msmbuilder/decomposition/tica.py:402).  Its contract
is BASELINE.json's: eigenvalues within 1e-5 of float64, components within cos >= 1 - 1e-4.  These
tests hold it to that against the float64 CUDA-core engine (the reference's arithmetic; itself pinned
to the oracle and the reference-written goldens in test_gpu_tica.py) at

  * config 2:  10M x 64 float32, lag 10                         (BASELINE.json configs[1])
  * config 4:  a 12.8M-frame slice of the 50M x 256 workload    (the bench line carries the full-size
               number: bench.py check.eig_err_vs_f64)
  * heavy-tailed, drifting and bursty features (the shift / scale sample must not be fooled).
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EIG_ATOL = 1e-5          # BASELINE.json: "eigenvalues vs reference within 1e-5"
COS_MIN = 1 - 1e-4


def _fit_pair(seqs, lag=10, k=4, engine="auto"):
    from msmbuilder_b200.decomposition import tICA
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = tICA(n_components=k, lag_time=lag, engine="simt_f64").fit(seqs)
        b = tICA(n_components=k, lag_time=lag, engine=engine).fit(seqs)
    return a, b


def _check(a, b, eig_atol=EIG_ATOL):
    assert a.n_observations_ == b.n_observations_ and a.n_sequences_ == b.n_sequences_
    err = float(np.abs(a.eigenvalues_ - b.eigenvalues_).max())
    assert err <= eig_atol, "eigenvalues differ from float64 by %g" % err
    cos = np.abs(np.sum(a.components_ * b.components_, axis=1)) / (
        np.linalg.norm(a.components_, axis=1) * np.linalg.norm(b.components_, axis=1))
    assert cos.min() >= COS_MIN, "component cosine %g" % cos.min()
    np.testing.assert_allclose(b.means_, a.means_, rtol=0,
                               atol=1e-6 * max(1.0, float(np.abs(a.means_).max())))
    return err


def test_config2_10M_x_64_default_engine_vs_float64():
    # BASELINE.json configs[1]: 10M frames x 64 features float32, eigenvalues within 1e-5
    from msmbuilder_b200.synthetic import ar1_device
    n_seq, L, D = 100, 100_000, 64
    X = ar1_device(n_seq, L, D, seed=2000)
    seqs = [X[i * L:(i + 1) * L] for i in range(n_seq)]
    a, b = _fit_pair(seqs)
    assert a.n_observations_ == 10_000_000
    _check(a, b)


def test_config4_slice_12p8M_x_256_default_engine_vs_float64():
    # the bench workload's generator and shape; 128 of its 500 sequences (float64 engine: ~0.6 s)
    from msmbuilder_b200.synthetic import ar1_device
    n_seq, L, D = 128, 100_000, 256
    X = ar1_device(n_seq, L, D, seed=1000)
    seqs = [X[i * L:(i + 1) * L] for i in range(n_seq)]
    a, b = _fit_pair(seqs)
    _check(a, b)
    # sharding must not move the result beyond the tolerance either: two halves, added
    from msmbuilder_b200.decomposition import tICA
    c = tICA(n_components=4, lag_time=10)
    c._initialize(D)
    for part in (seqs[:64], seqs[64:]):
        c._add_packed(c._accumulate_device(part).cpu().numpy())
    assert float(np.abs(c.eigenvalues_ - a.eigenvalues_).max()) <= EIG_ATOL


def _awkward(kind, n_seq=6, L=40_000, D=64, seed=7):
    from msmbuilder_b200.synthetic import ar1_numpy
    rs = np.random.RandomState(seed)
    seqs = ar1_numpy(n_seq, L, D, seed=seed)
    out = []
    for i, s in enumerate(seqs):
        s = s.astype(np.float64)
        if kind == "heavy_tailed":
            # Student-t(3)-like marginals: rare values far outside the bulk
            s = s * (1.0 + np.abs(rs.standard_t(3, size=s.shape)))
        elif kind == "early_outlier":
            # one huge value inside the very first rows, then ordinary data: a sample taken only
            # from the head would set a scale that wipes out the rest
            if i == 0:
                s[3, ::4] += 5.0e4
        elif kind == "drift":
            # non-stationary: the mean wanders by 50 sigma and the amplitude grows 30-fold
            t = np.linspace(0, 1, len(s))[:, None]
            s = s * (1.0 + 29.0 * t) + 50.0 * np.sin(3.0 * t * (i + 1))
        elif kind == "late_burst":
            # quiet at first, a burst 300x larger in the last sequence only
            if i == n_seq - 1:
                s[L // 2:L // 2 + 500] *= 300.0
        elif kind == "tiny_then_normal":
            # the first sequence is 1e-4 of the others
            if i == 0:
                s *= 1.0e-4
        out.append(s.astype(np.float32))
    return out


@pytest.mark.parametrize("kind", ["heavy_tailed", "early_outlier", "drift", "late_burst",
                                  "tiny_then_normal"])
def test_default_engine_on_awkward_features(kind):
    seqs = _awkward(kind)
    a, b = _fit_pair(seqs)
    _check(a, b)
    # the moments themselves, per unit of variance (what the eigenproblem feels)
    sd = np.sqrt(np.diag(a.covariance_))
    np.testing.assert_allclose(b.covariance_ / np.outer(sd, sd), a.covariance_ / np.outer(sd, sd),
                               rtol=0, atol=2e-5)
    np.testing.assert_allclose(b.offset_correlation_ / np.outer(sd, sd),
                               a.offset_correlation_ / np.outer(sd, sd), rtol=0, atol=2e-5)


def test_config1_dihedral_standin_runs_on_the_device_too():
    # BASELINE.json configs[0] is CPU plumbing (tests/test_config1_plumbing.py); here the same
    # D = 4 input goes through the C ABI (float64 CUDA-core engine: D < 32) and must match the oracle
    from oracle.tica_oracle import TicaOracle
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.synthetic import dihedral_standin_numpy
    seqs = dihedral_standin_numpy()
    ref = TicaOracle(n_components=4, lag_time=10).fit(seqs)
    m = tICA(n_components=4, lag_time=10).fit(seqs)
    np.testing.assert_allclose(m.eigenvalues_, ref.eigenvalues_, rtol=0, atol=1e-9)
    np.testing.assert_allclose(m.means_, ref.means_, rtol=1e-10, atol=1e-12)
