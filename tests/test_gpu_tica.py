"""GPU parity: tICA estimator (C ABI msmb200_tica_accumulate / _transform) vs the
float64 oracle, the goldens written from the reference's own tica.py, and the
reference's identity tests (msmbuilder/tests/test_decomposition.py:28-125,164-168;
tests/test_utils.py:59-79)."""
import os
import pickle
import warnings

import numpy as np
import pytest

from oracle.tica_oracle import TicaOracle
from msmbuilder_b200.synthetic import ar1_numpy

pytestmark = pytest.mark.gpu

# float64 engine: only the summation order differs from NumPy
ACC_RTOL = 1e-11
EIG_ATOL = 1e-9
# tensor-core engine: BASELINE.json tolerance
UMMA_EIG_ATOL = 1e-5


def _golden_inputs(g):
    seqs = ar1_numpy(int(g["n_seq"]), int(g["length"]), int(g["D"]), seed=int(g["seed"]),
                     dtype=np.dtype(str(g["dtype"])))
    for n_short in g["short"]:
        seqs.insert(1, seqs[0][:int(n_short)].copy())
    return seqs


def _assert_moments_close(m, ref, rtol):
    for name in ("_outer_0_to_T_lagged", "_outer_0_to_TminusTau", "_outer_offset_to_T",
                 "_sum_0_to_TminusTau", "_sum_tau_to_T", "_sum_0_to_T"):
        a, b = getattr(m, name), getattr(ref, name) if not isinstance(ref, dict) else ref[name]
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= rtol * scale, name


@pytest.mark.parametrize("name", ["tica_d6_lag3", "tica_d16_lag10", "tica_d64_lag10_f64",
                                  "tica_d256_lag10"])
@pytest.mark.parametrize("engine", ["simt_f64", "auto"])
def test_matches_golden(golden_dir, name, engine):
    from msmbuilder_b200.decomposition import tICA
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seqs = _golden_inputs(g)
    shrink = None if np.isnan(g["shrinkage"]) else float(g["shrinkage"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = tICA(n_components=int(g["k"]), lag_time=int(g["lag"]), shrinkage=shrink, engine=engine).fit(seqs)
    assert m.n_observations_ == int(g["n_observations"]) and m.n_sequences_ == int(g["n_sequences"])
    ref = {"_outer_0_to_T_lagged": g["C_tau"], "_outer_0_to_TminusTau": g["C_00"],
           "_outer_offset_to_T": g["C_tt"], "_sum_0_to_TminusTau": g["S_0"],
           "_sum_tau_to_T": g["S_tau"], "_sum_0_to_T": g["S"]}
    exact = engine == "simt_f64"
    _assert_moments_close(m, ref, ACC_RTOL if exact else 2e-6)
    np.testing.assert_allclose(m.eigenvalues_, g["eigenvalues"], rtol=0,
                               atol=EIG_ATOL if exact else UMMA_EIG_ATOL)
    np.testing.assert_allclose(m.means_, g["means"], rtol=0, atol=1e-9 if exact else 1e-6)
    proj = m.transform([seqs[0][:50]])[0]
    assert proj.dtype == np.float64 and proj.shape == (50, int(g["k"]))
    sign = np.sign((proj * g["proj50"]).sum(0))
    np.testing.assert_allclose(proj * sign, g["proj50"], atol=1e-6 if exact else 2e-3)


def test_partial_fit_equals_fit_and_pickles():
    from msmbuilder_b200.decomposition import tICA
    seqs = ar1_numpy(3, 900, 8, seed=3)
    a = tICA(n_components=2, lag_time=5, engine="simt_f64").fit(seqs)
    b = tICA(n_components=2, lag_time=5, engine="simt_f64")
    for s in seqs:
        b.partial_fit(s)
    np.testing.assert_allclose(a._outer_0_to_T_lagged, b._outer_0_to_T_lagged, rtol=1e-12)
    np.testing.assert_allclose(a.eigenvalues_, b.eigenvalues_, atol=1e-12)
    c = pickle.loads(pickle.dumps(a))
    np.testing.assert_array_equal(c.eigenvalues_, a.eigenvalues_)
    c.partial_fit(seqs[0])                      # resumable after unpickling
    assert c.n_sequences_ == 4
    assert isinstance(a.summarize(), str)


def test_short_sequences_and_errors():
    from msmbuilder_b200.decomposition import tICA
    rs = np.random.RandomState(0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = tICA(lag_time=10, engine="simt_f64").fit([rs.randn(100, 3), rs.randn(10, 3), rs.randn(4, 3)])
        assert sum("too short" in str(x.message) for x in w) == 2
    assert m.n_sequences_ == 1 and m.n_observations_ == 100        # tica.py:410-415
    with pytest.raises(ValueError):
        tICA(lag_time=10).fit([rs.randn(5, 3)])                     # tica.py:286-288
    with pytest.raises(ValueError):
        tICA(lag_time=1).fit([np.array([[1.0, np.nan], [0.0, 1.0], [2.0, 2.0]])])


def test_reference_identities():
    from msmbuilder_b200.decomposition import tICA
    rs = np.random.RandomState(0)
    X = rs.randn(100, 5)
    for n in range(1, 5):                                           # test_decomposition.py:59-67
        m = tICA(n_components=n, shrinkage=0, engine="simt_f64").fit([X])
        np.testing.assert_almost_equal(m.eigenvalues_.sum(), m.score([X]))
        np.testing.assert_almost_equal(m.eigenvalues_.sum(), m.score_)
    m = tICA(n_components=3, engine="simt_f64").fit([rs.randn(10, 3), rs.randn(10, 3)])
    assert m.eigenvalues_.shape == (3,) and m.eigenvectors_.shape == (3, 3)
    Y = rs.randn(401, 3).cumsum(0)                                  # test_utils.py:59-79
    a = tICA(lag_time=2, engine="simt_f64").fit([Y])
    b = tICA(lag_time=1, engine="simt_f64").fit([Y[0::2], Y[1::2]])
    np.testing.assert_allclose(a._outer_0_to_T_lagged, b._outer_0_to_T_lagged, atol=1e-8)
    k = tICA(n_components=2, lag_time=2, kinetic_mapping=True, engine="simt_f64").fit([Y])   # :101-111
    p = tICA(n_components=2, lag_time=2, engine="simt_f64").fit([Y])
    np.testing.assert_allclose(k.transform([Y])[0], p.transform([Y])[0] * p.eigenvalues_, atol=1e-9)
    c = tICA(n_components=2, lag_time=2, commute_mapping=True, engine="simt_f64").fit([Y])
    o = TicaOracle(n_components=2, lag_time=2, commute_mapping=True).fit([Y])
    np.testing.assert_allclose(np.abs(c.transform([Y])[0]), np.abs(o.transform([Y])[0]), atol=1e-8)


def test_pipeline_tica_then_kcenters():
    # docs/apipatterns.rst:103-118, test_decomposition.py:164-168
    from sklearn.pipeline import Pipeline
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters
    seqs = ar1_numpy(3, 1200, 12, seed=5)
    pipe = Pipeline([("tica", tICA(n_components=3, lag_time=4)), ("cluster", KCenters(n_clusters=5, random_state=0))])
    labels = pipe.fit_transform(seqs)
    assert [l.shape for l in labels] == [(1200,)] * 3
    # device tensors flow through transform without leaving the GPU
    import torch
    dseqs = [torch.from_numpy(s).cuda() for s in seqs]
    proj = pipe.named_steps["tica"].transform(dseqs)
    assert all(p.is_cuda and p.dtype == torch.float64 for p in proj)


def test_f32_f64_int_inputs():
    from msmbuilder_b200.decomposition import tICA
    seqs = ar1_numpy(2, 500, 5, seed=9, dtype=np.float64)
    o = TicaOracle(n_components=2, lag_time=3).fit(seqs)
    m = tICA(n_components=2, lag_time=3, engine="simt_f64").fit(seqs)
    _assert_moments_close(m, o, ACC_RTOL)
    ints = [np.round(s * 10).astype(np.int64) for s in seqs]
    oi = TicaOracle(n_components=2, lag_time=3).fit(ints)
    mi = tICA(n_components=2, lag_time=3, engine="simt_f64").fit(ints)
    _assert_moments_close(mi, oi, ACC_RTOL)


# ---- tensor-core engine (tcgen05, D = 256) vs the float64 CUDA-core engine ---------------
def _umma_vs_simt(seqs, lag, engine="umma_3xtf32"):
    from msmbuilder_b200.decomposition import tICA
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = tICA(n_components=6, lag_time=lag, engine="simt_f64").fit(seqs)
        b = tICA(n_components=6, lag_time=lag, engine=engine).fit(seqs)
    return a, b


@pytest.mark.parametrize("lag", [1, 3, 10, 37])
def test_umma_ragged_sequences_match_f64(lag):
    # lengths that are / are not multiples of 4, shorter than a tile, barely longer than lag
    lens = [4001, 130, 64, lag + 1, lag + 5, 777, 32, 2048, lag]      # the last one is skipped
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 4100, 256, seed=21), lens)]
    a, b = _umma_vs_simt(seqs, lag)
    assert a.n_observations_ == b.n_observations_ and a.n_sequences_ == b.n_sequences_
    _assert_moments_close(b, a, 5e-6)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=UMMA_EIG_ATOL)


@pytest.mark.parametrize("engine", ["auto", "umma_3xf16", "umma_3xtf32"])
def test_umma_long_run_eigenvalues_1e5(engine):
    # BASELINE config "eigenvalues vs reference within 1e-5", at a size the f64 engine finishes fast;
    # "auto" is the engine the bench and every default-constructed estimator run
    import torch
    from msmbuilder_b200.synthetic import ar1_device
    from msmbuilder_b200.decomposition import tICA
    X = ar1_device(20, 50000, 256, seed=5)
    seqs = [X[i * 50000:(i + 1) * 50000] for i in range(20)]
    a = tICA(n_components=8, lag_time=10, engine="simt_f64").fit(seqs)
    b = tICA(n_components=8, lag_time=10, engine=engine).fit(seqs)
    # raw moments: fp32 tensor-core accumulation over <= 1024-frame slabs (biased by
    # ~2^-25 per MMA step) bounds the relative error by ~3e-6
    _assert_moments_close(b, a, 5e-6)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=UMMA_EIG_ATOL)
    cos = np.abs(np.sum(a.components_ * b.components_, axis=1)) / (
        np.linalg.norm(a.components_, axis=1) * np.linalg.norm(b.components_, axis=1))
    assert cos.min() > 1 - 1e-4
    # additivity: fit(seqs) == partial_fit one sequence at a time (size-independent property)
    c = tICA(n_components=8, lag_time=10, engine=engine)
    for s in seqs:
        c.partial_fit(s)
    np.testing.assert_allclose(c.eigenvalues_, b.eigenvalues_, rtol=0, atol=2e-6)
    assert c.n_observations_ == b.n_observations_ == 1000000


def test_umma_single_pass_tf32_is_looser():
    seqs = ar1_numpy(3, 5000, 256, seed=22)
    a, b = _umma_vs_simt(seqs, 10, engine="umma_tf32")
    _assert_moments_close(b, a, 5e-3)


@pytest.mark.parametrize("engine,mom_tol", [("umma_3xbf16", 2e-5), ("umma_6xbf16", 5e-6)])
def test_umma_bf16_engines(engine, mom_tol):
    # bf16 split: 3 products ~2^-16 per element (unbiased, averages down with n), 6 products ~2^-24
    lens = [4001, 130, 64, 11, 15, 777, 32, 2048]
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 4100, 256, seed=23), lens)]
    a, b = _umma_vs_simt(seqs, 10, engine=engine)
    assert a.n_observations_ == b.n_observations_ and a.n_sequences_ == b.n_sequences_
    _assert_moments_close(b, a, mom_tol)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=UMMA_EIG_ATOL)


def test_umma_f16_engine_matches_f64():
    # fp16 h/l split of the power-of-two scaled frames: 3 products, ~2^-22 per element
    lens = [4001, 130, 64, 11, 15, 777, 32, 2048]
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 4100, 256, seed=24), lens)]
    a, b = _umma_vs_simt(seqs, 10, engine="umma_3xf16")
    assert a.n_observations_ == b.n_observations_ and a.n_sequences_ == b.n_sequences_
    _assert_moments_close(b, a, 5e-6)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=UMMA_EIG_ATOL)
    np.testing.assert_allclose(b.means_, a.means_, rtol=0, atol=1e-6)


def test_umma_f16_engine_feature_scales():
    # features spanning 1e-6 ... 1e6 (and an offset 1000x the spread): the per-feature scale
    # keeps every one of them inside fp16's range with 11-bit components
    rng = np.random.RandomState(3)
    scales = (10.0 ** rng.uniform(-6, 6, size=256)).astype(np.float32)
    offs = (scales * rng.uniform(-1000, 1000, size=256)).astype(np.float32)
    offs[::3] = 0
    seqs = [(s * scales + offs).astype(np.float32) for s in ar1_numpy(3, 3000, 256, seed=25)]
    a, b = _umma_vs_simt(seqs, 10, engine="umma_3xf16")
    # compare correlation-normalised moments: the features differ by 24 orders of magnitude
    sd = np.sqrt(np.diag(a.covariance_))
    np.testing.assert_allclose(b.covariance_ / np.outer(sd, sd), a.covariance_ / np.outer(sd, sd),
                               rtol=0, atol=2e-5)
    np.testing.assert_allclose(b.offset_correlation_ / np.outer(sd, sd),
                               a.offset_correlation_ / np.outer(sd, sd), rtol=0, atol=2e-5)
    np.testing.assert_allclose(b.means_, a.means_, rtol=1e-6, atol=0)


def test_umma_f16_engine_range_rescue():
    # an excursion 10^7 times the spread of the shift/scale sample: beyond 2^6 times the sample's
    # largest magnitude the converters raise the range flag and the call redoes itself with the
    # float64 engine on the stream (no host round trip).  The sample is frame floor(j * 8000 / 1024)
    # of the concatenated call (host side of tica_umma_accumulate), so the excursions sit on frames
    # between two sample points (global 7001..7005 and 7501).
    seqs = ar1_numpy(2, 4000, 256, seed=26)
    seqs[1] = seqs[1].copy()
    sampled = {(j * 8000) // 1024 for j in range(1024)}
    assert not (sampled & set(range(7001, 7006))) and 7501 not in sampled
    seqs[1][3001:3006, 5] += 3.0e7
    seqs[1][3501, 200] = -8.0e6
    a, b = _umma_vs_simt(seqs, 10, engine="umma_3xf16")
    assert a.n_observations_ == b.n_observations_
    # the rescue IS the float64 engine: same accumulators up to the order of its atomic adds
    for name in ("_outer_0_to_T_lagged", "_outer_0_to_TminusTau", "_outer_offset_to_T",
                 "_sum_0_to_TminusTau", "_sum_tau_to_T", "_sum_0_to_T"):
        np.testing.assert_allclose(getattr(b, name), getattr(a, name), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(b.eigenvalues_, a.eigenvalues_, rtol=0, atol=1e-9)
    # a milder outlier (100 x the sample's range) takes the same exact route
    seqs2 = ar1_numpy(2, 4000, 256, seed=28)
    seqs2[0] = seqs2[0].copy()
    seqs2[0][3, ::4] += 400.0
    a2, b2 = _umma_vs_simt(seqs2, 10, engine="umma_3xf16")
    np.testing.assert_allclose(b2.eigenvalues_, a2.eigenvalues_, rtol=0, atol=1e-9)
    # and the next call on the same estimator is clean again (flag is per call)
    d = _umma_vs_simt(seqs[:1], 10, engine="umma_3xf16")
    _assert_moments_close(d[1], d[0], 5e-6)
    assert np.abs(d[1].eigenvalues_ - d[0].eigenvalues_).max() > 0     # the tensor-core engine again


@pytest.mark.parametrize("engine", ["umma_3xf16", "umma_6xbf16"])
def test_umma_is_bit_reproducible(engine):
    # no atomics on the result path: single-writer reductions, per-pair column sums and per-slot edge
    # terms added in a fixed order -> two runs agree bit for bit
    from msmbuilder_b200.decomposition import tICA
    lens = [4001, 130, 9000, 777, 2048]
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 9000, 256, seed=27), lens)]
    a = tICA(n_components=4, lag_time=10, engine=engine).fit(seqs)
    for _ in range(3):
        b = tICA(n_components=4, lag_time=10, engine=engine).fit(seqs)
        for name in ("_outer_0_to_T_lagged", "_outer_0_to_TminusTau", "_outer_offset_to_T",
                     "_sum_0_to_TminusTau", "_sum_tau_to_T", "_sum_0_to_T"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name))


@pytest.mark.parametrize("D", [64, 96, 128, 224])
def test_umma_narrow_feature_counts(D):
    # D = 32k < 256 rides the 256-wide tensor-core tiles (TMA zero-fills the missing feature blocks)
    from msmbuilder_b200.decomposition import tICA
    lens = [3001, 257, 64, 12, 900]
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 3100, D, seed=30 + D), lens)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = tICA(n_components=4, lag_time=10, engine="simt_f64").fit(seqs)
        b = tICA(n_components=4, lag_time=10, engine="auto").fit(seqs)
        c = tICA(n_components=4, lag_time=10, engine="umma_3xtf32").fit(seqs)
    for m in (b, c):
        assert m.n_observations_ == a.n_observations_ and m.n_sequences_ == a.n_sequences_
        _assert_moments_close(m, a, 5e-6)
        np.testing.assert_allclose(m.eigenvalues_, a.eigenvalues_, rtol=0, atol=UMMA_EIG_ATOL)
        np.testing.assert_allclose(m.means_, a.means_, rtol=0, atol=1e-6)


def test_synthetic_dataset_is_independent_of_sharding():
    # bench.py generates rank r's sequences with first_seq = its offset: any N sees the same frames
    import torch
    from msmbuilder_b200.synthetic import ar1_device
    whole = ar1_device(7, 1500, 32, seed=11)
    parts = torch.cat([ar1_device(3, 1500, 32, seed=11, first_seq=0),
                       ar1_device(1, 1500, 32, seed=11, first_seq=3),
                       ar1_device(3, 1500, 32, seed=11, first_seq=4, seqs_per_chunk=2)])
    assert torch.equal(whole, parts)
    other = ar1_device(7, 1500, 32, seed=12)
    assert not torch.equal(whole, other)
    # statistics: lag-1 autocorrelation of the slowest latent direction is ~0.999
    x = whole[:1500].double()
    x = x - x.mean(0)
    c0 = (x[:-1] * x[:-1]).sum()
    c1 = (x[:-1] * x[1:]).sum()
    assert 0.5 < float(c1 / c0) < 1.0
