"""CPU: the C-ABI shared library loads without a GPU and exports every symbol that
include/msmb200.h declares; the host-only entry points (k-medoids) are checked
against the oracle here because they need no device."""
import ctypes
import os
import re

import numpy as np
import pytest

from msmbuilder_b200 import _lib
from oracle import libdistance_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "msmb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msmb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libmsmb200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "python binding table lacks %s" % s
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_sizes():
    lib = _lib.load()
    assert lib.msmb200_abi_version() == 1
    assert lib.msmb200_tica_acc_len(256) == 3 * 256 * 256 + 3 * 256 + 2
    assert lib.msmb200_candidate_bytes(256, _lib.F32) == 16 + 1024
    assert lib.msmb200_candidate_bytes(3, _lib.F64) % 16 == 0


def test_lookahead_host_helpers():
    # shape gate and blob sizes of the look-ahead k-centers (no device needed)
    lib = _lib.load()
    eu, sq, cb = (_lib.VECTOR_METRICS.index(m) for m in ("euclidean", "sqeuclidean", "cityblock"))
    for d, ok in ((256, 1), (128, 1), (64, 1), (16, 1), (512, 1), (5, 0), (12, 0), (8, 0), (1024, 0)):
        assert lib.msmb200_kcenters_lookahead_supported(d, d, _lib.F32, eu) == ok, d
    assert lib.msmb200_kcenters_lookahead_supported(256, 256, _lib.F32, sq) == 1
    assert lib.msmb200_kcenters_lookahead_supported(256, 256, _lib.F32, cb) == 0      # other metrics: one pass per centre
    assert lib.msmb200_kcenters_lookahead_supported(256, 256, _lib.F64, eu) == 0
    assert lib.msmb200_kcenters_lookahead_supported(256, 258, _lib.F32, eu) == 0      # rows not 16-byte aligned
    assert lib.msmb200_kcenters_set_bytes(256, 512) == 32 + 512 * 16 + 512 * 256 * 4
    assert lib.msmb200_kcenters_centers_bytes(256, 16) == 32 + 16 * 8 + 16 * 256 * 4
    assert lib.msmb200_kcenters_lane_bytes(0) >= 32 + 48 * 148 * 9 * 256


def test_no_cpu_fallback_is_loud():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters
    with pytest.raises(_lib.Msmb200Error):
        tICA(lag_time=1).fit([np.random.randn(50, 3)])
    with pytest.raises(_lib.Msmb200Error):
        KCenters(n_clusters=2).fit([np.random.randn(50, 3)])


def test_host_kmedoids_matches_oracle():
    from msmbuilder_b200 import _kernels as K
    rs = np.random.RandomState(0)
    for trial in range(30):
        n, k = rs.randint(6, 80), rs.randint(2, 8)
        X = rs.randn(n, 3)
        if trial % 4 == 0:
            X = np.round(X)
        dm = lo.pdist(X, "euclidean")
        cid = rs.randint(0, k, n)
        cid[:k] = np.arange(k)
        a = K.kmedoids(k, dm, 0, cid)
        b = lo.kmedoids(k, dm, 0, cid)
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1] == b[1] and a[2] == b[2]
        ca, cb = K.contigify_ids(a[0].copy()), lo.contigify_ids(b[0].copy())
        np.testing.assert_array_equal(ca[0], cb[0])
        assert ca[1] == cb[1]
    with pytest.raises(ValueError):
        K.kmedoids(3, np.zeros(4), 0, None)          # not a triangular number
    with pytest.raises(ValueError):
        K.kmedoids(9, np.zeros(3), 0, None)          # more clusters than elements


def test_host_kmedoids_restarts_match_oracle():
    # n_pass >= 1 (cluster/kmedoids.py:92-94): same RandomState draws, same kept solution
    from msmbuilder_b200 import _kernels as K
    rs = np.random.RandomState(1)
    for trial in range(20):
        n, k = int(rs.randint(6, 90)), int(rs.randint(1, 8))
        n_pass = int(rs.randint(1, 6))
        X = rs.randn(n, 3)
        if trial % 5 == 0:
            X = np.round(X)                # ties
        dm = lo.pdist(X, "euclidean")
        a = lo.kmedoids(k, dm, n_pass, random_state=trial)
        b = K.kmedoids(k, dm, n_pass, random_state=trial)
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1] == b[1] and a[2] == b[2], (trial, a[1:], b[1:])
    with pytest.raises(ValueError):
        K.kmedoids(3, np.zeros(10), -1)


def test_assign_engine_choice():
    # which filter float32 (sq)euclidean assign_nearest takes (include/msmb200.h): resident tcgen05
    # while the centres' fp16 tiles fit 96 KB, streamed chunks above, SIMT for small n / odd widths
    lib = _lib.load()
    assert lib.msmb200_assign_engine(10_000_000, 500, 16) == 0      # config 3
    assert lib.msmb200_assign_engine(10_000_000, 2000, 128) == 1    # the k = 2000 x D = 128 point
    assert lib.msmb200_assign_engine(50_000_000, 8, 256) == 1       # KCenters.predict at the bench width
    assert lib.msmb200_assign_engine(10_000_000, 2000, 16) == 1
    assert lib.msmb200_assign_engine(10_000_000, 256, 64) == 0
    assert lib.msmb200_assign_engine(1000, 500, 16) == 2            # latency bound: SIMT
    assert lib.msmb200_assign_engine(10_000_000, 500, 18) == 2      # d % 4 != 0
    assert lib.msmb200_assign_engine(10_000_000, 500, 300) == 2     # d > 256
    assert lib.msmb200_assign_engine(10_000_000, 12000, 128) == 2   # |c|^2 table beyond the shared-memory budget
