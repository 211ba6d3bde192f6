"""CPU, world_size 2, gloo: the multi-GPU protocol of msmbuilder_b200.parallel
(sequence sharding + one all-reduce for tICA; per-pass candidate all-gather +
deterministic select for KCenters) reproduces the single-process oracle.  The
per-rank compute callbacks are host stand-ins (the oracle); on the GPU box the
same driver code runs with the CUDA kernels (tests/test_gpu_parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msmbuilder_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        ret[rank] = fn(rank, ws)
    finally:
        dist.destroy_process_group()


def _run(fn, ws=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), fn, ret), nprocs=ws, join=True)
    return [ret[r] for r in range(ws)]


def test_shard_helpers():
    lens = [100, 5, 70, 70, 30, 1]
    owned = par.shard_sequences(lens, 3)
    assert sorted(sum(owned, [])) == list(range(6))
    loads = [sum(lens[i] for i in o) for o in owned]
    assert max(loads) - min(loads) <= 30
    assert par.shard_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert par.shard_rows(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def _tica_rank(rank, ws):
    from oracle.tica_oracle import TicaOracle
    from msmbuilder_b200.synthetic import ar1_numpy
    seqs = ar1_numpy(5, 400, 6, seed=1, dtype=np.float64) + [np.zeros((2, 6))]
    lens = [len(s) for s in seqs]
    mine = par.shard_sequences(lens, ws)[rank]
    o = TicaOracle(n_components=3, lag_time=7)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o._initialize(6)
        for i in mine:
            o.partial_fit(seqs[i])
    packed = torch.from_numpy(o.packed_moments().copy())
    par.allreduce_packed(packed)
    return packed.numpy()


def test_tica_allreduce_equals_single_process():
    from oracle.tica_oracle import TicaOracle
    from msmbuilder_b200.synthetic import ar1_numpy
    from msmbuilder_b200.decomposition import tICA
    import warnings
    res = _run(_tica_rank)
    np.testing.assert_array_equal(res[0], res[1])
    seqs = ar1_numpy(5, 400, 6, seed=1, dtype=np.float64) + [np.zeros((2, 6))]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = TicaOracle(n_components=3, lag_time=7).fit(seqs)
    np.testing.assert_allclose(res[0], ref.packed_moments(), rtol=1e-12)
    m = tICA(n_components=3, lag_time=7)
    m._initialize(6)
    m._add_packed(res[0])
    assert m.n_sequences_ == 5 and m.n_observations_ == 2000
    np.testing.assert_allclose(m.eigenvalues_, ref.eigenvalues_, atol=1e-12)


def _kc_rank(rank, ws):
    from oracle import libdistance_oracle as lo
    rs = np.random.RandomState(4)
    X = rs.randn(501, 3).astype(np.float32)
    X[77] = X[400]                         # a duplicate frame: tie across ranks
    k, seed = 12, 333
    start, stop = par.shard_rows(len(X), ws)[rank]
    Xl = X[start:stop]
    cand_bytes = 16 + 4 * 3 + 4            # padded to 16
    cand_bytes = (cand_bytes + 15) // 16 * 16
    distances = np.full(len(Xl), np.inf)
    labels = np.zeros(len(Xl), dtype=np.int64)

    def write(cand, value, gidx, row):
        b = cand.numpy()
        b[:8] = np.array([value], dtype=np.float64).view(np.uint8)
        b[8:16] = np.array([gidx], dtype=np.int64).view(np.uint8)
        b[16:28] = np.asarray(row, dtype=np.float32).view(np.uint8)

    def seed_fn(cand):
        if start <= seed < stop:
            write(cand, np.inf, seed, X[seed])
        else:
            write(cand, -np.inf, 0, np.zeros(3))

    def pass_fn(center_cand, label, out):
        c = center_cand.numpy()[16:28].copy().view(np.float32)
        d = lo.dist(Xl, c, "euclidean")
        m = d < distances
        distances[m] = d[m]
        labels[m] = label
        if len(Xl):
            a = int(np.argmax(distances))
            write(out, distances[a], start + a, Xl[a])
        else:
            write(out, -np.inf, 0, np.zeros(3))

    def select_fn(gathered, n, dst):
        w = par.select_candidate_host(gathered, n, cand_bytes)
        dst.copy_(gathered.reshape(n, cand_bytes)[w])

    ring = par.kcenters_fit_distributed(k, cand_bytes, seed_fn, pass_fn, select_fn,
                                        lambda nb: torch.zeros(int(nb), dtype=torch.uint8))
    ids = ring[:k, 8:16].contiguous().numpy().view(np.int64).reshape(-1).copy()
    return ids, labels, distances, (start, stop)


def test_kcenters_gather_select_equals_single_process():
    from oracle import cluster_oracle as co
    res = _run(_kc_rank)
    rs = np.random.RandomState(4)
    X = rs.randn(501, 3).astype(np.float32)
    X[77] = X[400]

    class FixedSeed(object):          # kcenters.py:84 draws randint(0, n): pin it to 333
        def randint(self, lo_, hi):
            return 333
    import sklearn.utils
    ref = None
    orig = co.check_random_state
    co.check_random_state = lambda rs_: FixedSeed()
    try:
        ref = co.kcenters_fit(X, 12, "euclidean")
    finally:
        co.check_random_state = orig
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][0], ref["cluster_ids_"])
    labels = np.concatenate([res[0][1], res[1][1]])
    dists = np.concatenate([res[0][2], res[1][2]])
    np.testing.assert_array_equal(labels, ref["labels_"])
    np.testing.assert_array_equal(dists, ref["distances_"])


# ------------------------------------------------------------------ look-ahead k-centers protocol
from _lookahead_host import HostLookahead as _HostLookahead   # host stand-in for the five kernel calls


def _kc_lookahead_rank(rank, ws):
    from msmbuilder_b200 import _kernels as K
    rs = np.random.RandomState(4)
    X = rs.randn(501, 3).astype(np.float32)
    X[77] = X[400]
    k, seed = 12, 333
    start, stop = par.shard_rows(len(X), ws)[rank]
    st = _HostLookahead(X[start:stop], start)
    gather_sets, bcast = par.lookahead_collectives()
    stats = {}
    ids, rows, distances, labels = K.kcenters_fit_lookahead(None, k, "euclidean", seed, gather_sets=gather_sets,
                                                            bcast=bcast, row_offset=start, stats=stats, state=st)
    return ids.numpy(), labels.numpy(), distances.numpy(), rows.numpy(), stats["passes"]


def test_kcenters_lookahead_protocol_equals_single_process():
    """all-gather of candidate sets + identical chain on every rank == the reference loop"""
    from oracle import cluster_oracle as co
    res = _run(_kc_lookahead_rank)
    rs = np.random.RandomState(4)
    X = rs.randn(501, 3).astype(np.float32)
    X[77] = X[400]

    class FixedSeed(object):
        def randint(self, lo_, hi):
            return 333
    orig = co.check_random_state
    co.check_random_state = lambda rs_: FixedSeed()
    try:
        ref = co.kcenters_fit(X, 12, "euclidean")
    finally:
        co.check_random_state = orig
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][0], ref["cluster_ids_"])
    np.testing.assert_array_equal(np.concatenate([res[0][1], res[1][1]]), ref["labels_"])
    np.testing.assert_array_equal(np.concatenate([res[0][2], res[1][2]]), ref["distances_"])
    np.testing.assert_array_equal(res[0][3], X[ref["cluster_ids_"]])
    assert res[0][4] == res[1][4] <= 12          # same number of passes on both ranks, never more than k
