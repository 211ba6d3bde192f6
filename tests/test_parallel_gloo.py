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
class _HostLookahead(object):
    """Host stand-in for _kernels.LookaheadState (same five methods, NumPy arithmetic through the
    oracle): rows are dealt to `n_lanes` strided lanes that keep their two largest running minima,
    exactly the information the CUDA pass leaves behind."""

    def __init__(self, Xl, row_offset, n_lanes=7, t_cap=5, j_cap=4):
        from oracle import libdistance_oracle as lo
        self.lo = lo
        self.X, self.row_offset = Xl, row_offset
        self.n, self.d = Xl.shape
        self.n_lanes, self.t_cap, self.j_cap = n_lanes, t_cap, j_cap
        self.distances = torch.full((self.n,), float("inf"), dtype=torch.float64)
        self.labels = torch.zeros((self.n,), dtype=torch.int32)
        # centres blob: [n, ids[j_cap], rows[j_cap * d]] as float64 (ids are exact in a double here)
        self.centers = torch.zeros(1 + j_cap + j_cap * self.d, dtype=torch.float64)
        # set blob: [count, tau, val[t_cap], idx[t_cap], rows[t_cap * d]]
        self.set_len = 2 + 2 * t_cap + t_cap * self.d

    def centers_ids(self):
        return self.centers[1:1 + self.j_cap].to(torch.int64)

    def centers_rows(self):
        return self.centers[1 + self.j_cap:].reshape(self.j_cap, self.d).to(torch.float32)

    def seed(self, global_row):
        self.centers.zero_()
        local = global_row - self.row_offset
        if 0 <= local < self.n:
            self.centers[0] = 1
            self.centers[1] = global_row
            self.centers[1 + self.j_cap:1 + self.j_cap + self.d] = torch.from_numpy(self.X[local].astype(np.float64))

    def multi_pass(self, n_centers, label0, first):
        dist_np, lab_np = self.distances.numpy(), self.labels.numpy()
        rows = self.centers_rows().numpy()
        for j in range(n_centers):
            if self.n == 0:
                break
            dv = self.lo.dist(self.X, rows[j], "euclidean")
            m = dv < dist_np
            dist_np[m] = dv[m]
            lab_np[m] = label0 + j

    def select(self):
        out = torch.zeros(self.set_len, dtype=torch.float64)
        dist_np = self.distances.numpy()
        tops, seconds = [], [-np.inf]
        for lane in range(self.n_lanes):
            rows = np.arange(lane, self.n, self.n_lanes)
            if len(rows) == 0:
                continue
            v = dist_np[rows]
            a = int(np.argmax(v))
            tops.append((v[a], self.row_offset + int(rows[a])))
            if len(rows) > 1:
                seconds.append(np.max(np.delete(v, a)))
        if not tops:
            out[1] = -np.inf
            return out
        best = min(tops, key=lambda t: (-t[0], t[1]))
        vals = sorted((t[0] for t in tops if t[0] > 0), reverse=True)
        cut = vals[self.t_cap - 1] if len(vals) >= self.t_cap else 0.0
        cands = [t for t in tops if t[0] > 0 and t[0] > cut]
        if best not in cands:
            cands.append(best)
        out[0] = len(cands)
        out[1] = max(max(seconds), cut)
        for c, (v, gi) in enumerate(cands):
            out[2 + c] = v
            out[2 + self.t_cap + c] = gi
            o = 2 + 2 * self.t_cap + c * self.d
            out[o:o + self.d] = torch.from_numpy(self.X[gi - self.row_offset].astype(np.float64))
        return out

    def chain(self, sets, n_sets, k_remaining):
        sets = sets.reshape(n_sets, self.set_len).numpy()
        vals, idx, rows = [], [], []
        tau = -np.inf
        for s in sets:
            c = int(s[0])
            tau = max(tau, s[1])
            vals += list(s[2:2 + c])
            idx += [int(i) for i in s[2 + self.t_cap:2 + self.t_cap + c]]
            rows += [s[2 + 2 * self.t_cap + i * self.d:2 + 2 * self.t_cap + (i + 1) * self.d].astype(np.float32)
                     for i in range(c)]
        vals = np.array(vals)
        rows = np.array(rows, dtype=np.float32).reshape(len(vals), self.d)
        self.centers.zero_()
        steps = 0
        while steps < min(k_remaining, self.j_cap) and len(vals):
            order = sorted(range(len(vals)), key=lambda c: (-vals[c], idx[c]))
            b = order[0]
            if steps > 0 and not vals[b] > tau:
                break
            self.centers[1 + steps] = idx[b]
            o = 1 + self.j_cap + steps * self.d
            self.centers[o:o + self.d] = torch.from_numpy(rows[b].astype(np.float64))
            steps += 1
            vals = np.minimum(vals, self.lo.dist(rows, rows[b], "euclidean"))
        self.centers[0] = steps
        return steps


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
