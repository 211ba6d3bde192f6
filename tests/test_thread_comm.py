"""CPU: the in-process communicator behind `devices=` (parallel.ThreadComm) -- the two collectives of
the k-centers loops between threads, on host tensors; and the whole rank-collective Gonzalez loop
(parallel.kcenters_fit_distributed) over it against the single-rank oracle."""
import threading

import numpy as np
import torch

from msmbuilder_b200 import parallel as P


def _run(n, fn):
    group = P.ThreadGroup(n)
    out, errs = [None] * n, []

    def work(r):
        try:
            out[r] = fn(group.comm(r), r)
        except BaseException as e:      # noqa: B902
            errs.append(e)
            group.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    return out


def test_all_gather_and_all_reduce_between_threads():
    def fn(comm, r):
        res = []
        for it in range(20):               # repeated rounds: slots are reused, barriers must hold
            local = torch.full((5,), float(10 * it + r))
            out = torch.empty(comm.ws * 5)
            comm.all_gather_into(out, local)
            t = torch.arange(4, dtype=torch.float64) * (r + 1) + it
            comm.all_reduce_sum(t)
            res.append((out.clone(), t.clone()))
        return res

    n = 3
    outs = _run(n, fn)
    for r in range(n):
        for it, (g, t) in enumerate(outs[r]):
            want = torch.cat([torch.full((5,), float(10 * it + q)) for q in range(n)])
            assert torch.equal(g, want)
            want_t = sum(torch.arange(4, dtype=torch.float64) * (q + 1) + it for q in range(n))
            assert torch.equal(t, want_t)


def test_gonzalez_loop_over_threads_matches_single_rank():
    # host mirror of the candidate protocol (the same one tests/test_parallel_gloo.py drives over gloo)
    rs = np.random.RandomState(0)
    X = rs.randn(600, 5)
    k, seed = 7, 123
    cand_bytes = P.CAND_HEADER + 8 * 5

    def reference():
        d = np.full(len(X), np.inf)
        ids, nxt = [], seed
        for _ in range(k):
            dd = np.sqrt(((X - X[nxt]) ** 2).sum(1))
            d = np.minimum(d, dd)
            ids.append(int(nxt))
            nxt = int(np.argmax(d))
        return ids

    def fn(comm, r):
        lo, hi = P.shard_rows(len(X), comm.ws)[r]
        Xl = X[lo:hi]
        d = np.full(len(Xl), np.inf)

        def write(cand, value, index, row):
            cand[:8].view(torch.float64)[0] = value
            cand[8:16].view(torch.int64)[0] = index
            cand[16:].view(torch.float64)[:] = torch.from_numpy(np.ascontiguousarray(row))

        def seed_fn(cand):
            if lo <= seed < hi:
                write(cand, float("inf"), seed, X[seed])
            else:
                cand.zero_()
                cand[:8].view(torch.float64)[0] = float("-inf")

        def pass_fn(center, label, out):
            c = center[16:].view(torch.float64).numpy()
            dd = np.sqrt(((Xl - c) ** 2).sum(1))
            np.minimum(d, dd, out=d)
            j = int(np.argmax(d))
            write(out, float(d[j]), lo + j, Xl[j])

        def select_fn(gathered, n, out):
            w = P.select_candidate_host(gathered, n, cand_bytes)
            out.copy_(gathered[w * cand_bytes:(w + 1) * cand_bytes])

        ring = P.kcenters_fit_distributed(k, cand_bytes, seed_fn, pass_fn, select_fn,
                                          lambda nb: torch.zeros(int(nb), dtype=torch.uint8), comm=comm)
        return [int(v) for v in ring[:k, 8:16].contiguous().view(torch.int64).reshape(k)]

    ref = reference()
    for ids in _run(4, fn):
        assert ids == ref
