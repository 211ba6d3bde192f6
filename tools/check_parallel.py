"""tools/check_parallel.py -- run under torchrun with >= 2 GPUs:
sharded tICA / KCenters (NCCL all-reduce / all-gather) == single-GPU estimators."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmbuilder_b200 import parallel as par
from msmbuilder_b200.cluster import KCenters
from msmbuilder_b200.decomposition import tICA
from msmbuilder_b200.synthetic import ar1_numpy


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    lens = [3000, 1700, 2600, 900, 2048, 1501, 777, 5]
    seqs = [s[:n] for s, n in zip(ar1_numpy(len(lens), 3000, 256, seed=7), lens)]

    # --- tICA: whole sequences dealt to ranks, one all-reduce
    owned = par.shard_sequences(lens, ws)[rank]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for engine in ("simt_f64", "auto"):
            t = par.tica_fit_sharded(tICA(n_components=4, lag_time=10, engine=engine), [seqs[i] for i in owned])
            ref = tICA(n_components=4, lag_time=10, engine=engine).fit(seqs)
            assert t.n_observations_ == ref.n_observations_ and t.n_sequences_ == ref.n_sequences_
            tol = 1e-12 if engine == "simt_f64" else 2e-6
            np.testing.assert_allclose(t._outer_0_to_T_lagged, ref._outer_0_to_T_lagged,
                                       rtol=0, atol=tol * np.abs(ref._outer_0_to_T_lagged).max())
            np.testing.assert_allclose(t.eigenvalues_, ref.eigenvalues_, rtol=0, atol=1e-9 if engine == "simt_f64" else 5e-6)

    # --- KCenters: contiguous frame shards, per-pass candidate all-gather
    X = np.concatenate(seqs[:7])
    n_total = len(X)
    bounds = par.shard_rows(n_total, ws)
    a, b = bounds[rank]
    for metric in ("euclidean", "cityblock"):
        kc = par.kcenters_fit_sharded(KCenters(n_clusters=13, metric=metric, random_state=3), [X[a:b]], a, n_total)
        ref = KCenters(n_clusters=13, metric=metric, random_state=3).fit([X])
        assert kc.cluster_ids_ == ref.cluster_ids_, (kc.cluster_ids_, ref.cluster_ids_)
        np.testing.assert_array_equal(kc.labels_[0], ref.labels_[0][a:b])
        np.testing.assert_array_equal(kc.distances_[0], ref.distances_[0][a:b])
        np.testing.assert_array_equal(kc.cluster_centers_, ref.cluster_centers_)
        assert abs(kc.inertia_ - ref.inertia_) <= 1e-10 * ref.inertia_
    # --- look-ahead KCenters across ranks (candidate sets all-gathered once per chain) against the
    #     one-pass-per-centre schedule on one GPU, on data where the chain certifies little
    from msmbuilder_b200 import _kernels as K
    rs = np.random.RandomState(11)
    cen = rs.randn(5, 64) * 8
    Y = (cen[rs.randint(0, 5, 30000)] + 0.1 * rs.randn(30000, 64)).astype(np.float32)
    (a, b) = par.shard_rows(len(Y), ws)[rank]
    stats = {}
    ids, distances, labels, ring = par.kcenters_fit_gpu(torch.from_numpy(Y[a:b]).cuda(), a, 40, "euclidean",
                                                        seed_global=777, stats=stats)
    rid, rdist, rlab = K.kcenters_fit(torch.from_numpy(Y).cuda(), 40, "euclidean", 777, lookahead=False)
    assert torch.equal(ids, rid), (ids, rid)
    assert torch.equal(labels, rlab[a:b]) and torch.equal(distances, rdist[a:b])
    assert stats["passes"] <= 40
    dist.barrier()
    if rank == 0:
        print("PARALLEL_OK world_size=%d (look-ahead passes for k=40: %d)" % (ws, stats["passes"]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
