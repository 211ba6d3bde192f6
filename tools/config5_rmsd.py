"""BASELINE.json configs[4]: KCenters(n_clusters=2000, metric='rmsd') on synthetic 5M frames x (100 atoms x 3),
frame-sharded over the GPUs of one box.

    python tools/config5_rmsd.py [--frames 5000000] [--atoms 100] [--k 2000] [--templates 2000] [--check-k 200]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5_rmsd.py ...

Every rank generates its shard of the same seeded data set on the device, centres it (K5) and runs the
k-centers loop of kcenters.py:79-102 with the pruned RMSD pass (K6c: triangle inequality, exact).  Prints one
JSON line: seconds, passes/s, the fraction of frame reads the pruning removed, and -- for the first
`--check-k` centres -- whether ids, labels and distances equal the plain pass-per-centre run bit for bit.
Times are CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=5_000_000)
    ap.add_argument("--atoms", type=int, default=100)
    ap.add_argument("--k", type=int, default=2000)
    ap.add_argument("--templates", type=int, default=2000)
    ap.add_argument("--noise", type=float, default=0.05)
    ap.add_argument("--check-k", type=int, default=200)
    ap.add_argument("--seed", type=int, default=5)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from msmbuilder_b200 import _kernels as K
    from msmbuilder_b200 import parallel as P
    from msmbuilder_b200.synthetic import rmsd_conformations_device

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lo, hi = P.shard_rows(a.frames, ws)[rank] if ws > 1 else (0, a.frames)
    # the same global data set at every world size: generated in fixed chunks of 250k frames, chunk c
    # seeded with seed + c, each rank keeps the rows of its shard
    chunk = 250_000
    g = torch.Generator(device="cuda")
    g.manual_seed(a.seed)
    bank = torch.randn((a.templates, a.atoms, 3), generator=g, device="cuda") * 0.3
    parts = []
    for c0 in range(0, a.frames, chunk):
        c1 = min(a.frames, c0 + chunk)
        if c1 <= lo or c0 >= hi:
            continue
        x, _ = rmsd_conformations_device(c1 - c0, a.atoms, seed=a.seed * 1000 + 1 + c0 // chunk,
                                         noise=a.noise, templates=bank)
        parts.append(x[max(lo, c0) - c0:min(hi, c1) - c0])
    xyz = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
    del parts
    traces = K.rmsd_center(xyz)
    torch.cuda.synchronize()
    seed_row = int(np.random.RandomState(0).randint(0, a.frames))

    def fit(k, prune):
        if prune:
            os.environ.pop("MSMB200_RMSD_NO_PRUNE", None)
        else:
            os.environ["MSMB200_RMSD_NO_PRUNE"] = "1"
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if ws > 1:
            ids, d, lab, ring = P.kcenters_fit_gpu(xyz, lo, k, "rmsd", seed_row, traces=traces)
        else:
            ids, d, lab = K.kcenters_fit(xyz, k, "rmsd", seed_row, traces=traces)
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if ws > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ids.cpu().numpy(), d, lab

    fit(min(a.k, 20), True)                                  # warm-up (allocator, attributes)
    ms, ids, d, lab = fit(a.k, True)
    out = {"config": "KCenters(k=%d, 'rmsd') on %d x (%d atoms x 3) f32, %d GPU(s), frame-sharded"
                     % (a.k, a.frames, a.atoms, ws),
           "seconds": ms / 1e3, "passes_per_s": a.k / (ms / 1e3), "n_gpus": ws,
           "distinct_centres": int(len(set(ids.tolist()))),
           "frame_bytes_if_every_pass_read_everything": float(a.k) * a.frames * (12 * a.atoms + 16)}
    out["equivalent_read_rate_TBps"] = out["frame_bytes_if_every_pass_read_everything"] / (ms / 1e3) / 1e12
    if a.check_k > 0:
        ck = min(a.check_k, a.k)
        ms_p, ids_p, d_p, lab_p = fit(ck, True)
        ms_u, ids_u, d_u, lab_u = fit(ck, False)
        same = bool((ids_p == ids_u).all()) and bool(torch.equal(d_p, d_u)) and bool(torch.equal(lab_p, lab_u))
        flag = torch.tensor([1 if same else 0], device="cuda")
        if ws > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["check"] = {"k": ck, "pruned_ms": ms_p, "plain_ms": ms_u,
                        "plain_ms_per_pass": ms_u / ck,
                        "plain_pass_hbm_TBps": a.frames / ws * (12 * a.atoms + 16) / (ms_u / ck / 1e3) / 1e12,
                        "ids_labels_distances_bit_identical": bool(flag.item()),
                        "ids_equal_prefix_of_full_run": bool((ids[:ck] == ids_u).all())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
