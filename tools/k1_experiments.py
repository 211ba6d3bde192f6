"""K1 experiments on the GPU box: each environment-knob variant of the tcgen05 tICA kernel is
checked against the float64 engine and timed (CUDA events) on the same frames.

    python tools/k1_experiments.py [--frames 8000000] [--features 256] VARIANT ...

VARIANT = name:ENV=VAL,ENV=VAL   (e.g.  red2:MSMB200_UMMA_FLUSH_RED=2)
"""
import argparse
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8_000_000)
    ap.add_argument("--features", type=int, default=256)
    ap.add_argument("--seq-len", type=int, default=100_000)
    ap.add_argument("--lag", type=int, default=10)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("variants", nargs="*")
    a = ap.parse_args()
    import torch
    from msmbuilder_b200.synthetic import ar1_device
    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200 import _lib
    L, D = a.seq_len, a.features
    n_seq = a.frames // L
    X = ar1_device(n_seq, L, D, seed=1000)
    seqs = [X[i * L:(i + 1) * L] for i in range(n_seq)]
    lib = _lib.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = tICA(n_components=4, lag_time=a.lag, engine="simt_f64").fit(seqs)
    print("frames %d x %d, f64 eigenvalues %s" % (n_seq * L, D, ref.eigenvalues_), flush=True)
    variants = a.variants or ["default:"]
    for v in variants:
        name, _, envs = v.partition(":")
        sets = dict(kv.split("=") for kv in envs.split(",") if kv)
        old = {k: os.environ.get(k) for k in sets}
        os.environ.update(sets)
        try:
            est = tICA(n_components=4, lag_time=a.lag, engine=a.engine)
            est._initialize(D)
            acc = torch.zeros(int(lib.msmb200_tica_acc_len(D)), dtype=torch.float64, device="cuda")
            for _ in range(2):
                acc.zero_()
                est._accumulate_device(seqs, acc=acc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(a.reps):
                acc.zero_()
                e0.record()
                est._accumulate_device(seqs, acc=acc)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            est._add_packed(acc.cpu().numpy())
            err = float(np.abs(est.eigenvalues_ - ref.eigenvalues_).max())
            DD = D * D
            b = ref._outer_0_to_TminusTau
            sd = np.sqrt(np.abs(np.diag(b)))
            merr = max(float((np.abs(getattr(est, n) - getattr(ref, n)) / np.outer(sd, sd)).max())
                       for n in ("_outer_0_to_T_lagged", "_outer_0_to_TminusTau", "_outer_offset_to_T"))
            print("%-28s min %.3f ms  median %.3f ms  (%.1f Mframes/s, %.0f TFLOP/s alg)  eig_err %.2e  "
                  "moment_err/var %.2e" % (name, min(ts), float(np.median(ts)),
                                           n_seq * L / min(ts) / 1e3, 4.0 * D * D * n_seq * L / min(ts) / 1e9,
                                           err, merr / (n_seq * L)), flush=True)
        except Exception as e:  # keep going: a failing variant must not hide the others
            print("%-28s FAILED: %s" % (name, e), flush=True)
            torch.cuda.synchronize()
        finally:
            for k, val in old.items():
                if val is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = val


if __name__ == "__main__":
    main()
