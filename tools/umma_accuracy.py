"""tools/umma_accuracy.py -- accuracy + speed of the tcgen05 tICA engine vs the f64
engine on device-generated data; prints one line per (passes, slab) setting."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmbuilder_b200.decomposition import tICA
from msmbuilder_b200.synthetic import ar1_device


def packed(m):
    return np.concatenate([m._outer_0_to_T_lagged.ravel(), m._outer_0_to_TminusTau.ravel(),
                           m._outer_offset_to_T.ravel()])


def main():
    n_seq, L, D = int(os.environ.get("NSEQ", 40)), 100000, 256
    X = ar1_device(n_seq, L, D, seed=3)
    seqs = [X[i * L:(i + 1) * L] for i in range(n_seq)]
    torch.cuda.synchronize()
    t0 = time.time()
    ref = tICA(n_components=8, lag_time=10, engine="simt_f64").fit(seqs)
    torch.cuda.synchronize()
    print("simt_f64: %.1f ms for %d frames; eig %s" % (1e3 * (time.time() - t0), n_seq * L, ref.eigenvalues_[:4]))
    pr = packed(ref)
    engines = os.environ.get("ENGINES", "umma_3xtf32,umma_tf32,umma_3xbf16,umma_6xbf16").split(",")
    slabs = [int(v) for v in os.environ.get("SLABS", "16,32,64,128,100000").split(",")]
    for engine in engines:
        for slab in slabs:
            os.environ["MSMB200_UMMA_SLAB_TILES"] = str(slab)
            m = tICA(n_components=8, lag_time=10, engine=engine)
            m.fit(seqs)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            m2 = tICA(n_components=8, lag_time=10, engine=engine)
            m2._initialize(D)
            ev[0].record()
            m2._accumulate_device(seqs)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            pm = packed(m)
            rel = np.abs(pm - pr).max() / np.abs(pr).max()
            cov_err = np.abs(m.covariance_ - ref.covariance_).max() / np.abs(ref.covariance_).max()
            oc_err = np.abs(m.offset_correlation_ - ref.offset_correlation_).max() / np.abs(ref.offset_correlation_).max()
            de = np.abs(m.eigenvalues_ - ref.eigenvalues_).max()
            print("%s slab_tiles=%-6d: %.2f ms (%.1f Mframes/s, %.0f TFLOP/s algorithmic) | raw-moment rel err %.2e | "
                  "cov rel err %.2e | offset-corr rel err %.2e | max |d eig| %.2e"
                  % (engine, slab, ms, n_seq * L / ms / 1e3, 4 * D * D * n_seq * L / ms / 1e9, rel, cov_err, oc_err, de))
    os.environ.pop("MSMB200_UMMA_SLAB_TILES", None)


if __name__ == "__main__":
    main()
