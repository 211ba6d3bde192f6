"""First contact of the second-generation tcgen05 tICA kernel with the hardware: small shapes against
the float64 engine, one process per shape so a hang costs one timeout and nothing else.

    python tools/v2_smoke.py D n_seq seq_len [lag]
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    D, n_seq, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    lag = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    import torch
    from msmbuilder_b200.synthetic import ar1_device
    from msmbuilder_b200.decomposition import tICA
    X = ar1_device(n_seq, L, D, seed=7)
    lens = [L - (37 * i) % 61 for i in range(n_seq)]           # ragged
    seqs = [X[i * L:i * L + n] for i, n in enumerate(lens)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = tICA(n_components=4, lag_time=lag, engine="simt_f64").fit(seqs)
        est = tICA(n_components=4, lag_time=lag, engine="umma_3xf16").fit(seqs)
    torch.cuda.synchronize()
    sd = np.sqrt(np.abs(np.diag(ref._outer_0_to_TminusTau)))
    out = []
    for n in ("_outer_0_to_T_lagged", "_outer_0_to_TminusTau", "_outer_offset_to_T"):
        out.append(float((np.abs(getattr(est, n) - getattr(ref, n)) / np.outer(sd, sd)).max()))
    sums = max(float(np.abs(getattr(est, n) - getattr(ref, n)).max() / (np.abs(getattr(ref, n)).max() + 1e-30))
               for n in ("_sum_0_to_TminusTau", "_sum_tau_to_T", "_sum_0_to_T"))
    eig = float(np.abs(est.eigenvalues_ - ref.eigenvalues_).max())
    ok = max(out) < 2e-5 and eig < 1e-5 and sums < 1e-5 and est.n_observations_ == ref.n_observations_
    print("D=%d n_seq=%d L=%d lag=%d: moment err/var %s  sums %.2e  eig %.2e  %s"
          % (D, n_seq, L, lag, ["%.2e" % o for o in out], sums, eig, "OK" if ok else "MISMATCH"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
