"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Writes tests/golden/*.npz.

Runs the reference's OWN code (verbatim tica.py / kcenters.py /
minibatchkmedoids.py through oracle/ref_loader.py, and its libdistance C++
through oracle/_ref/libref.so) on seeded inputs and stores inputs' seeds +
outputs.  Only runnable where /root/reference exists (the build container);
the fixtures travel, the reference does not.

    python -m oracle.gen_golden
"""
import os
import warnings

import numpy as np

from . import ref_loader
from . import libdistance_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def tica_inputs(seed, n_seq, length, D, dtype):
    from msmbuilder_b200.synthetic import ar1_numpy
    seqs = ar1_numpy(n_seq, length, D, seed=seed, dtype=dtype)
    return seqs


def gen_tica():
    tICA = ref_loader.load_tica()
    cases = [
        dict(name="tica_d6_lag3", seed=11, n_seq=4, length=700, D=6, lag=3, k=3, shrinkage=None,
             dtype="float32", short=[2]),
        dict(name="tica_d16_lag10", seed=12, n_seq=3, length=1500, D=16, lag=10, k=4, shrinkage=None,
             dtype="float32", short=[]),
        dict(name="tica_d64_lag10_f64", seed=13, n_seq=2, length=2500, D=64, lag=10, k=4,
             shrinkage=0.0, dtype="float64", short=[]),
        dict(name="tica_d256_lag10", seed=14, n_seq=2, length=3000, D=256, lag=10, k=4,
             shrinkage=None, dtype="float32", short=[5]),
    ]
    for c in cases:
        seqs = tica_inputs(c["seed"], c["n_seq"], c["length"], c["D"], np.dtype(c["dtype"]))
        for n_short in c["short"]:
            seqs.insert(1, seqs[0][:n_short].copy())    # a too-short sequence: skipped, not counted
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = tICA(n_components=c["k"], lag_time=c["lag"], shrinkage=c["shrinkage"]).fit(seqs)
            proj = m.transform([seqs[0][:50]])[0]
        np.savez_compressed(
            os.path.join(GOLD, c["name"] + ".npz"),
            seed=c["seed"], n_seq=c["n_seq"], length=c["length"], D=c["D"], lag=c["lag"],
            k=c["k"], shrinkage=np.nan if c["shrinkage"] is None else c["shrinkage"],
            dtype=c["dtype"], short=np.array(c["short"], dtype=np.int64),
            eigenvalues=m.eigenvalues_, eigenvectors=m.eigenvectors_, means=m.means_,
            timescales=m.timescales_, n_observations=m.n_observations_,
            n_sequences=m.n_sequences_, shrinkage_=m.shrinkage_,
            C_tau=m._outer_0_to_T_lagged, C_00=m._outer_0_to_TminusTau,
            C_tt=m._outer_offset_to_T, S_0=m._sum_0_to_TminusTau, S_tau=m._sum_tau_to_T,
            S=m._sum_0_to_T, proj50=proj)
        print("wrote", c["name"], m.eigenvalues_)


def gen_config1():
    """BASELINE.json configs[0]: tICA(lag_time=10, n_components=4) on AlanineDipeptide dihedral
    features through the reference's CPU path.  The dataset itself needs mdtraj + a download
    (example_datasets/alanine_dipeptide.py:18-54), so the seeded stand-in of the same shape is used:
    10 trajectories x 9,999 frames x [sin phi, cos phi, sin psi, cos psi]."""
    from msmbuilder_b200.synthetic import dihedral_standin_numpy
    tICA = ref_loader.load_tica()
    seqs = dihedral_standin_numpy()
    m = tICA(n_components=4, lag_time=10).fit(seqs)
    proj = m.transform([seqs[0][:50]])[0]
    mk = tICA(n_components=2, lag_time=10, kinetic_mapping=True).fit(seqs)
    np.savez_compressed(
        os.path.join(GOLD, "tica_config1_dihedral.npz"),
        eigenvalues=m.eigenvalues_, eigenvectors=m.eigenvectors_, means=m.means_,
        timescales=m.timescales_, n_observations=m.n_observations_, n_sequences=m.n_sequences_,
        shrinkage_=m.shrinkage_, covariance=m.covariance_, offset_correlation=m.offset_correlation_,
        proj50=proj, proj50_kinetic=mk.transform([seqs[0][:50]])[0],
        score=m.score(seqs[:3]), summary=np.array(m.summarize()))
    print("wrote tica_config1_dihedral", m.eigenvalues_)


def gen_agglomerative():
    """LandmarkAgglomerative fit + predict from the reference class (cluster/agglomerative.py),
    over the compiled reference libdistance: 4 linkages x (ward predictors) x 2 metrics/dtypes."""
    LA = ref_loader.load_agglomerative()
    out = {}
    cases = []
    for linkage in ("average", "complete", "single", "ward"):
        for metric, dtype in (("euclidean", "float32"), ("cityblock", "float64")):
            preds = ("ward", "average") if linkage == "ward" else (None,)
            for wp in preds:
                cases.append((linkage, metric, dtype, wp))
    for ci, (linkage, metric, dtype, wp) in enumerate(cases):
        seqs = cluster_inputs(40 + ci, 3, 300, 6, np.dtype(dtype))
        kw = dict(n_clusters=5, n_landmarks=60, linkage=linkage, metric=metric,
                  landmark_strategy="stride" if ci % 2 == 0 else "random", random_state=7)
        if wp is not None:
            kw["ward_predictor"] = wp
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = LA(**kw).fit(seqs)
            new = cluster_inputs(90 + ci, 2, 250, 6, np.dtype(dtype))
            pred = m.predict(new)
        key = "ag%d_" % ci
        out[key + "landmark_labels"] = m.landmark_labels_
        out[key + "landmarks"] = m.landmarks_
        out[key + "centers"] = m.cluster_centers_
        out[key + "cardinality"] = m.cardinality_
        out[key + "sqsum"] = m.squared_distances_within_cluster_
        out[key + "labels"] = np.concatenate(m.labels_) if hasattr(m, "labels_") else np.zeros(0)
        out[key + "pred"] = np.concatenate(pred)
        # margin between the best and second-best pooled distance of every predicted frame,
        # so the parity test can exempt genuine near-ties
        out[key + "kw"] = np.array(repr(sorted(kw.items())))
    out["n_cases"] = len(cases)
    out["cases"] = np.array([repr(c) for c in cases])
    np.savez_compressed(os.path.join(GOLD, "agglomerative.npz"), **out)
    print("wrote agglomerative", len(cases), "cases")


def cluster_inputs(seed, n_seq, length, D, dtype):
    rs = np.random.RandomState(seed)
    centers = rs.randn(7, D) * 3
    return [(centers[rs.randint(0, 7, size=length)] + rs.randn(length, D)).astype(dtype)
            for _ in range(n_seq)]


def gen_cluster():
    KCenters, MiniBatchKMedoids, _ = ref_loader.load_cluster()
    out = {}
    for metric in lo.VECTOR_METRICS:
        for dtype in ("float32", "float64"):
            seqs = cluster_inputs(21, 3, 400, 5, np.dtype(dtype))
            if metric in ("hamming", "jaccard"):
                seqs = [np.round(s).astype(dtype) for s in seqs]
            kc = KCenters(n_clusters=9, metric=metric, random_state=3).fit(seqs)
            key = "kc_%s_%s_" % (metric, dtype)
            out[key + "ids"] = np.array(kc.cluster_ids_)
            out[key + "labels"] = np.concatenate(kc.labels_)
            out[key + "distances"] = np.concatenate(kc.distances_)
            out[key + "inertia"] = kc.inertia_
            out[key + "predict"] = np.concatenate(kc.predict(seqs))
            mb = MiniBatchKMedoids(n_clusters=6, batch_size=40, max_iter=3, metric=metric,
                                   random_state=5).fit(seqs)
            key = "mb_%s_%s_" % (metric, dtype)
            out[key + "ids"] = np.array(mb.cluster_ids_)
            out[key + "labels"] = np.concatenate(mb.labels_)
            out[key + "inertia"] = mb.inertia_
    np.savez_compressed(os.path.join(GOLD, "cluster_small.npz"), **out)
    print("wrote cluster_small with", len(out), "arrays")


def gen_cluster_more():
    """RegularSpatial / KMedoids from the reference classes, verbatim."""
    RegularSpatial, KMedoids = ref_loader.load_more_clusterers()
    out = {}
    for metric, d_min in (("euclidean", 4.0), ("cityblock", 8.0), ("chebyshev", 2.5)):
        for dtype in ("float32", "float64"):
            seqs = cluster_inputs(31, 3, 300, 5, np.dtype(dtype))
            rs = RegularSpatial(d_min=d_min, metric=metric).fit(seqs)
            key = "rs_%s_%s_" % (metric, dtype)
            out[key + "d_min"] = d_min
            out[key + "ids"] = np.asarray(rs.cluster_center_indices_)
            out[key + "centers"] = rs.cluster_centers_
            out[key + "predict"] = np.concatenate(rs.predict(seqs))
            for n_passes in (1, 4):
                km = KMedoids(n_clusters=5, n_passes=n_passes, metric=metric, random_state=7).fit(seqs)
                key = "km%d_%s_%s_" % (n_passes, metric, dtype)
                out[key + "ids"] = np.asarray(km.cluster_ids_)
                out[key + "labels"] = np.concatenate(km.labels_)
                out[key + "inertia"] = km.inertia_
    np.savez_compressed(os.path.join(GOLD, "cluster_more.npz"), **out)
    print("wrote cluster_more with", len(out), "arrays")


def msm_inputs(case):
    rs = np.random.RandomState(100 + case)
    n_states = (4, 9, 70, 200)[case]
    seqs = []
    for _ in range(5):
        n = int(rs.randint(1, 3000))
        y = rs.randint(0, n_states, size=n)
        if case == 1:
            y = y * 5 - 7                       # non-contiguous, negative labels
        seqs.append(y.astype(np.int64))
    return seqs


def gen_msm():
    """_transition_counts (msm/core.py:487-602) from the reference function, verbatim."""
    tc = ref_loader.load_transition_counts()
    out = {"n_cases": 4}
    for case, (lag, sliding) in enumerate([(1, True), (3, True), (4, False), (7, True)]):
        seqs = msm_inputs(case)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c, m = tc(seqs, lag_time=lag, sliding_window=sliding)
        out["lens_%d" % case] = np.array([len(s) for s in seqs])
        out["labels_%d" % case] = np.concatenate(seqs)
        out["lag_%d" % case] = lag
        out["sliding_%d" % case] = sliding
        out["counts_%d" % case] = c
        out["classes_%d" % case] = np.array(sorted(int(k) for k in m.keys()))
    np.savez_compressed(os.path.join(GOLD, "msm_counts.npz"), **out)
    print("wrote msm_counts")


def main():
    import sys
    if not ref_loader.available():
        raise SystemExit("reference tree absent: goldens can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["tica", "cluster", "cluster_more", "msm", "config1", "agglomerative"]
    for name in which:
        {"tica": gen_tica, "cluster": gen_cluster, "cluster_more": gen_cluster_more,
         "msm": gen_msm, "config1": gen_config1, "agglomerative": gen_agglomerative}[name]()


if __name__ == "__main__":
    main()
