"""oracle/cluster_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restatement of the single-array clustering loops of the hot path on top of the
libdistance oracle.  These travel to the GPU box (the reference .py files cannot).

  kcenters_fit ............ msmbuilder/cluster/kcenters.py:79-102
      (seed draw :84, strict `d < distances_` update :93-95, first-max argmax :97)
  minibatch_kmedoids_fit .. msmbuilder/cluster/minibatchkmedoids.py:90-140
      (RandomState call order :99,:100,:108 is part of the contract)
  regular_spatial_fit ..... msmbuilder/cluster/regularspatial.py:70-77
  kmedoids_fit ............ msmbuilder/cluster/kmedoids.py:80-100
  split / split_indices ... msmbuilder/cluster/base.py:76-88

Pinned in tests/test_oracle_cluster.py against the reference's own
kcenters.py / minibatchkmedoids.py loaded verbatim (oracle/ref_loader.py) and
against tests/golden/cluster_*.npz generated from them.
"""
from operator import itemgetter

import numpy as np
from sklearn.utils import check_random_state

from . import libdistance_oracle as lo


def _as_float(X):
    if isinstance(X, np.ndarray) and X.dtype not in (np.float32, np.float64):
        X = X.astype("float64")
    return X


def kcenters_fit(X, n_clusters, metric="euclidean", random_state=None,
                 dist_fn=None, impl="port"):
    """Returns dict(cluster_ids_, cluster_centers_, labels_, distances_, inertia_)."""
    X = _as_float(X)
    if dist_fn is None:
        dist_fn = lambda X_, y_: lo.dist(X_, y_, metric, impl=impl)
    n = len(X)
    nxt = check_random_state(random_state).randint(0, n)
    labels = np.zeros(n, dtype=int)
    distances = np.full(n, np.inf, dtype=float)
    ids = []
    for i in range(n_clusters):
        d = dist_fn(X, X[nxt])
        closer = d < distances
        distances[closer] = d[closer]
        labels[closer] = i
        ids.append(int(nxt))
        nxt = np.argmax(distances)
    return dict(cluster_ids_=ids, cluster_centers_=X[ids], labels_=labels,
                distances_=distances, inertia_=float(np.sum(distances)))


def minibatch_kmedoids_fit(X, n_clusters, max_iter=5, batch_size=100,
                           metric="euclidean", max_no_improvement=10,
                           random_state=None, pdist_fn=None, assign_fn=None,
                           impl="port"):
    """Returns dict(cluster_ids_, cluster_centers_, labels_, inertia_, n_iter_)."""
    X = _as_float(X)
    if pdist_fn is None:
        pdist_fn = lambda X_, idx: lo.pdist(X_, metric, X_indices=idx, impl=impl)
    if assign_fn is None:
        assign_fn = lambda X_, Y_: lo.assign_nearest(X_, Y_, metric, impl=impl)
    n = len(X)
    n_batches = int(np.ceil(float(n) / batch_size))
    n_iter = int(max_iter * n_batches)
    rs = check_random_state(random_state)

    cluster_ids = rs.randint(0, n, size=n_clusters)
    labels = rs.randint(0, n_clusters, size=n)

    stale = 0
    done = 0
    for _ in range(n_iter):
        done += 1
        mb = np.concatenate([cluster_ids, rs.randint(0, n, batch_size)])
        dmat = pdist_fn(X, np.array(mb, dtype=np.intp))
        mb_labels = np.array(np.concatenate([np.arange(n_clusters),
                                             labels[mb[n_clusters:]]]), dtype=np.intp)
        ids, _, _ = lo.kmedoids(n_clusters, dmat, 0, mb_labels, impl=impl)
        mb_labels, mapping = lo.contigify_ids(ids, impl=impl)
        mb_cluster_ids = np.array(sorted(mapping.items(), key=itemgetter(1)))[:, 0]
        cluster_ids = mb[mb_cluster_ids]
        n_changed = np.sum(labels[mb] != mb_labels)
        if n_changed == 0:
            stale += 1
        else:
            labels[mb] = mb_labels
            stale = 0
        if stale >= max_no_improvement:
            break

    centers = X[cluster_ids]
    final_labels, inertia = assign_fn(X, centers)
    return dict(cluster_ids_=cluster_ids, cluster_centers_=centers,
                labels_=final_labels, inertia_=inertia, n_iter_=done)


def regular_spatial_fit(X, d_min, metric="euclidean", impl="port"):
    """regularspatial.py:70-77: frame i joins the centres when all its distances to
    the centres so far exceed d_min.  Returns (cluster_center_indices_, cluster_centers_)."""
    X = _as_float(X)
    ids = [0]
    for i in range(1, len(X)):
        d = lo.dist(X, X[i], metric, np.array(ids, dtype=np.intp), impl=impl)
        if np.all(d > d_min):
            ids.append(i)
    return ids, X[np.array(ids)]


def kmedoids_fit(X, n_clusters, n_passes=1, metric="euclidean", random_state=None, impl="port"):
    """kmedoids.py:80-100: full pdist, restarted k-medoids, relabel by first appearance.
    Returns dict(cluster_ids_, labels_, inertia_, cluster_centers_)."""
    X = _as_float(X)
    dmat = lo.pdist(X, metric, impl=impl)
    ids, inertia, _ = lo.kmedoids(n_clusters, dmat, n_passes, random_state=random_state, impl=impl)
    labels, mapping = lo.contigify_ids(ids, impl=impl)
    smapping = sorted(mapping.items(), key=itemgetter(1))
    cluster_ids = np.array(smapping)[:, 0]
    return dict(cluster_ids_=cluster_ids, labels_=labels, inertia_=inertia,
                cluster_centers_=X[cluster_ids])


def split(concat, lengths):
    return [concat[cl - l: cl] for (cl, l) in zip(np.cumsum(lengths), lengths)]


def split_indices(concat_inds, lengths):
    clengths = np.append([0], np.cumsum(lengths))
    mapping = np.zeros((clengths[-1], 2), dtype=int)
    for traj_i, (start, end) in enumerate(zip(clengths[:-1], clengths[1:])):
        mapping[start:end, 0] = traj_i
        mapping[start:end, 1] = np.arange(end - start)
    return mapping[concat_inds]
