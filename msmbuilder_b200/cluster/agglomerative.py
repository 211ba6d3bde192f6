"""LandmarkAgglomerative on the device distance kernels.

Mirror of ``msmbuilder.cluster.LandmarkAgglomerative`` (msmbuilder/cluster/agglomerative.py:76-300):
cluster ``n_landmarks`` landmark frames hierarchically, then give every frame to the cluster that
minimises the linkage function between the frame and that cluster's landmarks.

Where the time goes in the reference is ``predict`` (agglomerative.py:234-269): a host
``libdistance.cdist`` of ALL frames against the landmarks -- an (n_frames, n_landmarks) float64
matrix -- followed by one NumPy pooling pass per cluster.  Here frames stay on the device: the
distances come chunk by chunk from ``msmb200_cdist`` / the RMSD kernel (same arithmetic as the
reference's libdistance) and ``msmb200_pooled_assign`` pools and arg-mins each chunk in place; only
the int labels leave the GPU.  ``fit`` keeps the reference's structure: landmark ``pdist`` on the
device (K4), the linkage tree on the host (SciPy's ``linkage`` -- the reference imports
``fastcluster.linkage``, a drop-in for it that returns the same stepwise dendrogram), and the
within-cluster squared-distance sums for the ward predictor.
"""
from __future__ import absolute_import, print_function, division

import warnings

import numpy as np
from scipy.cluster.hierarchy import fcluster, linkage
from sklearn.base import ClusterMixin, TransformerMixin
from sklearn.utils import check_random_state

from .base import MultiSequenceClusterMixin
from .kcenters import _prepare
from ..base import BaseEstimator
from .. import _lib

__all__ = ['LandmarkAgglomerative']

_POOLS = {'average': 0, 'complete': 1, 'single': 2, 'ward': 3}
_CHUNK_BYTES = 256 << 20          # of (chunk, n_landmarks) float64 distances on the device


def _to_host_int(labels):
    import torch
    from .. import _device as dev
    return dev.to_host(labels, torch.int64).astype(int, copy=False)


class _LandmarkAgglomerative(ClusterMixin, TransformerMixin):
    """Landmark-based agglomerative hierarchical clustering.

    Parameters
    ----------
    n_clusters : int
        The number of clusters to find.
    n_landmarks : int, optional
        Cluster only this many landmark frames (chosen by ``landmark_strategy``) and assign
        the rest by their distances to the landmarks.  None = every frame is a landmark.
    linkage : {'single', 'complete', 'average', 'ward'}, default='average'
        Linkage criterion of the tree; it is also the pooling function ``predict`` applies to
        the distances between a frame and the landmarks of a cluster.
    metric : string, default='euclidean'
        Any libdistance metric, or 'rmsd' on (n_frames, n_atoms, 3) coordinates.
    landmark_strategy : {'stride', 'random'}, default='stride'
    random_state : integer or numpy.RandomState, optional
        Used by landmark_strategy='random'.
    max_landmarks : int, optional, default=None
        If n_clusters exceeds n_landmarks, use max_landmarks landmarks instead.
    ward_predictor : {'single', 'complete', 'average', 'ward'}, default='ward'
        Pooling used by ``predict`` after a ward fit.

    Attributes
    ----------
    landmark_labels_ : np.array, [n_landmarks]
    landmarks_ : np.array, [n_landmarks, X.shape]
    cluster_centers_ : np.array, [n_clusters, X.shape]
        Mean of each cluster's landmarks (unless RMSD is the metric)
    """

    def __init__(self, n_clusters, n_landmarks=None, linkage='average',
                 metric='euclidean', landmark_strategy='stride',
                 random_state=None, max_landmarks=None, ward_predictor='ward'):
        self.n_clusters = n_clusters
        self.n_landmarks = n_landmarks
        self.metric = metric
        self.landmark_strategy = landmark_strategy
        self.random_state = random_state
        self.linkage = linkage
        self.max_landmarks = max_landmarks
        self.ward_predictor = ward_predictor

        self.landmark_labels_ = None
        self.landmarks_ = None
        self.cluster_centers_ = None

    # ------------------------------------------------------------------ fit
    def fit(self, X, y=None):
        import torch
        from .. import _kernels as K
        if callable(self.metric):
            raise TypeError("callable metrics run on the host in the reference; this device "
                            "implementation takes the libdistance metric names and 'rmsd'")
        if self.metric != 'rmsd':
            _lib.metric_id(self.metric)
        if self.max_landmarks is not None:
            if self.n_clusters > self.n_landmarks:
                self.n_landmarks = self.max_landmarks

        data, traces = _prepare(X, self.metric)
        n = int(data.shape[0])
        if self.n_landmarks is None:
            rows = None
            n_land = n
        else:
            if self.landmark_strategy == 'random':
                rows = check_random_state(self.random_state).randint(n, size=self.n_landmarks)
            else:
                rows = np.arange(n)[::(n // self.n_landmarks)][:self.n_landmarks]
            n_land = len(rows)

        if self.metric == 'rmsd':
            condensed = K.rmsd_pdist(data, traces, rows=rows).cpu().numpy()
        else:
            condensed = K.pdist(data, self.metric, rows=rows).cpu().numpy()
        tree = linkage(condensed, method=self.linkage)
        self.landmark_labels_ = fcluster(tree, criterion='maxclust', t=self.n_clusters) - 1
        self.cardinality_ = np.bincount(self.landmark_labels_)

        # sum of squared distances between landmarks that share a cluster (ward predictor);
        # pair k of the condensed matrix is (i, j), i < j, in row-major upper-triangle order
        iu, ju = np.triu_indices(n_land, k=1)
        same = self.landmark_labels_[iu] == self.landmark_labels_[ju]
        self.squared_distances_within_cluster_ = np.zeros(self.n_clusters)
        # unbuffered, in pair order: the same sequence of float64 additions as the reference's loop
        np.add.at(self.squared_distances_within_cluster_, self.landmark_labels_[iu[same]],
                  condensed[same] ** 2)

        if rows is None:
            land = data
        else:
            land = data[torch.from_numpy(np.asarray(rows, dtype=np.int64)).cuda()]
        # landmarks_ keeps the caller's view of the frames (uncentred for rmsd: predict centres
        # its own copy, like libdistance.cdist does)
        if self.metric == 'rmsd':
            src = X if rows is None else X[np.asarray(rows)] if isinstance(X, np.ndarray) else \
                X[torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(X.device)]
            self.landmarks_ = src.cpu().numpy() if hasattr(src, 'cpu') else np.asarray(src)
        else:
            self.landmarks_ = land.cpu().numpy()
            self.cluster_centers_ = np.array([
                list(np.mean(self.landmarks_[self.landmark_labels_ == i], axis=0))
                for i in range(self.n_clusters)])
        return self

    # -------------------------------------------------------------- predict
    def predict(self, X):
        """Predict the closest cluster each sample in X belongs to.

        Returns
        -------
        labels : array, shape [n_samples,]
        """
        import torch
        from .. import _kernels as K
        from .. import _device as dev
        pool_name = self.ward_predictor if self.linkage == 'ward' else self.linkage
        if pool_name not in _POOLS:
            raise ValueError("linkage {} is not supported".format(pool_name))
        data, traces = _prepare(X, self.metric)
        n = int(data.shape[0])
        land_labels = np.asarray(self.landmark_labels_)
        order = np.argsort(land_labels, kind='stable')       # landmarks grouped by cluster
        counts = np.bincount(land_labels, minlength=self.n_clusters)[:self.n_clusters]
        for i in np.nonzero(counts == 0)[0]:
            print("No data points were assigned to cluster {}".format(i))
        offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).cuda()
        card = torch.from_numpy(np.asarray(self.cardinality_, dtype=np.float64)).cuda()
        if card.numel() < self.n_clusters:
            card = torch.cat([card, torch.zeros(self.n_clusters - card.numel(), dtype=torch.float64,
                                                device="cuda")])
        sqs = torch.from_numpy(np.ascontiguousarray(self.squared_distances_within_cluster_,
                                                    dtype=np.float64)).cuda()
        land, land_tr = _prepare(np.ascontiguousarray(self.landmarks_[order]), self.metric)
        if self.metric != 'rmsd' and land.dtype != data.dtype:
            raise TypeError('XA and XB must be both float32 or float64')
        L = int(land.shape[0])
        labels = torch.empty(n, dtype=torch.int32, device="cuda")
        neg = torch.zeros(1, dtype=torch.int32, device="cuda")
        chunk = max(1, min(n, _CHUNK_BYTES // (8 * L)))
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            if self.metric == 'rmsd':
                d = torch.empty((b - a, L), dtype=torch.float64, device="cuda")
                tr_host = land_tr.cpu().numpy()
                for j in range(L):
                    d[:, j] = K.rmsd_dist(data[a:b], traces[a:b], land[j], float(tr_host[j]))
            else:
                d = K.cdist(data[a:b], land, self.metric)
            _lib.call("msmb200_pooled_assign", dev.ptr(d), b - a, L, dev.ptr(offsets),
                      int(self.n_clusters), _POOLS[pool_name], dev.ptr(card), dev.ptr(sqs),
                      dev.ptr(labels[a:b]), None, dev.ptr(neg), dev.stream_ptr())
        if int(neg.item()):
            warnings.warn("Distance shouldn't be negative.")
        return _to_host_int(labels)

    def fit_predict(self, X):
        """``fit(X)`` followed by ``predict(X)``."""
        self.fit(X)
        return self.predict(X)


class LandmarkAgglomerative(MultiSequenceClusterMixin, _LandmarkAgglomerative, BaseEstimator):
    __doc__ = _LandmarkAgglomerative.__doc__
    _allow_trajectory = True

    def fit_predict(self, sequences, y=None):
        self.fit(sequences)
        return self.predict(sequences)
