"""K-medoids with the all-pairs distance matrix computed on a B200.

Drop-in for ``msmbuilder.cluster.KMedoids`` (cluster/kmedoids.py:144) and
``_KMedoids`` (:20-141): same constructor, fitted attributes (``cluster_ids_``,
``cluster_centers_``, ``labels_``, ``inertia_``) and ``predict``.  ``fit`` is
``libdistance.pdist`` over ALL frames (kmedoids.py:91; N(N-1)/2 distances: the
GPU part, ``msmb200_pdist``) followed by ``n_passes`` randomly restarted descents
on the condensed matrix (kmedoids.py:92-94 -> kmedoids.cc:160-250), which are
sequential and stay on the host (``msmb200_kmedoids_restarts``); the random
starts come from the same RandomState calls in the same order as the reference's.
"""
from __future__ import absolute_import, print_function, division

from operator import itemgetter

import numpy as np
from sklearn.base import ClusterMixin, TransformerMixin

from .base import MultiSequenceClusterMixin
from .kcenters import _prepare
from ..base import BaseEstimator


def _to_host_intp(labels):
    import torch
    from .._device import to_host
    return to_host(labels, torch.int64)

__all__ = ['KMedoids']


class _KMedoids(ClusterMixin, TransformerMixin):
    """K-medoids: minimise the summed distance of every frame to the medoid (an
    actual frame) of its cluster, from ``n_passes`` random starts.  Needs the
    full pairwise distance matrix, so memory grows with the square of the
    number of frames.

    Parameters
    ----------
    n_clusters : int, optional, default: 8
        Number of clusters.
    n_passes : int, default=1
        Random restarts; the best solution is kept.
    metric : {"euclidean", "sqeuclidean", "cityblock", "chebyshev", "canberra",
              "braycurtis", "hamming", "jaccard", "cityblock", "rmsd"}
        Distance. 'rmsd' takes trajectories / (n, n_atoms, 3) coordinates.
    random_state : integer or numpy.RandomState, optional
        Draws the random starts; an integer fixes the seed.

    Attributes
    ----------
    cluster_ids_ : array, [n_clusters]
        Frame index of each medoid.
    cluster_centers_ : array, [n_clusters, n_features]
        The medoids themselves.
    labels_ : array, [n_samples,]
        Cluster number of each frame.
    inertia_ : float
        Summed distance of the frames to their medoid.
    """

    def __init__(self, n_clusters=8, n_passes=1, metric='euclidean', random_state=None):
        self.n_clusters = n_clusters
        self.n_passes = n_passes
        self.metric = metric
        self.random_state = random_state

    def fit(self, X, y=None):
        import torch
        from .. import _kernels as K
        if self.n_passes < 1:
            raise ValueError('n_passes must be greater than 0. got %s' % self.n_passes)
        if self.n_clusters < 1:
            raise ValueError('n_passes must be greater than 0. got %s' % self.n_clusters)

        data, traces = _prepare(X, self.metric)
        if self.metric == 'rmsd':
            dmat = K.rmsd_pdist(data, traces)
        else:
            dmat = K.pdist(data, self.metric)
        ids, self.inertia_, _ = K.kmedoids(self.n_clusters, dmat.cpu().numpy(), self.n_passes,
                                           random_state=self.random_state)
        self.labels_, mapping = K.contigify_ids(ids)
        smapping = sorted(mapping.items(), key=itemgetter(1))
        self.cluster_ids_ = np.array(smapping)[:, 0]
        idx = torch.as_tensor(self.cluster_ids_, dtype=torch.int64, device="cuda")
        self.cluster_centers_ = data[idx].cpu().numpy()
        return self

    def predict(self, X):
        """Index of the closest medoid of each frame of X."""
        import torch
        from .. import _kernels as K
        data, traces = _prepare(X, self.metric)
        if self.metric == 'rmsd':
            cent, ctr = _prepare(self.cluster_centers_, 'rmsd')
            labels, _, _ = K.rmsd_assign_nearest(data, traces, cent, ctr)
        else:
            cent = torch.from_numpy(np.ascontiguousarray(self.cluster_centers_)).cuda()
            if cent.dtype != data.dtype:
                raise TypeError('X and y must be both float32 or float64')
            labels, _, _ = K.assign_nearest(data, cent, self.metric)
        return _to_host_intp(labels)

    def fit_predict(self, X, y=None):
        return self.fit(X, y).labels_


class KMedoids(MultiSequenceClusterMixin, _KMedoids, BaseEstimator):
    __doc__ = _KMedoids.__doc__
    _allow_trajectory = True

    def fit(self, sequences, y=None):
        """Fit the clustering on a list of sequences (kmedoids.py:148-164)."""
        MultiSequenceClusterMixin.fit(self, sequences)
        self.cluster_ids_ = self._split_indices(self.cluster_ids_)
        return self
