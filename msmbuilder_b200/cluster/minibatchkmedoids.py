"""Mini-batch k-medoids with the distance work on a B200.

Drop-in for ``msmbuilder.cluster.MiniBatchKMedoids``
(msmbuilder/cluster/minibatchkmedoids.py:170) and its single-array core
(minibatchkmedoids.py:24-167).  Per iteration: the (n_clusters + batch_size)
gathered rows go through the device ``pdist`` (K4), the condensed matrix comes
back in one D2H copy, the tiny sequential k-medoids step runs on the host
(``msmb200_kmedoids``), and at the end ONE device ``assign_nearest`` pass (K3)
labels every frame.  The NumPy RandomState call order of the reference
(minibatchkmedoids.py:99,100,108) is preserved so that equal seeds give equal
medoids.
"""
from __future__ import absolute_import, print_function, division

from operator import itemgetter

import numpy as np
from sklearn.utils import check_random_state
from sklearn.base import ClusterMixin, TransformerMixin

from .base import MultiSequenceClusterMixin
from .kcenters import _prepare
from ..base import BaseEstimator
from .. import _lib


def _to_host_intp(labels):
    import torch
    from .._device import to_host
    return to_host(labels, torch.int64)

__all__ = ['MiniBatchKMedoids']


class _MiniBatchKMedoids(ClusterMixin, TransformerMixin):
    """Mini-batch k-medoids: k-medoids sweeps on small random batches that always
    contain the current medoids, followed by one full assignment pass.

    Parameters
    ----------
    n_clusters : int, optional, default: 8
        Number of medoids.
    max_iter : int, optional, default=5
        Passes over the data, in units of n_samples / batch_size batches.
    batch_size : int, optional, default=100
        Random frames per batch; memory grows with its square.
    metric : {"euclidean", "sqeuclidean", "cityblock", "chebyshev", "canberra",
              "braycurtis", "hamming", "jaccard", "cityblock", "rmsd"}
        Distance. 'rmsd' takes trajectories / (n, n_atoms, 3) coordinates.
    max_no_improvement : int, default=10
        Stop after this many consecutive batches that change no label.
    random_state : integer or numpy.RandomState, optional
        Generator for the initial medoids / labels and the batches.

    Attributes
    ----------
    cluster_ids_ : array, [n_clusters]
        Index of the frame that each medoid is.
    labels_ : array, [n_samples,]
        Medoid number of each frame.
    inertia_ : float
        Sum of distances of frames to their medoid.
    """

    def __init__(self, n_clusters=8, max_iter=5, batch_size=100,
                 metric='euclidean', max_no_improvement=10, random_state=None):
        self.n_clusters = n_clusters
        self.batch_size = batch_size
        self.max_iter = max_iter
        self.max_no_improvement = max_no_improvement
        self.metric = metric
        self.random_state = random_state

    def fit(self, X, y=None):
        from .. import _kernels as K
        if self.metric != 'rmsd':
            _lib.metric_id(self.metric)   # ValueError on an unknown metric, before any work
        data, traces = _prepare(X, self.metric)
        n_samples = int(data.shape[0])
        n_batches = int(np.ceil(float(n_samples) / self.batch_size))
        n_iter = int(self.max_iter * n_batches)
        random_state = check_random_state(self.random_state)

        cluster_ids_ = random_state.randint(0, n_samples, size=self.n_clusters)
        labels_ = random_state.randint(0, self.n_clusters, size=n_samples)

        n_iters_no_improvement = 0
        for kk in range(n_iter):
            # batch = current medoids + fresh random frames
            minibatch_indices = np.concatenate([
                cluster_ids_,
                random_state.randint(0, n_samples, self.batch_size),
            ])
            rows = np.array(minibatch_indices, dtype=np.intp)
            if self.metric == 'rmsd':
                dmat = K.rmsd_pdist(data, traces, rows=rows).cpu().numpy()
            else:
                dmat = K.pdist(data, self.metric, rows=rows).cpu().numpy()
            minibatch_labels = np.array(np.concatenate([
                np.arange(self.n_clusters),
                labels_[minibatch_indices[self.n_clusters:]]
            ]), dtype=np.intp)

            ids, intertia, _ = K.kmedoids(self.n_clusters, dmat, 0, minibatch_labels,
                                          random_state=random_state)
            minibatch_labels, m = K.contigify_ids(ids)

            # new medoids, in label order
            minibatch_cluster_ids = np.array(sorted(m.items(), key=itemgetter(1)))[:, 0]
            cluster_ids_ = minibatch_indices[minibatch_cluster_ids]

            n_changed = np.sum(labels_[minibatch_indices] != minibatch_labels)
            if n_changed == 0:
                n_iters_no_improvement += 1
            else:
                labels_[minibatch_indices] = minibatch_labels
                n_iters_no_improvement = 0
            if n_iters_no_improvement >= self.max_no_improvement:
                break

        import torch
        self.cluster_ids_ = cluster_ids_
        idx = torch.from_numpy(np.asarray(cluster_ids_, dtype=np.int64)).cuda()
        centers = data[idx].contiguous()
        self.cluster_centers_ = centers.cpu().numpy()
        if self.metric == 'rmsd':
            labels, _, inertia = K.rmsd_assign_nearest(data, traces, centers, traces[idx].contiguous())
        else:
            labels, _, inertia = K.assign_nearest(data, centers, self.metric)
        self.labels_ = _to_host_intp(labels)
        self.inertia_ = float(inertia)
        return self

    def predict(self, X):
        """Index of the closest medoid of each frame of X ([n_samples, n_features])."""
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


class MiniBatchKMedoids(MultiSequenceClusterMixin, _MiniBatchKMedoids, BaseEstimator):
    _allow_trajectory = True
    __doc__ = _MiniBatchKMedoids.__doc__[: _MiniBatchKMedoids.__doc__.find('Attributes')] + \
    '''
    Attributes
    ----------
    `cluster_centers_` : array, [n_clusters, n_features]
        Coordinates of cluster centers

    `labels_` : list of arrays, each of shape [sequence_length, ]
        Medoid number of each frame, one array per sequence.
    '''

    def fit(self, sequences, y=None):
        """Fit the clustering on the data

        Parameters
        ----------
        sequences : list of array-like, each of shape [sequence_length, n_features]
            A list of multivariate timeseries, or trajectories for metric='rmsd'.

        Returns
        -------
        self
        """
        MultiSequenceClusterMixin.fit(self, sequences)
        self.cluster_ids_ = self._split_indices(self.cluster_ids_)
        return self

    def summarize(self):
        return """MiniBatchKMedoids clustering
----------------------------
n_clusters : {n_clusters}
metric     : {metric}

Inertia    : {inertia_}
""".format(**self.__dict__)
