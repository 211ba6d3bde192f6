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


class _MedoidSearch(object):
    """The sequential part of MiniBatchKMedoids.fit (minibatchkmedoids.py:90-134): a host loop over
    small problems -- k current medoids plus `batch_size` random frames -- whose pairwise distances
    come from the device (K4, gathered rows, nothing but the (k + batch)^2 / 2 distances leaves
    the GPU) and whose k-medoids sweep runs in msmb200_kmedoids (host C++).

    Parity with the reference depends on the ORDER of the RandomState draws: first the k starting
    medoids, then a provisional medoid number for every frame, then one draw of `batch_size` frame
    indices per round (the k-medoids call itself draws nothing with npass = 0).  Everything else
    here is bookkeeping and is written around three arrays:

      medoids[k]    frame index of each medoid
      member_of[n]  medoid number last given to each frame (only touched frames are meaningful)
      quiet         consecutive rounds that changed no frame's medoid number
    """

    def __init__(self, est, data, traces):
        self.k = int(est.n_clusters)
        self.batch = int(est.batch_size)
        self.metric = est.metric
        self.patience = est.max_no_improvement
        self.data, self.traces = data, traces
        self.n = int(data.shape[0])
        rounds_per_sweep = -(-self.n // self.batch)              # ceil(n / batch)
        self.max_rounds = int(est.max_iter) * rounds_per_sweep
        self.rng = check_random_state(est.random_state)
        self.medoids = self.rng.randint(0, self.n, size=self.k)          # draw 1
        self.member_of = self.rng.randint(0, self.k, size=self.n)        # draw 2
        self.quiet = 0

    def _distances(self, frames):
        from .. import _kernels as K
        rows = np.asarray(frames, dtype=np.intp)
        if self.metric == 'rmsd':
            return K.rmsd_pdist(self.data, self.traces, rows=rows).cpu().numpy()
        return K.pdist(self.data, self.metric, rows=rows).cpu().numpy()

    def _round(self):
        """One mini-batch: returns False when the search has gone quiet for long enough."""
        from .. import _kernels as K
        k = self.k
        fresh = self.rng.randint(0, self.n, self.batch)                  # draw 3, 4, ...
        frames = np.concatenate([self.medoids, fresh])                   # positions 0..k-1 = medoids
        # starting assignment inside the batch: a medoid belongs to itself, a fresh frame keeps
        # the number it had
        start = np.concatenate([np.arange(k), self.member_of[fresh]]).astype(np.intp)
        solved, _, _ = K.kmedoids(k, self._distances(frames), 0, start, random_state=self.rng)
        # the sweep names a cluster by the batch position of its medoid; renumber 0..k-1 in order
        # of first appearance and remember which position carries each number
        numbered, position_to_number = K.contigify_ids(solved)
        by_number = sorted(position_to_number, key=position_to_number.get)
        self.medoids = frames[np.asarray(by_number, dtype=np.intp)]
        if np.array_equal(self.member_of[frames], numbered):
            self.quiet += 1
        else:
            self.member_of[frames] = numbered
            self.quiet = 0
        return self.quiet < self.patience

    def run(self):
        for _ in range(self.max_rounds):
            if not self._round():
                break


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
        search = _MedoidSearch(self, data, traces)
        search.run()

        import torch
        self.cluster_ids_ = search.medoids
        idx = torch.from_numpy(np.asarray(search.medoids, dtype=np.int64)).cuda()
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
