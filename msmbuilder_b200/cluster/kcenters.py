"""K-centers clustering with every pass over the frames on a B200.

Drop-in for ``msmbuilder.cluster.KCenters`` (msmbuilder/cluster/kcenters.py:132)
and its single-array core ``_KCenters`` (kcenters.py:21-129): same constructor,
fitted attributes (``cluster_ids_``, ``cluster_centers_``, ``labels_``,
``distances_``, ``inertia_``) and ``predict``.  Each of the k sequential passes
of Gonzalez' algorithm (kcenters.py:91-97: one-to-all distance, strict running
minimum, label update, arg-max) is ONE fused streaming kernel
(``msmb200_kcenters_pass``); the arg-max that names the next centre stays on
the device, so the k launches queue back to back without a host round trip.
"""
from __future__ import absolute_import, print_function, division

import numpy as np
from sklearn.utils import check_random_state
from sklearn.base import ClusterMixin, TransformerMixin

from .base import MultiSequenceClusterMixin
from ..base import BaseEstimator
from ..utils import is_tensor

__all__ = ['KCenters']


def _prepare(X, metric):
    """-> (device tensor, traces or None).  rmsd: centred private copy + traces."""
    import torch
    from .. import _device as dev
    from .. import _kernels as K
    if metric == 'rmsd':
        t = X if is_tensor(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
        if t.ndim != 3 or t.shape[2] != 3:
            raise ValueError("metric='rmsd' needs coordinates of shape (n_frames, n_atoms, 3)")
        t = t.to(device="cuda", dtype=torch.float32).contiguous().clone()
        return t, K.rmsd_center(t)
    t = dev.to_device(X)
    if t.ndim != 2:
        raise ValueError("expected a 2-D array of shape (n_samples, n_features)")
    return t, None


class _KCenters(ClusterMixin, TransformerMixin):
    """Gonzalez k-centers: repeatedly promote the frame farthest from all current
    centres to be the next centre.  Runtime O(k N); 2-approximation of the
    minimax radius.

    Parameters
    ----------
    n_clusters : int, optional, default: 8
        Number of centres to pick.
    metric : {"euclidean", "sqeuclidean", "cityblock", "chebyshev", "canberra",
              "braycurtis", "hamming", "jaccard", "cityblock", "rmsd"}
        Distance. 'rmsd' takes trajectories / (n, n_atoms, 3) coordinates.
    random_state : integer or numpy.RandomState, optional
        Draws the first centre; an integer fixes the seed.

    Attributes
    ----------
    cluster_ids_ : array, [n_clusters]
        Index of the frame that each centre is.
    cluster_centers_ : array, [n_clusters, n_features]
        The centres themselves.
    labels_ : array, [n_samples,]
        Centre number of each frame.
    distances_ : array, [n_samples,]
        Distance of each frame to its centre.
    inertia_ : float
        Sum of ``distances_``.
    """

    def __init__(self, n_clusters=8, metric='euclidean', random_state=None):
        self.n_clusters = n_clusters
        self.metric = metric
        self.random_state = random_state

    def fit(self, X, y=None):
        from .. import _kernels as K
        data, traces = _prepare(X, self.metric)
        n_samples = int(data.shape[0])
        seed = check_random_state(self.random_state).randint(0, n_samples)   # kcenters.py:84

        ids, distances, labels = K.kcenters_fit(data, self.n_clusters, self.metric, seed,
                                                traces=traces)
        cluster_ids = ids.cpu().numpy()
        self.cluster_ids_ = [int(c) for c in cluster_ids]
        from .. import _device as dev
        import torch
        self.labels_ = dev.to_host(labels, torch.int64)       # == .astype(int) of the reference dtype
        self.distances_ = dev.to_host(distances)
        centers = data[ids]
        self.cluster_centers_ = centers.cpu().numpy()
        # float64 sum on the device, fixed tree order (reference: np.sum on the host)
        self.inertia_ = float(distances.sum().item())
        return self

    def predict(self, X):
        """Index of the closest centre of each frame of X ([n_samples, n_features])."""
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
        from .. import _device as dev
        return dev.to_host(labels, torch.int64)

    def fit_predict(self, X, y=None):
        return self.fit(X, y).labels_


class KCenters(MultiSequenceClusterMixin, _KCenters, BaseEstimator):
    _allow_trajectory = True
    __doc__ = _KCenters.__doc__[: _KCenters.__doc__.find('Attributes')] + \
    '''
    Attributes
    ----------
    `cluster_centers_` : array, [n_clusters, n_features]
        Coordinates of cluster centers

    `labels_` : list of arrays, each of shape [sequence_length, ]
        Centre number of each frame, one array per sequence.

    `distances_` : list of arrays, each of shape [sequence_length, ]
        Distance of each frame to its centre, one array per sequence.
    '''

    def fit(self, sequences, y=None):
        """Fit the kcenters clustering on the data

        Parameters
        ----------
        sequences : list of array-like, each of shape [sequence_length, n_features]
            A list of multivariate timeseries (NumPy, or torch on host / GPU), or
            trajectories for metric='rmsd'.

        Returns
        -------
        self
        """
        MultiSequenceClusterMixin.fit(self, sequences)
        self.distances_ = self._split(self.distances_)
        return self

    def summarize(self):
        return """KCenters clustering
--------------------
n_clusters : {n_clusters}
metric     : {metric}

Inertia       : {inertia}
Mean distance : {mean_distance}
Max  distance : {max_distance}
""".format(n_clusters=self.n_clusters, metric=self.metric,
           inertia=self.inertia_, mean_distance=np.mean(np.concatenate(self.distances_)),
           max_distance=np.max(np.concatenate(self.distances_)))
