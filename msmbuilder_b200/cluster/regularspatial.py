"""Regular spatial clustering with the distance work on a B200.

Drop-in for ``msmbuilder.cluster.RegularSpatial`` (cluster/regularspatial.py:106)
and ``_RegularSpatial`` (:19-103): same constructor and fitted attributes
(``cluster_center_indices_``, ``cluster_centers_``, ``n_clusters_``) and
``predict``.  The reference walks the frames one by one and measures each against
the growing centre list (regularspatial.py:70-77, one ``libdistance.dist`` call
per frame); here every accepted centre triggers one streaming pass that lowers
the running minimum of all LATER frames (``msmb200_kcenters_pass``) and one scan
for the first frame that is still farther than ``d_min`` from everything
(``msmb200_first_above``).  Same centres, n_clusters_ passes instead of N calls.
"""
from __future__ import absolute_import, print_function, division

import numpy as np
from sklearn.base import ClusterMixin, TransformerMixin

from .base import MultiSequenceClusterMixin
from .kcenters import _prepare
from ..base import BaseEstimator


def _to_host_intp(labels):
    import torch
    from .._device import to_host
    return to_host(labels, torch.int64)

__all__ = ['RegularSpatial']


class _RegularSpatial(ClusterMixin, TransformerMixin):
    """Pick centres so that no two of them are closer than ``d_min``: the first
    frame is a centre; a later frame becomes one when it is farther than
    ``d_min`` from every centre chosen before it.

    Parameters
    ----------
    d_min : float
        Minimum distance between cluster centres.
    metric : {"euclidean", "sqeuclidean", "cityblock", "chebyshev", "canberra",
              "braycurtis", "hamming", "jaccard", "cityblock", "rmsd"}
        Distance. 'rmsd' takes trajectories / (n, n_atoms, 3) coordinates.

    Attributes
    ----------
    cluster_center_indices_ : list
        Frame index of each centre (for the multi-sequence class: array of
        (sequence, frame) pairs).
    cluster_centers_ : array, [n_clusters, n_features]
        The centres themselves.
    n_clusters_ : int
        How many were found.
    """

    def __init__(self, d_min, metric='euclidean'):
        self.d_min = d_min
        self.metric = metric

    def fit(self, X, y=None):
        import torch
        from .. import _kernels as K
        data, traces = _prepare(X, self.metric)
        ids = K.regular_spatial_fit(data, self.d_min, self.metric, traces=traces)
        self.cluster_center_indices_ = ids
        idx = torch.as_tensor(ids, dtype=torch.int64, device="cuda")
        self.cluster_centers_ = data[idx].cpu().numpy()
        self.n_clusters_ = len(ids)
        return self

    def predict(self, X):
        """Index of the closest centre of each frame of X."""
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
        return self.fit(X, y=y).predict(X)


class RegularSpatial(MultiSequenceClusterMixin, _RegularSpatial, BaseEstimator):
    __doc__ = _RegularSpatial.__doc__
    _allow_trajectory = True

    def fit(self, sequences, y=None):
        """Fit the clustering on a list of sequences (regularspatial.py:110-126)."""
        MultiSequenceClusterMixin.fit(self, sequences)
        self.cluster_center_indices_ = self._split_indices(self.cluster_center_indices_)
        return self
