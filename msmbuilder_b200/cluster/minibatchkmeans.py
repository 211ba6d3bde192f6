"""MiniBatchKMeans whose assignment passes run on a B200.

The reference class is a three-line mix of MultiSequenceClusterMixin and
scikit-learn's estimator (msmbuilder/cluster/__init__.py:67-69); all of its
arithmetic is scikit-learn's.  Here the centre updates stay with scikit-learn on
the host (small mini-batches, sequential), while the two full-data steps -- the
final labelling pass of ``fit`` and every ``predict`` -- are the Euclidean
arg-min over centres of the hot path and run through ``msmb200_assign_nearest``.
"""
from __future__ import absolute_import, print_function, division

import numpy as np
from sklearn import cluster as _skc

from .base import MultiSequenceClusterMixin
from ..base import BaseEstimator
from ..utils import is_tensor

__all__ = ['MiniBatchKMeans']


class _GpuAssignMiniBatchKMeans(_skc.MiniBatchKMeans):
    def fit(self, X, y=None, sample_weight=None):
        import torch
        from .. import _kernels as K
        from .. import _device as dev
        X_dev = dev.to_device(X)
        X_host = X_dev.cpu().numpy() if is_tensor(X) else np.ascontiguousarray(X)
        want_labels = self.compute_labels
        self.compute_labels = False          # skip sklearn's full-data pass
        try:
            super(_GpuAssignMiniBatchKMeans, self).fit(X_host, y, sample_weight=sample_weight)
        finally:
            self.compute_labels = want_labels
        if want_labels:
            cent = torch.from_numpy(np.ascontiguousarray(self.cluster_centers_)).cuda().to(X_dev.dtype)
            labels, dmin, _ = K.assign_nearest(X_dev, cent, 'sqeuclidean', want_min_dist=True)
            self.labels_ = labels.cpu().numpy()
            self.inertia_ = float(dmin.sum().item())
        return self

    def predict(self, X):
        import torch
        from .. import _kernels as K
        from .. import _device as dev
        X_dev = dev.to_device(X)
        cent = torch.from_numpy(np.ascontiguousarray(self.cluster_centers_)).cuda().to(X_dev.dtype)
        labels, _, _ = K.assign_nearest(X_dev, cent, 'sqeuclidean')
        return labels.cpu().numpy()


class MiniBatchKMeans(MultiSequenceClusterMixin, _GpuAssignMiniBatchKMeans, BaseEstimator):
    __doc__ = _skc.MiniBatchKMeans.__doc__
