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
    devices : None, 'all' or list of CUDA ordinals, optional (``KCenters`` only)
        GPUs to shard the frames over from ONE process (a thread per GPU, candidate exchange by
        peer copies): ``fit`` then uses the whole box without torchrun.  None = current device.
        Same centres, labels and distances as on one GPU.

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

    def __init__(self, n_clusters=8, metric='euclidean', random_state=None, devices=None):
        _KCenters.__init__(self, n_clusters=n_clusters, metric=metric, random_state=random_state)
        self.devices = devices

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
        from .._device import resolve_devices
        devs = resolve_devices(self.devices)
        if devs is not None:
            return self._fit_on_devices(sequences, devs)
        MultiSequenceClusterMixin.fit(self, sequences)
        self.distances_ = self._split(self.distances_)
        return self

    def _fit_on_devices(self, sequences, devs):
        """One process, one thread per GPU: contiguous row ranges of the concatenated frames go to
        the devices (no host concatenation: every thread uploads slices of the caller's arrays into
        its own FrameStore), then the rank-collective loop of parallel.kcenters_fit_gpu runs over a
        thread communicator -- the protocol torchrun + NCCL would run, minus the processes."""
        import threading
        import torch
        from ..utils import check_iter_of_sequences
        from .base import _frames_of
        from .. import parallel as P
        from .. import _kernels as K
        from .._device import FrameStore, to_host
        check_iter_of_sequences(sequences, allow_trajectory=self._allow_trajectory)
        seqs = [_frames_of(s) for s in sequences]
        if len(seqs) == 0:
            raise TypeError('sequences must be a list of numpy arrays or ``md.Trajectory``s')
        lengths = [int(len(s)) for s in seqs]
        n_total = int(sum(lengths))
        # a device needs a few thousand frames to be worth a thread
        n_dev = max(1, min(len(devs), n_total // 4096))
        devs = devs[:n_dev]
        if n_dev == 1:
            with torch.cuda.device(devs[0]):
                MultiSequenceClusterMixin.fit(self, sequences)
                self.distances_ = self._split(self.distances_)
            return self
        starts = np.concatenate([[0], np.cumsum(lengths)])
        ranges = P.shard_rows(n_total, n_dev)
        seed = check_random_state(self.random_state).randint(0, n_total)      # kcenters.py:84
        k, metric = int(self.n_clusters), self.metric
        group = P.ThreadGroup(n_dev)
        results, errors = [None] * n_dev, []

        def work(r):
            try:
                lo, hi = ranges[r]
                with torch.cuda.device(devs[r]):
                    parts = []
                    for s, a in zip(seqs, starts[:-1]):
                        b0, b1 = max(lo, int(a)) - int(a), min(hi, int(a) + len(s)) - int(a)
                        if b1 > b0:
                            parts.append(s[b0:b1])
                    store = FrameStore(parts)
                    data, traces = store.data, None
                    if metric == 'rmsd':
                        if data.ndim != 3 or data.shape[2] != 3:
                            raise ValueError("metric='rmsd' needs coordinates of shape (n_frames, n_atoms, 3)")
                        data = data.to(torch.float32).clone()
                        traces = K.rmsd_center(data)
                    elif data.ndim != 2:
                        raise ValueError("expected a 2-D array of shape (n_samples, n_features)")
                    ids, distances, labels, ring = P.kcenters_fit_gpu(
                        data, lo, k, metric, seed, traces=traces, comm=group.comm(r))
                    row_elems = int(np.prod(data.shape[1:]))
                    es = data.element_size()
                    cent = None
                    if r == 0:
                        cent = ring[:k, P.CAND_HEADER:P.CAND_HEADER + row_elems * es].contiguous() \
                            .view(data.dtype).reshape((k,) + tuple(data.shape[1:])).cpu().numpy()
                    results[r] = (ids.cpu().numpy(), to_host(labels, torch.int64), to_host(distances),
                                  float(distances.sum().item()), cent)
            except BaseException as e:        # noqa: B902 -- the other threads must not wait for this one
                errors.append(e)
                group.abort()

        threads = [threading.Thread(target=work, args=(r,), name="msmb200-kcenters-%d" % devs[r])
                   for r in range(n_dev)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            first = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
            raise (first or errors)[0]
        self.cluster_ids_ = [int(c) for c in results[0][0]]
        self.cluster_centers_ = results[0][4]
        labels = np.concatenate([res[1] for res in results])
        distances = np.concatenate([res[2] for res in results])
        self.inertia_ = float(sum(res[3] for res in results))
        self._MultiSequenceClusterMixin__lengths = lengths
        self.labels_ = self._split(labels)
        self.distances_ = self._split(distances)
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
