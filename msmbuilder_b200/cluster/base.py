"""Multi-sequence adaptor for clusterers, device-resident.

Same contract as ``msmbuilder.cluster.base.MultiSequenceClusterMixin``
(msmbuilder/cluster/base.py:17-173): ``fit`` takes a list of sequences,
concatenates them, runs the single-array ``fit`` of the class it is mixed into
and splits ``labels_`` back per sequence; ``predict`` works sequence by
sequence; ``transform`` / ``partial_transform`` / ``fit_transform`` are aliases.

What differs is WHERE the concatenation lives: the reference materialises one
host array (np.concatenate, base.py:58); here every sequence is copied straight
into its slot of one device buffer (``_device.FrameStore``), or adopted without
a copy when the caller already holds back-to-back CUDA tensors.
"""
from __future__ import absolute_import, print_function, division

import numpy as np

from ..utils import check_iter_of_sequences, is_trajectory, is_tensor

__all__ = ['MultiSequenceClusterMixin']


def _frames_of(seq):
    """ndarray / tensor view of one sequence (trajectory -> its coordinates)."""
    return seq.xyz if is_trajectory(seq) else seq


class MultiSequenceClusterMixin(object):
    _allow_trajectory = False

    def fit(self, sequences, y=None):
        """Fit the clustering on a list of sequences.

        Parameters
        ----------
        sequences : list of array-like, each of shape [sequence_length, n_features]
            Sequences may differ in length but not in width.

        Returns
        -------
        self
        """
        check_iter_of_sequences(sequences, allow_trajectory=self._allow_trajectory)
        super(MultiSequenceClusterMixin, self).fit(self._concat(sequences))

        if hasattr(self, 'labels_'):
            self.labels_ = self._split(self.labels_)

        return self

    def _concat(self, sequences):
        from .._device import FrameStore
        if hasattr(sequences, 'to_device'):
            # io.NumpyDirStream: files go straight into their slots of one device buffer
            seqs = sequences.to_device()
        else:
            seqs = list(sequences)
        self.__lengths = [len(s) for s in seqs]
        if len(seqs) == 0:
            raise TypeError('sequences must be a list of numpy arrays '
                            'or ``md.Trajectory``s')
        first = seqs[0]
        if not (isinstance(first, np.ndarray) or is_tensor(first) or is_trajectory(first)):
            raise TypeError('sequences must be a list of numpy arrays '
                            'or ``md.Trajectory``s')
        store = FrameStore([_frames_of(s) for s in seqs])
        assert sum(self.__lengths) == store.n
        return store.data

    def _split(self, concat):
        return [concat[cl - l: cl] for (cl, l) in zip(np.cumsum(self.__lengths), self.__lengths)]

    def _split_indices(self, concat_inds):
        """Indices in concatenated space -> (traj_i, frame_i) pairs (base.py:79-88),
        without materialising an N x 2 table."""
        clengths = np.append([0], np.cumsum(self.__lengths))
        concat_inds = np.asarray(concat_inds, dtype=np.int64)
        traj = np.searchsorted(clengths, concat_inds, side='right') - 1
        out = np.zeros((len(concat_inds), 2), dtype=int)
        out[:, 0] = traj
        out[:, 1] = concat_inds - clengths[traj]
        return out

    def predict(self, sequences, y=None):
        """Closest cluster centre of every frame of every sequence.

        Returns
        -------
        Y : list of arrays, each of shape [sequence_length,]
        """
        predictions = []
        check_iter_of_sequences(sequences, allow_trajectory=self._allow_trajectory)
        for X in sequences:
            predictions.append(self.partial_predict(X))
        return predictions

    def partial_predict(self, X, y=None):
        """Closest cluster centre of every frame of ONE sequence."""
        return super(MultiSequenceClusterMixin, self).predict(_frames_of(X))

    def fit_predict(self, sequences, y=None):
        """Cluster the sequences and return their labels."""
        if hasattr(super(MultiSequenceClusterMixin, self), 'fit_predict'):
            check_iter_of_sequences(sequences, allow_trajectory=self._allow_trajectory)
            labels = super(MultiSequenceClusterMixin, self).fit_predict(sequences)
        else:
            self.fit(sequences)
            labels = self.predict(sequences)

        if not isinstance(labels, list):
            labels = self._split(labels)
        return labels

    def transform(self, sequences):
        """Alias for predict"""
        return self.predict(sequences)

    def partial_transform(self, X):
        """Alias for partial_predict"""
        return self.partial_predict(X)

    def fit_transform(self, sequences, y=None):
        """Alias for fit_predict"""
        return self.fit_predict(sequences, y)
