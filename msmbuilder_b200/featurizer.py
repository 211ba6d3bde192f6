"""RMSDFeaturizer on the device RMSD kernel.

Mirror of ``msmbuilder.featurizer.RMSDFeaturizer`` (msmbuilder/featurizer/featurizer.py:255-321):
the feature vector of a frame is its RMSD, after optimal superposition, to every frame of a
reference trajectory -- ``libdistance.cdist(traj, reference, 'rmsd')`` in the reference
(featurizer.py:318-320), the same call on the device here (K6, csrc/rmsd.cu).

Trajectories are anything with an ``.xyz`` array of shape (n_frames, n_atoms, 3) (an
``md.Trajectory`` when mdtraj is installed) or such an array / CUDA tensor itself; mdtraj is not
required.  Like the reference, the featurizer never modifies its inputs: frames are centred in a
private device copy.
"""
from __future__ import absolute_import, print_function, division

import warnings

import numpy as np

from .base import BaseEstimator
from .utils import is_tensor, is_trajectory

__all__ = ['RMSDFeaturizer']


def _coords(traj):
    xyz = traj.xyz if is_trajectory(traj) else traj
    if not is_tensor(xyz):
        xyz = np.asarray(xyz)
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise ValueError("expected coordinates of shape (n_frames, n_atoms, 3)")
    return xyz


class RMSDFeaturizer(BaseEstimator):
    """Featurizer based on RMSD to one or more reference structures.

    Parameters
    ----------
    reference_traj : trajectory or array, shape=(n_ref_frames, n_atoms, 3)
        The reference conformations to superpose each frame with respect to
    atom_indices : np.ndarray, shape=(n_atoms,), dtype=int
        The indices of the atoms to superpose and compute the distances with.
        If not specified, all atoms are used.
    trj0
        Deprecated. Please use reference_traj.
    """

    def __init__(self, reference_traj=None, atom_indices=None, trj0=None):
        if trj0 is not None:
            warnings.warn("trj0 is deprecated. Please use reference_traj", DeprecationWarning)
            reference_traj = trj0
        elif reference_traj is None:
            raise ValueError("Please specify a reference trajectory")
        self.reference_traj = reference_traj
        self.trj0 = None
        ref = _coords(reference_traj)
        self.atom_indices = atom_indices
        if self.atom_indices is not None:
            idx = np.asarray(self.atom_indices, dtype=np.int64)
            self.sliced_reference_traj = ref[:, idx] if not is_tensor(ref) else \
                ref[:, _index_like(ref, idx)]
        else:
            self.sliced_reference_traj = ref
            self.atom_indices = [i for i in range(int(ref.shape[1]))]

    def _transform(self, value):
        return value

    def partial_transform(self, traj):
        """RMSD of every frame of `traj` to every reference frame.

        Returns
        -------
        features : np.ndarray, shape=(n_frames, n_ref_frames), float64
        """
        from . import libdistance
        xyz = _coords(traj)
        idx = np.asarray(self.atom_indices, dtype=np.int64)
        if len(idx) != int(xyz.shape[1]) or not np.array_equal(idx, np.arange(len(idx))):
            xyz = xyz[:, idx] if not is_tensor(xyz) else xyz[:, _index_like(xyz, idx)]
        return self._transform(libdistance.cdist(xyz, self.sliced_reference_traj, 'rmsd'))

    def fit(self, traj_list, y=None):
        return self

    def transform(self, traj_list, y=None):
        """Featurize a several trajectories: a list of (n_frames_i, n_ref_frames) arrays."""
        return [self.partial_transform(traj) for traj in traj_list]

    def fit_transform(self, traj_list, y=None):
        return self.fit(traj_list).transform(traj_list)

    def summarize(self):
        return "RMSDFeaturizer: %d reference frames x %d atoms" % (
            int(self.sliced_reference_traj.shape[0]), len(self.atom_indices))


def _index_like(t, idx):
    import torch
    return torch.from_numpy(idx).to(t.device)
