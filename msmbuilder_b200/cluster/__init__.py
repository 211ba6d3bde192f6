"""Clusterers of the hot path (see msmbuilder/cluster/__init__.py for the full
reference list; only the libdistance-driven assignment loops live here)."""
from .base import MultiSequenceClusterMixin
from .kcenters import KCenters
from .minibatchkmedoids import MiniBatchKMedoids
from .minibatchkmeans import MiniBatchKMeans
from .regularspatial import RegularSpatial
from .kmedoids import KMedoids
from .agglomerative import LandmarkAgglomerative

__all__ = ['KCenters', 'MiniBatchKMedoids', 'MiniBatchKMeans', 'RegularSpatial', 'KMedoids', 'LandmarkAgglomerative',
           'MultiSequenceClusterMixin']
