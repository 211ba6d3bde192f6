"""msmbuilder_b200 -- the tICA + clustering hot path of MSMBuilder on NVIDIA B200.

    from msmbuilder_b200.decomposition import tICA
    from msmbuilder_b200.cluster import KCenters, MiniBatchKMedoids, MiniBatchKMeans
    from msmbuilder_b200 import libdistance

Same estimator API as msmbuilder (fit / partial_fit / transform / predict on
lists of sequences); the arithmetic runs as hand-written sm_100a CUDA behind the
C ABI in include/msmb200.h.  No CPU fallback.
"""
__version__ = "0.1.0"


def device_sequences(sequences):
    """List of host arrays -> list of CUDA tensors in one device allocation
    (upload once, then hand the same list to tICA and to the clusterers)."""
    from ._device import device_sequences as _ds
    return _ds(sequences)
