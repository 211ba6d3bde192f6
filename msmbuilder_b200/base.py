"""Estimator base class of the drop-in boundary.

Mirrors msmbuilder/base.py:6-11 (sklearn BaseEstimator + ``summarize``).  When a
real ``msmbuilder`` is importable its BaseEstimator is mixed in as well, so that
``isinstance(est, msmbuilder.base.BaseEstimator)`` (checked by the reference's
tests/test_estimator_subclassing.py:52-55) holds for the replacements too.
"""
from sklearn.base import BaseEstimator as _SklearnBaseEstimator

try:  # pragma: no cover - msmbuilder is not installable in the build image
    from msmbuilder.base import BaseEstimator as _RefBaseEstimator
    if getattr(__import__("msmbuilder"), "__stub__", False):
        raise ImportError
except Exception:  # noqa: BLE001
    _RefBaseEstimator = None


if _RefBaseEstimator is not None:  # pragma: no cover
    class BaseEstimator(_RefBaseEstimator):
        pass
else:
    class BaseEstimator(_SklearnBaseEstimator):
        def summarize(self):
            """Return some diagnostic summary statistics about this model."""
            return 'NotImplemented'
