from .tica import tICA

__all__ = ['tICA']
