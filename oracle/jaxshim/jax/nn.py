"""jax.nn stand-in (TEST INFRASTRUCTURE ONLY)."""
import numpy as _np

from ._core import narrow


def relu(x):
    return narrow(_np.maximum(_np.asarray(x), 0))


def elu(x, alpha=1.0):
    x = _np.asarray(x)
    safe = _np.where(x > 0, _np.zeros_like(x), x)
    return narrow(_np.where(x > 0, x, alpha * _np.expm1(safe)))


class initializers:  # only referenced by the fitting code, never called on the hot path
    @staticmethod
    def glorot_normal():
        raise NotImplementedError

    @staticmethod
    def normal():
        raise NotImplementedError
