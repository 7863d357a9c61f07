"""NumPy stand-in for the parts of `jax` the reference's hot path uses (TEST INFRASTRUCTURE ONLY).

See oracle/jaxshim/README.md.  Not a general JAX emulator.
"""
import numpy as _np

from . import _core
from ._core import Arr, narrow, to_arr
from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, nn, random, scipy  # noqa: F401

__version__ = "0.0-numpy-shim"


def jit(fun=None, **kwargs):
    """Identity: no tracing, no compilation (static_argnames / donate_argnums ignored)."""
    if fun is None:
        return lambda f: f
    return fun


def _leaf_index(x, i):
    if x is None:
        return None
    if isinstance(x, (tuple, list)):
        return type(x)(_leaf_index(v, i) for v in x)
    if isinstance(x, dict):
        return {k: _leaf_index(v, i) for k, v in x.items()}
    return x[i]


def _leaf_len(x):
    if x is None:
        return None
    if isinstance(x, (tuple, list)):
        for v in x:
            n = _leaf_len(v)
            if n is not None:
                return n
        return None
    if isinstance(x, dict):
        return _leaf_len(tuple(x.values()))
    return x.shape[0]


def _stack(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([o[k] for o in outs]) for k in range(len(first)))
    if first is None:
        return None
    return narrow(_np.stack([_np.asarray(o) for o in outs], axis=0))


def _zero_elem(x):
    if x is None:
        return None
    if isinstance(x, (tuple, list)):
        return type(x)(_zero_elem(v) for v in x)
    return narrow(_np.zeros(x.shape[1:], x.dtype))


def _empty_like_tree(x):
    if isinstance(x, (tuple, list)):
        return type(x)(_empty_like_tree(v) for v in x)
    if x is None:
        return None
    a = _np.asarray(x)
    return narrow(_np.zeros((0,) + a.shape, a.dtype))


def vmap(fun, in_axes=0, out_axes=0):
    """Map over the leading axis of every argument with a Python loop, stack the outputs."""
    assert in_axes == 0 and out_axes == 0, "shim: only leading-axis vmap"

    def mapped(*args):
        args = tuple(to_arr(a) if not isinstance(a, (tuple, list, dict)) else a for a in args)
        n = _leaf_len(args)
        if n == 0:
            # empty batch: evaluate once on a zero element to learn the output structure
            probe = fun(*_zero_elem(args))
            return _empty_like_tree(probe)
        outs = [fun(*_leaf_index(args, i)) for i in range(n)]
        return _stack(outs)

    return mapped


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
