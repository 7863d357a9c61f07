"""jax.numpy stand-in on NumPy float32 (TEST INFRASTRUCTURE ONLY)."""
import numpy as _np

from .._core import Arr, narrow, to_arr
from . import linalg  # noqa: F401

ndarray = _np.ndarray
pi = _np.pi
inf = _np.inf
float32 = _np.float32
int32 = _np.int32


def _raw(x):
    if isinstance(x, (tuple, list)):
        return type(x)(_raw(v) for v in x)
    return _np.asarray(x) if isinstance(x, Arr) else x


def _wrap(f):
    def g(*args, **kwargs):
        args = tuple(_raw(a) for a in args)
        kwargs = {k: _raw(v) for k, v in kwargs.items()}
        return narrow(f(*args, **kwargs))
    g.__name__ = getattr(f, "__name__", "wrapped")
    return g


def __getattr__(name):
    f = getattr(_np, name)
    if callable(f) and not isinstance(f, type):
        return _wrap(f)
    return f


def array(x, dtype=None):
    return narrow(_np.array(_raw(x), dtype=dtype))


def asarray(x, dtype=None):
    return narrow(_np.asarray(_raw(x), dtype=dtype))


def zeros(shape, dtype=None):
    return narrow(_np.zeros(shape, dtype=_np.float32 if dtype is None else _dt(dtype)))


def ones(shape, dtype=None):
    return narrow(_np.ones(shape, dtype=_np.float32 if dtype is None else _dt(dtype)))


def full(shape, fill_value, dtype=None):
    return narrow(_np.full(shape, fill_value, dtype=None if dtype is None else _dt(dtype)))


def _dt(dtype):
    if dtype is int:
        return _np.int32
    if dtype is float:
        return _np.float32
    if dtype is bool:
        return _np.bool_
    return dtype


def arange(start, stop=None, step=None, dtype=None):
    if stop is None:
        start, stop = 0, start
    return narrow(_np.arange(start, stop, step, dtype=None if dtype is None else _dt(dtype)))


def eye(n, dtype=None):
    return narrow(_np.eye(n, dtype=_np.float32 if dtype is None else _dt(dtype)))


def linspace(start, stop, num=50, dtype=None):
    """JAX's float32 linspace: start*(1-s) + stop*s for s = i/(num-1), endpoint appended; integer dtype: floor."""
    start = _np.asarray(_raw(start), _np.float32)
    stop = _np.asarray(_raw(stop), _np.float32)
    div = _np.float32(num - 1)
    s = (_np.arange(num - 1, dtype=_np.float32) / div).astype(_np.float32)
    shp = (num - 1,) + (1,) * start.ndim
    s = s.reshape(shp)
    body = (start[None, ...] * (_np.float32(1) - s)).astype(_np.float32) + (stop[None, ...] * s).astype(_np.float32)
    out = _np.concatenate((body, _np.broadcast_to(stop, start.shape)[None, ...]), axis=0)
    if dtype is not None and _np.issubdtype(_dt(dtype), _np.integer):
        out = _np.floor(out).astype(_dt(dtype))
    return narrow(out)


def clip(x, a_min=None, a_max=None):
    x = _np.asarray(_raw(x))
    if a_min is not None:
        x = _np.maximum(x, _raw(a_min))
    if a_max is not None:
        x = _np.minimum(x, _raw(a_max))
    return narrow(x)


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    return narrow(_np.nan_to_num(_np.asarray(_raw(x)), nan=nan, posinf=posinf, neginf=neginf))


def nonzero(a, size=None, fill_value=None):
    a = _np.asarray(_raw(a))
    idx = _np.nonzero(a)
    if size is None:
        return tuple(narrow(i) for i in idx)
    out = []
    for k, i in enumerate(idx):
        fv = 0 if fill_value is None else (fill_value[k] if isinstance(fill_value, (tuple, list)) else fill_value)
        o = _np.full((size,), fv, dtype=_np.int64)
        m = min(size, i.shape[0])
        o[:m] = i[:m]
        out.append(narrow(o))
    return tuple(out)


def argsort(a, axis=-1, kind=None, order=None):
    return narrow(_np.argsort(_np.asarray(_raw(a)), axis=axis, kind="stable"))


def meshgrid(*xs, indexing="xy"):
    return [narrow(g) for g in _np.meshgrid(*[_raw(x) for x in xs], indexing=indexing)]


def product(a, axis=None):
    return narrow(_np.prod(_raw(a), axis=axis))


def sum(a, axis=None, keepdims=False):  # noqa: A001
    a = _np.asarray(_raw(a))
    if a.dtype == _np.bool_:
        return narrow(_np.sum(a, axis=axis, keepdims=keepdims, dtype=_np.int32))
    return narrow(_np.sum(a, axis=axis, keepdims=keepdims))


def where(c, x=None, y=None):
    if x is None:
        return nonzero(c)
    return narrow(_np.where(_raw(c), _raw(x), _raw(y)))
