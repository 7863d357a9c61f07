"""jax.lax stand-in (TEST INFRASTRUCTURE ONLY)."""
import numpy as _np

from ._core import narrow, to_arr


def _tree_to_arr(x):
    if isinstance(x, (tuple, list)):
        return type(x)(_tree_to_arr(v) for v in x)
    return to_arr(x)


def fori_loop(lower, upper, body_fun, init_val):
    val = _tree_to_arr(init_val)        # JAX turns Python scalars into arrays (so `~False` is logical)
    for i in range(int(lower), int(upper)):
        val = _tree_to_arr(body_fun(i, val))
    return val


def _clamp_start(n, start, size):
    start = int(start)
    if start < 0:
        start += n
    return max(0, min(start, n - size))


def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    a = _np.asarray(operand)
    s = _clamp_start(a.shape[axis], start_index, slice_size)
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(s, s + slice_size)
    return narrow(a[tuple(sl)].copy())


def dynamic_update_slice_in_dim(operand, update, start_index, axis=0):
    a = _np.array(operand, copy=True)
    u = _np.asarray(update)
    s = _clamp_start(a.shape[axis], start_index, u.shape[axis])
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(s, s + u.shape[axis])
    a[tuple(sl)] = u
    return narrow(a)


def map(f, xs):  # noqa: A001
    outs = [f(xs[i]) for i in range(xs.shape[0])]
    return narrow(_np.stack([_np.asarray(o) for o in outs], axis=0))


def pow(x, y):  # noqa: A001
    return narrow(_np.power(_np.asarray(x, _np.float32), _np.asarray(y, _np.float32)))


def top_k(x, k):
    x = _np.asarray(x)
    idx = _np.argsort(-x, kind="stable")[..., :k]
    return narrow(_np.take_along_axis(x, idx, axis=-1)), narrow(idx)
