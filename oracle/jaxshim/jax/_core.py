"""Array type and helpers for the NumPy-backed jax stand-in (TEST INFRASTRUCTURE ONLY)."""
import numpy as np

_NARROW = {np.dtype(np.float64): np.float32, np.dtype(np.int64): np.int32,
           np.dtype(np.uint64): np.uint32, np.dtype(np.complex128): np.complex64}


def narrow(x):
    """float64/int64 -> float32/int32 (x64 disabled); wraps ndarrays as Arr, leaves other objects."""
    if isinstance(x, (tuple, list)):
        return type(x)(narrow(v) for v in x)
    if isinstance(x, np.ndarray):
        t = _NARROW.get(x.dtype)
        if t is not None:
            x = x.astype(t)
        return x.view(Arr)
    if isinstance(x, np.generic):
        t = _NARROW.get(x.dtype)
        return t(x) if t is not None else x
    return x


def to_arr(x):
    """Python scalars / lists / arrays -> Arr with JAX default dtypes."""
    if isinstance(x, Arr):
        return x
    if x is None:
        return None
    return narrow(np.asarray(x))


def _oob_fill(dtype):
    if np.issubdtype(dtype, np.inexact):
        return np.nan
    if np.issubdtype(dtype, np.bool_):
        return True
    if np.issubdtype(dtype, np.signedinteger):
        return np.iinfo(dtype).min
    return np.iinfo(dtype).max


def _norm_index(arr, idx):
    """Split an index into numpy-ready form and an in-bounds mask for its integer-array parts.

    Supports: integer array / tuple-of-arrays-as-one-axis (jnp.nonzero output), ints, slices, Ellipsis,
    and several integer arrays (broadcast together).  Negative indices wrap first (NumPy/JAX rule)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    # expand ellipsis to know the axis of each entry
    n_spec = sum(1 for i in idx if i is not Ellipsis and i is not None)
    out = []
    axis = 0
    ok = None
    for i in idx:
        if i is Ellipsis:
            n_fill = arr.ndim - n_spec
            out.extend([slice(None)] * n_fill)
            axis += n_fill
            continue
        if i is None:
            out.append(None)
            continue
        if isinstance(i, (tuple, list)) or (isinstance(i, np.ndarray) and i.dtype != np.bool_) \
                or isinstance(i, np.integer):
            a = np.asarray(i)
            if a.dtype == np.bool_:
                out.append(a)
                axis += a.ndim
                continue
            a = a.astype(np.int64)
            n = arr.shape[axis]
            a = np.where(a < 0, a + n, a)
            good = (a >= 0) & (a < n)
            ok = good if ok is None else (ok & good)
            out.append(a)
            axis += 1
        else:
            out.append(i)
            axis += 1
    return tuple(out), ok


class _AtIndexed:
    def __init__(self, arr, idx):
        self.arr = arr
        self.idx = idx

    def _scatter(self, vals, mode, op):
        out = np.array(self.arr, copy=True)
        idx, ok = _norm_index(out, self.idx)
        vals = np.asarray(vals)
        if ok is None or ok.all():
            if op == "set":
                out[idx] = vals
            elif op == "min":
                np.minimum.at(out, idx, vals)
            elif op == "max":
                np.maximum.at(out, idx, vals)
            elif op == "add":
                np.add.at(out, idx, vals)
            return narrow(out)
        # out-of-bounds updates are dropped (JAX scatter default and mode='drop')
        int_pos = [k for k, i in enumerate(idx) if isinstance(i, np.ndarray) and i.dtype != np.bool_]
        assert len({idx[k].shape for k in int_pos}) == 1 or all(idx[k].ndim <= 1 for k in int_pos)
        okb = np.broadcast_to(ok, np.broadcast_shapes(*[idx[k].shape for k in int_pos]))
        sel = np.nonzero(okb.reshape(-1))[0]
        new_idx = list(idx)
        for k in int_pos:
            new_idx[k] = np.broadcast_to(idx[k], okb.shape).reshape(-1)[sel]
        # result shape of the indexed view, via a dry gather with clipped indices
        clipped = list(idx)
        for k in int_pos:
            clipped[k] = np.clip(idx[k], 0, out.shape[self._axis_of(idx, k)] - 1)
        full_shape = out[tuple(clipped)].shape
        vals_b = np.broadcast_to(vals, full_shape)
        # the integer-array axes land at the front of the result iff they are adjacent & first; we only
        # support the layouts the reference uses: index arrays first, trailing slices.
        assert int_pos == list(range(len(int_pos))), "shim: unsupported scatter layout"
        lead = okb.shape
        vals_sel = vals_b.reshape((-1,) + full_shape[len(lead):])[sel]
        if op == "set":
            out[tuple(new_idx)] = vals_sel
        elif op == "min":
            np.minimum.at(out, tuple(new_idx), vals_sel)
        elif op == "max":
            np.maximum.at(out, tuple(new_idx), vals_sel)
        elif op == "add":
            np.add.at(out, tuple(new_idx), vals_sel)
        return narrow(out)

    @staticmethod
    def _axis_of(idx, k):
        ax = 0
        for j in range(k):
            if idx[j] is None:
                continue
            ax += 1
        return ax

    def set(self, vals, mode=None, indices_are_sorted=False, unique_indices=False):
        return self._scatter(vals, mode, "set")

    def min(self, vals, mode=None, **kw):
        return self._scatter(vals, mode, "min")

    def max(self, vals, mode=None, **kw):
        return self._scatter(vals, mode, "max")

    def add(self, vals, mode=None, **kw):
        return self._scatter(vals, mode, "add")

    def get(self, mode=None, fill_value=None, indices_are_sorted=False, unique_indices=False):
        a = np.asarray(self.arr)
        idx, ok = _norm_index(a, self.idx)
        if ok is None or ok.all():
            return narrow(a[idx])
        int_pos = [k for k, i in enumerate(idx) if isinstance(i, np.ndarray) and i.dtype != np.bool_]
        clipped = list(idx)
        for k in int_pos:
            clipped[k] = np.clip(idx[k], 0, a.shape[self._axis_of(idx, k)] - 1)
        res = np.array(a[tuple(clipped)], copy=True)
        if mode in ("fill", "drop"):
            fv = _oob_fill(a.dtype) if fill_value is None else fill_value
            assert int_pos == list(range(len(int_pos))), "shim: unsupported gather layout"
            okb = np.broadcast_to(ok, np.broadcast_shapes(*[idx[k].shape for k in int_pos]))
            res[~okb] = np.asarray(fv).astype(a.dtype)
        # mode None / 'clip' / 'promise_in_bounds': clamped gather (JAX default for gather)
        return narrow(res)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndexed(self.arr, idx)


class Arr(np.ndarray):
    """ndarray with `.at[]`, narrowing of every ufunc result, and a no-op block_until_ready()."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(np.asarray(i) if isinstance(i, Arr) else i for i in inputs)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, Arr) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        if res is NotImplemented:
            return NotImplemented
        if isinstance(res, tuple):
            return tuple(narrow(r) for r in res)
        return narrow(res)

    def __getitem__(self, idx):
        # jnp-style gather: out-of-range integer-array indices clamp instead of raising
        if isinstance(idx, tuple) and any(isinstance(i, tuple) for i in idx):
            idx = tuple(np.asarray(i) if isinstance(i, tuple) else i for i in idx)
        try:
            res = np.ndarray.__getitem__(self, idx)
        except IndexError:
            a = np.asarray(self)
            nidx, ok = _norm_index(a, idx)
            int_pos = [k for k, i in enumerate(nidx) if isinstance(i, np.ndarray) and i.dtype != np.bool_]
            clipped = list(nidx)
            for k in int_pos:
                clipped[k] = np.clip(nidx[k], 0, a.shape[_AtIndexed._axis_of(nidx, k)] - 1)
            res = a[tuple(clipped)]
        return narrow(res) if isinstance(res, (np.ndarray, np.generic)) else res
