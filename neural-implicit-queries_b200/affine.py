"""Mirrors the public surface of /root/reference/src/affine.py: AffineContext (:62-76) and
AffineImplicitFunction (:17-55).  The arithmetic (coordinates_in_general_box :109-117, the layer rules
of src/affine_layers.py, apply_linear_approx :164-193, truncate_affine :127-162, may_contain_bounds
:119-125) runs in the CUDA kernels of csrc/ (niq_engine.cuh: interval / affine_fixed;
niq_grow.cuh: affine_all / affine_truncate)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

import _niq
import implicit_function
from implicit_function import SIGN_NEGATIVE, SIGN_POSITIVE, SIGN_UNKNOWN  # noqa: F401


@dataclass(frozen=True)
class AffineContext:
    mode: str = "affine_fixed"
    truncate_count: int = -777
    truncate_policy: str = "absolute"
    affine_domain_terms: int = 0
    n_append: int = 0

    def __post_init__(self):
        if self.mode not in ["interval", "affine_fixed", "affine_truncate", "affine_append", "affine_all"]:
            raise ValueError("invalid mode")
        if self.mode == "affine_truncate":
            if self.truncate_count is None:
                raise ValueError("must specify truncate count")


class AffineImplicitFunction(implicit_function.ImplicitFunction):
    def __init__(self, affine_func, ctx):
        super().__init__("classify-only")
        self.affine_func = affine_func
        self.ctx = ctx
        self.mode_dict = {"ctx": self.ctx}

    def __call__(self, params, x):
        """f(x): x (3,) -> scalar, or (..., 3) -> (...) (replaces vmap)."""
        import mlp
        return mlp.eval_points(params, x)

    def classify_general_box(self, params, box_center, box_vecs, offset=0.):
        lab = self.bound_general_box(params, box_center, box_vecs, offset)[0]
        return lab

    def bound_general_box(self, params, box_center, box_vecs, offset=0., ctx=None):
        """-> (label i32, lower f32, upper f32, near_tie bool), each of the leading batch shape.
        box_center (..., 3), box_vecs (..., v, 3).  (The reference exposes only the label; the bounds of
        src/affine.py:119-125 and the near-tie flag are extras the parity tests use.)"""
        ctx = ctx or _niq.default_context()
        center = np.ascontiguousarray(box_center, np.float32)
        vecs = np.ascontiguousarray(box_vecs, np.float32)
        d = center.shape[-1]
        assert d == 3, "bad box_center shape"
        v = vecs.shape[-2]
        assert vecs.shape == center.shape[:-1] + (v, d), "bad box_vecs shape"
        lead = center.shape[:-1]
        c2 = center.reshape(-1, 3)
        v2 = vecs.reshape(-1, v, 3)
        n = c2.shape[0]
        lab = np.empty(n, np.int32)
        lo = np.empty(n, np.float32)
        up = np.empty(n, np.float32)
        tie = np.empty(n, np.uint8)
        cfg = _niq.mode_cfg(self.ctx)
        m = ctx.mlp(params)
        _niq.check(_niq.lib().niq_classify_general_boxes(
            ctx.handle, m.handle, C.byref(cfg), C.c_int64(n), _niq.ptr(c2), _niq.ptr(v2), C.c_int32(v),
            C.c_float(offset), _niq.ptr(lab), _niq.ptr(lo), _niq.ptr(up), _niq.ptr(tie), C.c_int(_niq.MEM_HOST)))
        return lab.reshape(lead), lo.reshape(lead), up.reshape(lead), tie.reshape(lead).astype(bool)

    def bound_box(self, params, box_lower, box_upper, offset=0., ctx=None):
        """Axis-aligned variant through niq_classify_boxes (lo/hi (..., 3))."""
        ctx = ctx or _niq.default_context()
        lo_in = np.ascontiguousarray(box_lower, np.float32)
        hi_in = np.ascontiguousarray(box_upper, np.float32)
        assert lo_in.shape == hi_in.shape and lo_in.shape[-1] == 3
        lead = lo_in.shape[:-1]
        n = int(np.prod(lead, dtype=np.int64)) if lead else 1
        lab = np.empty(n, np.int32)
        lo = np.empty(n, np.float32)
        up = np.empty(n, np.float32)
        tie = np.empty(n, np.uint8)
        cfg = _niq.mode_cfg(self.ctx)
        m = ctx.mlp(params)
        _niq.check(_niq.lib().niq_classify_boxes(
            ctx.handle, m.handle, C.byref(cfg), C.c_int64(n), _niq.ptr(lo_in.reshape(-1, 3)),
            _niq.ptr(hi_in.reshape(-1, 3)), C.c_float(offset), _niq.ptr(lab), _niq.ptr(lo), _niq.ptr(up),
            _niq.ptr(tie), C.c_int(_niq.MEM_HOST)))
        return lab.reshape(lead), lo.reshape(lead), up.reshape(lead), tie.reshape(lead).astype(bool)

    def classify_box(self, params, box_lower, box_upper, offset=0.):
        return self.bound_box(params, box_lower, box_upper, offset)[0]
