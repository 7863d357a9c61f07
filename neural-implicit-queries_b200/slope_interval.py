"""Mirrors /root/reference/src/slope_interval.py: SlopeIntervalImplicitFunction (:15-163) -- the bounder that carries a
primal value, slope centres and slope widths (one per box vector) through the net (rules of
src/slope_interval_layers.py: dense :11-33, relu :35-58, elu :60-83, squeeze_last, spatial_transformation) and
classifies from primal -+ sum_v max(upper_v, -lower_v).  The propagation runs in the CUDA engine as a 7-row tile
[primal, centre x3, width x3] (csrc/niq_engine.cuh TileSlope3, csrc/niq_kernels.cuh k_classify_slope); boxes with up to
3 vectors.  min_distance_to_zero / min_distance_to_zero_in_direction (:52-163; no query of the hot path calls them) take
the propagated form from the same kernel (niq_slope_forward) and finish with the reference's closed-form arithmetic in
float32 on the host; they accept one box like the reference or a leading batch dimension."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

import _niq
import affine


@dataclass(frozen=True)
class SlopeIntervalContext:
    mode: str = "slope_interval"


class SlopeIntervalImplicitFunction(affine.AffineImplicitFunction):
    """Same call surface as the reference class for classification; bound_box / bound_general_box (ours) return
    (label, may_lower, may_upper, near_tie)."""

    def __init__(self, slope_interval_func):
        super().__init__(slope_interval_func, SlopeIntervalContext())
        self.style = "classify-and-distance"
        self.slope_interval_func = slope_interval_func

    # ---- the propagated form: primal (n,), slope_lower (n,v), slope_upper (n,v)  (slope_bounds, src/slope_interval.py:196-199) ----
    def _slope_form(self, params, center, vecs, ctx=None):
        ctx = ctx or _niq.default_context()
        center = np.ascontiguousarray(center, np.float32).reshape(-1, 3)
        vecs = np.ascontiguousarray(vecs, np.float32)
        v = vecs.shape[-2]
        vecs = vecs.reshape(-1, v, 3)
        n = center.shape[0]
        raw = np.empty((n, 7), np.float32)
        m = ctx.mlp(params)
        _niq.check(_niq.lib().niq_slope_forward(ctx.handle, m.handle, C.c_int64(n), _niq.ptr(center), _niq.ptr(vecs), C.c_int32(v),
                                                _niq.ptr(raw), None, C.c_int(_niq.MEM_HOST)))
        sc, sw = raw[:, 1:1 + v], raw[:, 4:4 + v]
        return raw[:, 0].copy(), (sc - sw).astype(np.float32), (sc + sw).astype(np.float32)

    def min_distance_to_zero(self, params, box_center, box_axis_vec, return_source_value=False, ctx=None):
        """src/slope_interval.py:52-78: for an axis-aligned box (centre, half extents) a distance from the centre within which
        f keeps its sign: |f(centre)| / (sum_axes max|slope| in world units), at most the smallest half extent."""
        f32 = np.float32
        c = np.ascontiguousarray(box_center, f32)
        a = np.ascontiguousarray(box_axis_vec, f32)
        lead = c.shape[:-1]
        c2, a2 = c.reshape(-1, 3), a.reshape(-1, 3)
        lower, upper = c2 - a2, c2 + a2                       # coordinates_in_box(lower, upper), :177-181
        center = (f32(0.5) * (lower + upper)).astype(f32)
        half = (upper - center).astype(f32)
        vecs = np.zeros((c2.shape[0], 3, 3), f32)
        for i in range(3):
            vecs[:, i, i] = half[:, i]
        raw_primal, sl, su = self._slope_form(params, center, vecs, ctx)
        with np.errstate(divide="ignore", invalid="ignore"):
            primal = np.where(raw_primal >= 0, raw_primal, -raw_primal)
            dec = np.maximum(np.abs(sl), np.abs(su))
            vec_len = np.abs(a2)
            min_len = vec_len.min(axis=-1)
            dec = (dec / vec_len).astype(f32)
            dec = np.maximum(dec, f32(0.))
            axis_decrease = ((dec[:, 0] + dec[:, 1]) + dec[:, 2]).astype(f32)
            dist = np.minimum((primal / axis_decrease).astype(f32), min_len)
            dist = np.where(dist == 0, f32(0.), dist).astype(f32)
        raw_primal, dist = raw_primal.reshape(lead), dist.reshape(lead)
        if not lead:
            raw_primal, dist = f32(raw_primal), f32(dist)
        return (raw_primal, dist) if return_source_value else dist

    def min_distance_to_zero_in_direction(self, params, source_point, bound_vec, source_range=None, return_source_value=False,
                                          ctx=None):
        """src/slope_interval.py:81-163: how far one can move from `source_point` (or from anywhere in the box
        source_point + source_range) along `bound_vec` before f can reach zero, from the slope bound along that direction over
        the swept region; at most |bound_vec|."""
        f32 = np.float32
        src = np.ascontiguousarray(source_point, f32)
        bv = np.ascontiguousarray(bound_vec, f32)
        lead = src.shape[:-1]
        s2, b2 = src.reshape(-1, 3), bv.reshape(-1, 3)
        n = s2.shape[0]
        fwd = (b2 * f32(0.5)).astype(f32)
        center = (s2 + fwd).astype(f32)
        rng = None
        if source_range is not None:
            rng = np.ascontiguousarray(source_range, f32)
            rng = rng.reshape(n, rng.shape[-2], 3)
        vecs = fwd[:, None, :] if rng is None else np.concatenate((fwd[:, None, :], rng), axis=1)
        _, sl, su = self._slope_form(params, center, vecs, ctx)
        blen = np.sqrt(((b2[:, 0] * b2[:, 0] + b2[:, 1] * b2[:, 1]) + b2[:, 2] * b2[:, 2]).astype(f32)).astype(f32)
        shape = (lambda x: f32(x.reshape(lead)) if not lead else x.reshape(lead))
        with np.errstate(divide="ignore", invalid="ignore"):
            if rng is not None:
                sp, ssl, ssu = self._slope_form(params, s2, rng, ctx)
                prad = np.maximum(ssu, -ssl).sum(axis=1, dtype=f32)        # primal_may_contain_bounds, :201-206
                s_lo, s_up = (sp - prad).astype(f32), (sp + prad).astype(f32)
                is_pos = s_lo >= 0
                val = np.where(is_pos, s_lo, -s_up)
                slope = np.where(is_pos, sl[:, 0], -su[:, 0])
                slope = (f32(2.) * slope / blen).astype(f32)
                dec = np.maximum(-slope, f32(0.))
                dist = np.minimum((val / dec).astype(f32), blen)
                dist = np.where((s_lo <= 0) & (s_up >= 0), f32(0.), dist).astype(f32)
                return (shape(s_lo), shape(s_up), shape(dist)) if return_source_value else shape(dist)
            sval = np.asarray(self(params, s2), f32).reshape(-1)
            is_pos = sval >= 0
            val = np.abs(sval)
            slope = np.where(is_pos, sl[:, 0], -su[:, 0])
            slope = (f32(2.) * slope / blen).astype(f32)
            dec = np.maximum(-slope, f32(0.))
            dist = np.minimum((val / dec).astype(f32), blen)
            dist = np.where(sval == 0, f32(0.), dist).astype(f32)
        return (shape(sval), shape(dist)) if return_source_value else shape(dist)
