"""Mirrors /root/reference/src/slope_interval.py: SlopeIntervalImplicitFunction (:15-50) -- the bounder that carries a
primal value, slope centres and slope widths (one per box vector) through the net (rules of
src/slope_interval_layers.py: dense :11-33, relu :35-58, elu :60-83, squeeze_last, spatial_transformation) and
classifies from primal -+ sum_v max(upper_v, -lower_v).  The propagation runs in the CUDA engine as a 7-row tile
[primal, centre x3, width x3] (csrc/niq_engine.cuh TileSlope3, csrc/niq_kernels.cuh k_classify_slope); boxes with up to
3 vectors.  The min_distance_to_zero* helpers of the reference (:52-163; no query of the hot path calls them) are not
provided."""
from dataclasses import dataclass

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

    def min_distance_to_zero(self, *a, **k):
        raise _niq.NiqError(_niq.NIQ_EUNSUPPORTED, "min_distance_to_zero is outside this backend (no hot-path query uses it)")

    min_distance_to_zero_in_direction = min_distance_to_zero
