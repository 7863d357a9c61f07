"""Mirrors /root/reference/src/sdf.py: WeakSDFImplicitFunction (:17-50) -- classify a box from ONE evaluation of the
function at the box centre and a Lipschitz bound: the sign can change inside the box only if
|f(centre)| - lipschitz * radius < 0, radius = sqrt(sum_v ||vec_v||^2).  The evaluation and the test run on the GPU
(niq_classify_*boxes with NIQ_MODE_SDF: k_eval_points on the centres + k_sdf_labels); the same function object drives
the tree / intersection / closest-point queries and cast_rays like any other bounder."""
from dataclasses import dataclass

import affine


@dataclass(frozen=True)
class SdfContext:
    """What the C ABI needs to know about this bounder (the reference keeps lipschitz_bound on the function object)."""
    lipschitz_bound: float = 1.
    mode: str = "sdf"


class WeakSDFImplicitFunction(affine.AffineImplicitFunction):
    """Same call surface as the reference class; bound_box / bound_general_box (ours) return
    (label, f - L*radius, f + L*radius, near_tie)."""

    def __init__(self, sdf_func, lipschitz_bound=1.):
        super().__init__(sdf_func, SdfContext(float(lipschitz_bound)))
        self.sdf_func = sdf_func
        self.lipschitz_bound = lipschitz_bound
