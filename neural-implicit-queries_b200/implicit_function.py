"""Mirrors /root/reference/src/implicit_function.py:11-37 (SIGN_* codes, ImplicitFunction base)."""
import numpy as np

SIGN_UNKNOWN = 0    # could be anything
SIGN_POSITIVE = 1   # definitely positive throughout
SIGN_NEGATIVE = 2   # definitely negative throughout


class ImplicitFunction:
    def __init__(self, style):
        if style not in ["classify-only", "classify-and-distance"]:
            raise ValueError("unrecognized style")
        self.style = style

    def __call__(self, params, x):
        raise RuntimeError("ImplicitFunction does not implement a __call__() operator. Subclasses must provide an "
                           "implementation if is to be used.")

    def classify_box(self, params, box_lower, box_upper, offset=0.):
        """src/implicit_function.py:28-37.  Leading batch dims replace the reference's vmap."""
        box_lower = np.asarray(box_lower, np.float32)
        box_upper = np.asarray(box_upper, np.float32)
        center = (np.float32(0.5) * (box_lower + box_upper)).astype(np.float32)
        pos_vec = (box_upper - center).astype(np.float32)
        vecs = np.zeros(pos_vec.shape + (pos_vec.shape[-1],), np.float32)
        idx = np.arange(pos_vec.shape[-1])
        vecs[..., idx, idx] = pos_vec
        return self.classify_general_box(params, center, vecs, offset=offset)

    def classify_general_box(self, params, box_center, box_vecs, offset=0.):
        raise RuntimeError("ImplicitFunction does not implement classify_general_box(). Subclasses must provide an "
                           "implementation if is to be used.")
