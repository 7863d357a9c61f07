"""Mirrors /root/reference/src/implicit_mlp_utils.py:12-64 for the affine-family modes (the hot path + affine_append)."""
import _niq
import affine
import mlp


def generate_implicit_from_file(input_path, mode, **kwargs):
    if input_path.endswith(".npz"):
        params = mlp.load(input_path)
    else:
        raise ValueError("unrecognized filetype")
    return generate_implicit_from_params(params, mode, **kwargs), params


def generate_implicit_from_params(params, mode, **kwargs):
    """The mode -> ImplicitFunction half of the factory (ours: lets callers skip the file)."""
    if mode == "interval":
        ctx = affine.AffineContext("interval")
    elif mode == "affine_fixed":
        ctx = affine.AffineContext("affine_fixed")
    elif mode == "affine_truncate":
        ctx = affine.AffineContext("affine_truncate", truncate_count=kwargs["affine_n_truncate"],
                                   truncate_policy=kwargs["affine_truncate_policy"])
    elif mode == "affine_all":
        ctx = affine.AffineContext("affine_all")
    elif mode == "affine_append":
        ctx = affine.AffineContext("affine_append", n_append=kwargs["affine_n_append"])
    elif mode == "sdf":
        import sdf
        return sdf.WeakSDFImplicitFunction(mlp.func_from_spec(mode="default"), lipschitz_bound=kwargs.get("sdf_lipschitz", 1.))
    elif mode == "slope_interval":
        import slope_interval
        return slope_interval.SlopeIntervalImplicitFunction(mlp.func_from_spec(mode="default"))
    else:
        raise RuntimeError("unrecognized mode")
    return affine.AffineImplicitFunction(mlp.func_from_spec(mode="affine"), ctx)
