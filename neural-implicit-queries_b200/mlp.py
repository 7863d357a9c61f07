"""MLP-as-dict format: the reference's `mlp` module surface for the range-analysis hot path.

Mirrors /root/reference/src/mlp.py (names, signatures, key grammar, error types):
  build_spec :14-24, get_op_data :117-131, n_ops :134-144, prepend_op :149-167, load :173-185,
  save :187-196, op constructors dense :218-250, relu :283, elu :289, sin :296, pow2_frequency_encode :304-315, squeeze_last :328,
  spatial_transformation :335-340, quick_mlp_spec :73-94, func_from_spec :96-113.
Arrays are NumPy float32 (the reference holds jnp arrays; NumPy arrays are accepted there too).
Evaluation does not happen here: `func_from_spec` returns a callable that packs `params` into a
device handle (content-hashed, see _niq.Context.mlp) and runs the CUDA point-evaluation kernel.
"""
import numpy as np

import _niq


def build_spec(mlp_op_list):
    out_params = {}
    for i_op, op in enumerate(mlp_op_list):
        for key, val in op.items():
            out_params[f"{i_op:04d}." + key] = val
    return out_params


def get_op_data(params, i_op):
    i_op_str = f"{i_op:04d}"
    name = ""
    args = {}
    for key in params:
        if key.startswith(i_op_str):
            tokens = key.split(".")
            name = tokens[1]
            if len(tokens) > 2:
                args[tokens[2]] = params[key]
    if name == "":
        raise ValueError(f"didn't find op {i_op}")
    return name, args


def n_ops(params):
    n = 0
    for key in params:
        vals = key.split(".")
        try:
            i_op = int(vals[0])
        except ValueError:
            raise ValueError(f"Could not parse out key {key}. Is this a valid mlp spec? Did you make a mistake "
                             "passing params dictionaries around?")
        n = max(n, i_op + 1)
    return n


def prepend_op(params, op):
    new_params = {}
    for key in params:
        vals = key.split(".")
        vals[0] = f"{int(vals[0]) + 1:04d}"
        new_params[".".join(vals)] = params[key]
    for key, val in op.items():
        new_params[f"{0:04d}." + key] = val
    return new_params


def load(filename):
    out_params = {}
    param_count = 0
    with np.load(filename) as data:
        for key, val in data.items():
            if isinstance(val, np.ndarray):
                param_count += val.size
                val = np.array(val)
            out_params[key] = val
    print(f"Loaded MLP with {param_count} params")
    return out_params


def save(filename, params):
    np.savez(filename, **{k: np.asarray(v) for k, v in params.items()})


# ---- op constructors (same dict fragments as the reference) ------------------------------------

def dense(in_dim, out_dim, with_bias=True, A=None, b=None):
    if not with_bias and b is not None:
        raise ValueError("cannot specifify 'b' and 'with_bias=False'")
    if A is None:
        A = (in_dim, out_dim)                       # initialised later (initialize_params)
    else:
        A = np.asarray(A, np.float32)
        if A.shape != (in_dim, out_dim):
            raise ValueError(f"A should have shape ({in_dim},{out_dim}). Has shape {A.shape}.")
    if b is None and with_bias:
        b = (out_dim,)
    elif b is not None:
        b = np.asarray(b, np.float32)
        if b.shape != (out_dim,):
            raise ValueError(f"b should have shape ({out_dim}). Has shape {b.shape}.")
    sub = {"dense.A": A}
    if with_bias:
        sub["dense.b"] = b
    return sub


def relu():
    return {"relu._": np.zeros((0,), np.float32)}


def elu():
    return {"elu._": np.zeros((0,), np.float32)}


def sin():
    return {"sin._": np.zeros((0,), np.float32)}


def tanh():
    """OURS, parity unpinned: the reference's README (:24) names TanH MLPs but src/mlp.py registers no tanh op.
    Key "<i>.tanh._", point rule tanh(x), bound rules csrc/niq_engine.cuh tanh_lin."""
    return {"tanh._": np.zeros((0,), np.float32)}


def pow2_frequency_encode(count_pow2, start_pow=0, with_shift=True):
    """src/mlp.py:304-315: positional encoding coefficients 2^k * pi (and the pi shift that turns every second sin into a
    cos); followed by sin() in the reference's fitting script (src/main_fit_implicit.py:113-115)."""
    pows = np.power(np.float32(2.), np.arange(start_pow, start_pow + count_pow2, dtype=np.float32)).astype(np.float32)
    coefs = (pows * np.float32(np.pi)).astype(np.float32)
    if with_shift:
        coefs = np.repeat(coefs, 2)
        shift = np.zeros_like(coefs)
        shift[1::2] = np.float32(np.pi)
        return {"pow2_frequency_encode.coefs": coefs, "pow2_frequency_encode.shift": shift}
    return {"pow2_frequency_encode.coefs": coefs}


def squeeze_last():
    return {"squeeze_last._": np.zeros((0,), np.float32)}


def spatial_transformation():
    return {"spatial_transformation.R": np.eye(3, dtype=np.float32),
            "spatial_transformation.t": np.zeros(3, dtype=np.float32)}


def quick_mlp_spec(layer_sizes, activation):
    """src/mlp.py:73-94 (relu / elu as there; 'tanh' is ours: see tanh())."""
    spec_list = []
    for i in range(len(layer_sizes) - 1):
        spec_list.append(dense(layer_sizes[i], layer_sizes[i + 1]))
        if i + 2 != len(layer_sizes):
            if activation == "relu":
                spec_list.append(relu())
            elif activation == "elu":
                spec_list.append(elu())
            elif activation == "tanh":
                spec_list.append(tanh())
            else:
                raise ValueError("unrecognized activation")
    spec_list.append(squeeze_last())
    return spec_list


def initialize_params(params, rngkey):
    """src/mlp.py:26-53 + initialize_dense :260-277: glorot-normal A, b ~ N(0, 1e-2^2).
    `rngkey` is an int seed or a numpy Generator (the JAX PRNG does not exist here: same
    distributions, not the same bits)."""
    if rngkey is None:
        raise ValueError("to initialize model weights, must pass an RNG key")
    rng = rngkey if isinstance(rngkey, np.random.Generator) else np.random.default_rng(rngkey)
    out = {}
    for i_op in range(n_ops(params)):
        name, args = get_op_data(params, i_op)
        for a, val in args.items():
            if name == "dense" and isinstance(val, tuple):
                if a == "A":
                    std = np.sqrt(2.0 / (val[0] + val[1]))
                    val = (rng.standard_normal(val) * std).astype(np.float32)
                else:
                    val = (rng.standard_normal(val) * 1e-2).astype(np.float32)
            out[f"{i_op:04d}.{name}.{a}"] = val
    return out


# ---- evaluation entry point ---------------------------------------------------------------------

def func_from_spec(mode="default"):
    """src/mlp.py:96-113.  mode 'default' -> point evaluation f(params, x) on the GPU; x (3,) or (...,3).
    mode 'affine' returns a marker consumed by affine.AffineImplicitFunction (the affine interpreter
    itself is the CUDA bound-propagation kernel; it is not exposed as a Python-level function)."""
    if mode == "default":
        def eval_spec(params, x, mode_dict=None):
            return eval_points(params, x)
        return eval_spec
    if mode == "affine":
        def affine_marker(params, x, mode_dict=None):
            raise RuntimeError("the affine interpreter runs inside the CUDA kernels; use "
                               "AffineImplicitFunction.classify_box / classify_general_box")
        affine_marker.niq_mode = "affine"
        return affine_marker
    raise _niq.NiqError(_niq.NIQ_EUNSUPPORTED, f"func_from_spec mode '{mode}' is outside this backend")


def eval_points(params, x, return_scale=False, ctx=None):
    """f(x) for x (..., 3) -> (...) float32 through niq_eval_points."""
    import ctypes as C
    ctx = ctx or _niq.default_context()
    x = np.ascontiguousarray(x, np.float32)
    if x.shape[-1] != 3:
        raise ValueError("points must have shape (..., 3)")
    lead = x.shape[:-1]
    flat = x.reshape(-1, 3)
    n = flat.shape[0]
    f = np.empty(n, np.float32)
    s = np.empty(n, np.float32) if return_scale else None
    m = ctx.mlp(params)
    _niq.check(_niq.lib().niq_eval_points(ctx.handle, m.handle, C.c_int64(n), _niq.ptr(flat), _niq.ptr(f),
                                          _niq.ptr(s), C.c_int(_niq.MEM_HOST)))
    if return_scale:
        return f.reshape(lead), s.reshape(lead)
    return f.reshape(lead)
