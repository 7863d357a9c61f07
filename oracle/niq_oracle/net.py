"""CPU oracle, part 1: MLP dict format, point evaluation and affine / interval bound propagation.

TEST INFRASTRUCTURE ONLY.  This is a float32 NumPy restatement of the reference's range-analysis
core, with an explicit leading batch axis in place of `jax.vmap`.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product never does.

Pinning: the reference ships no tests or golden vectors.  This restatement is pinned against the
UNMODIFIED reference sources executed on a NumPy-backed `jax` stand-in (oracle/jaxshim, run by
oracle/tools/gen_golden.py in the build container) -- see tests/golden/README.md.  Real JAX/XLA is not
installable here, so XLA's own reduction order is not reproduced ("parity pinned to the reference's
Python code, not to XLA bits").

Reference files followed (all under /root/reference/src):
  mlp.py:96-144 (op-list interpreter, key grammar), :149-167 (prepend_op), :173-185 (load),
  mlp.py:253-347 (point-evaluation rules), affine.py:85-193, affine_layers.py:11-97,164-179,
  implicit_function.py:28-37.
"""
from dataclasses import dataclass

import numpy as np

F32 = np.float32

SIGN_UNKNOWN = 0   # implicit_function.py:11-13
SIGN_POSITIVE = 1
SIGN_NEGATIVE = 2

MODES = ("interval", "affine_fixed", "affine_truncate", "affine_all")


class precision:
    """Context manager (tests only): evaluate this module's functions in another float type, e.g.
    `with net.precision(np.float64): ...` gives the exact-arithmetic yardstick of the same formulas."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global F32
        self.saved = F32
        F32 = self.dtype
        return self

    def __exit__(self, *exc):
        global F32
        F32 = self.saved
        return False


# ----------------------------------------------------------------------------------------------
# params dict format  (mlp.py:14-24, 117-167, 173-185)
# ----------------------------------------------------------------------------------------------

def load_npz(path):
    """mlp.py:173-185 -- npz -> {key: float32 array}; no conversion beyond np.asarray."""
    out = {}
    with np.load(path) as data:
        for key in data.files:
            out[key] = np.asarray(data[key])
    return out


def n_ops(params):
    """mlp.py:134-144"""
    n = 0
    for key in params:
        head = key.split(".")[0]
        try:
            i_op = int(head)
        except ValueError:
            raise ValueError(f"Could not parse out key {key}. Is this a valid mlp spec?")
        n = max(n, i_op + 1)
    return n


def op_list(params):
    """Ordered [(name, {arg: array})] following mlp.py:117-131 (prefix match on the 4-digit index,
    op name = token 1, arg = token 2, the '_' placeholder dropped as in :108-109)."""
    ops = []
    for i_op in range(n_ops(params)):
        prefix = f"{i_op:04d}"
        name, args = "", {}
        for key in params:
            if key.startswith(prefix):
                tok = key.split(".")
                name = tok[1]
                if len(tok) > 2:
                    args[tok[2]] = params[key]
        if name == "":
            raise ValueError(f"didn't find op {i_op}")
        args.pop("_", None)
        ops.append((name, args))
    return ops


def prepend_op(params, op):
    """mlp.py:149-167"""
    new = {}
    for key, val in params.items():
        tok = key.split(".")
        tok[0] = f"{int(tok[0]) + 1:04d}"
        new[".".join(tok)] = val
    for key, val in op.items():
        new["0000." + key] = val
    return new


def spatial_transformation(R=None, t=None):
    """mlp.py:335-340 -- identity transform op dict (R, t optional overrides)."""
    return {"spatial_transformation.R": np.eye(3, dtype=F32) if R is None else np.asarray(R, F32),
            "spatial_transformation.t": np.zeros(3, dtype=F32) if t is None else np.asarray(t, F32)}


def _spatial_as_dense(R, t):
    """affine_layers.py:175-179 / mlp.py:342-346: A = inv(R) used as x @ A, b = inv(R) @ (-t)."""
    R = np.asarray(R, F32)
    t = np.asarray(t, F32)
    R_inv = np.linalg.inv(R).astype(F32)
    t_inv = (R_inv @ (-t)).astype(F32)
    return R_inv, t_inv


def random_mlp(layer_sizes, activation="relu", seed=0):
    """Synthetic MLP of the shape `mlp.quick_mlp_spec` builds (mlp.py:73-94), initialised with the
    distributions of `mlp.initialize_dense` (mlp.py:260-277): glorot-normal A, b ~ N(0, 1e-2^2).
    NumPy's generator replaces the JAX PRNG: same distribution, not the same bits."""
    rng = np.random.default_rng(seed)
    params = {}
    i_op = 0
    for i in range(len(layer_sizes) - 1):
        d_in, d_out = layer_sizes[i], layer_sizes[i + 1]
        std = np.sqrt(2.0 / (d_in + d_out))
        params[f"{i_op:04d}.dense.A"] = (rng.standard_normal((d_in, d_out)) * std).astype(F32)
        params[f"{i_op:04d}.dense.b"] = (rng.standard_normal((d_out,)) * 1e-2).astype(F32)
        i_op += 1
        if i + 2 != len(layer_sizes):
            params[f"{i_op:04d}.{activation}._"] = np.zeros((0,), F32)
            i_op += 1
    params[f"{i_op:04d}.squeeze_last._"] = np.zeros((0,), F32)
    return params


# ----------------------------------------------------------------------------------------------
# point evaluation (mlp.py 'default' rules)
# ----------------------------------------------------------------------------------------------

def _elu(x):
    # jax.nn.elu: where(x > 0, x, expm1(where(x > 0, 0, x)))
    safe = np.where(x > 0, F32(0), x)
    return np.where(x > 0, x, np.expm1(safe)).astype(F32)


PI = float(np.pi)


def _pow2_encode(x, coefs, shift, with_shift):
    """mlp.py:316-322 / affine_layers.py:147-152, batched over leading axes: (..., d) -> (..., d*c)."""
    coefs = np.asarray(coefs, F32)
    out = (np.asarray(x, F32)[..., :, None] * coefs).astype(F32)
    if with_shift and shift is not None:
        out = (out + np.asarray(shift, F32)).astype(F32)
    return out.reshape(out.shape[:-2] + (out.shape[-2] * out.shape[-1],))      # (explicit: interval mode has 0 aff rows)


def _sin_bound(lower, upper):
    """utils.py:209-229"""
    f_lower, f_upper = np.sin(lower), np.sin(upper)
    lower = (lower / F32(2. * PI)).astype(F32)
    upper = (upper / F32(2. * PI)).astype(F32)
    contains_min = np.ceil(lower - F32(.75)) < (upper - F32(.75))
    contains_max = np.ceil(lower - F32(.25)) < (upper - F32(.25))
    out_lower = np.where(contains_min, F32(-1.), np.minimum(f_lower, f_upper))
    out_upper = np.where(contains_max, F32(1.), np.maximum(f_lower, f_upper))
    return out_lower.astype(F32), out_upper.astype(F32)


def _cos_bound(lower, upper):
    """utils.py:230-231"""
    return _sin_bound((lower + F32(PI / 2)).astype(F32), (upper + F32(PI / 2)).astype(F32))


def eval_points(params, x):
    """f(x) for x of shape (N,3) -> (N,) float32.  mlp.py:99-111 with the 'default' rules
    (dense :253-258, relu :283-287, elu :289-293, sin :296-300, pow2_frequency_encode :316-322, squeeze_last :328-332,
    spatial :342-346)."""
    h = np.ascontiguousarray(x, dtype=F32)
    for name, args in op_list(params):
        if name == "dense":
            h = h @ np.asarray(args["A"], F32)
            if "b" in args and args["b"] is not None:
                h = h + np.asarray(args["b"], F32)
        elif name == "spatial_transformation":
            A, b = _spatial_as_dense(args["R"], args["t"])
            h = h @ A + b
        elif name == "relu":
            h = np.maximum(h, F32(0))
        elif name == "elu":
            h = _elu(h)
        elif name == "sin":
            h = np.sin(h)
        elif name == "tanh":                       # OURS (parity unpinned): the reference has no tanh op, see _tanh_coeffs
            h = np.tanh(h)
        elif name == "pow2_frequency_encode":
            h = _pow2_encode(h, args["coefs"], args.get("shift"), True)
        elif name == "squeeze_last":
            assert h.shape[-1] == 1
            h = h[..., 0]
        else:
            raise ValueError(f"oracle: unsupported op '{name}'")
        h = h.astype(F32, copy=False)
    return h


# ----------------------------------------------------------------------------------------------
# affine arithmetic (affine.py, affine_layers.py)
# ----------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class AffineContext:
    """affine.py:62-76"""
    mode: str = "affine_fixed"
    truncate_count: int = -777
    truncate_policy: str = "absolute"
    n_append: int = 0
    sdf_lipschitz: float = 1.0          # mode "sdf" (ours: the reference keeps it on sdf.WeakSDFImplicitFunction)

    def __post_init__(self):
        if self.mode not in ("interval", "affine_fixed", "affine_truncate", "affine_append", "affine_all", "sdf",
                             "slope_interval"):
            raise ValueError("invalid mode")
        if self.mode == "affine_truncate" and self.truncate_count is None:
            raise ValueError("must specify truncate count")


_RANK_TIE_RECORDER = None


def truncate_rank_near_tie(params, ctx, center, vecs, rel=1e-5, chunk=1024):
    """(N,) bool: some truncation of the box's propagation decided keep/drop between rows whose L1 norms agree to
    `rel` (relative).  Such boxes are excluded from strict parity like bound near-ties (counted, reported)."""
    global _RANK_TIE_RECORDER
    center = np.ascontiguousarray(center, F32)
    vecs = np.ascontiguousarray(vecs, F32)
    out = np.zeros(center.shape[0], bool)
    for s0 in range(0, center.shape[0], chunk):
        _RANK_TIE_RECORDER = {"rel": F32(rel), "flags": []}
        try:
            affine_forward(params, ctx, center[s0:s0 + chunk], vecs[s0:s0 + chunk])
            for f in _RANK_TIE_RECORDER["flags"]:
                out[s0:s0 + chunk] |= f
        finally:
            _RANK_TIE_RECORDER = None
    return out


def _radius(aff, err):
    """affine.py:85-91 -- sum_r |aff_r| + err, batched: aff (N,k,w), err (N,w)."""
    return (np.abs(aff).sum(axis=1, dtype=F32) + err).astype(F32)


def _truncate(ctx, base, aff, err):
    """affine.py:127-162, 'absolute' policy, stable descending sort by row L1 norm."""
    if ctx.mode != "affine_truncate":
        return base, aff, err
    n_keep = ctx.truncate_count
    if aff.shape[1] <= n_keep:
        return base, aff, err
    if ctx.truncate_policy != "absolute":
        # affine.py:146 divides (k,) by (w,): only shape-valid by accident, used by no caller
        raise RuntimeError("oracle: only the 'absolute' truncate policy is supported")
    mags = np.abs(aff).sum(axis=-1, dtype=F32)                       # (N,k)
    order = np.argsort(-mags, axis=-1, kind="stable")                # (N,k)
    if _RANK_TIE_RECORDER is not None:
        # diagnostic (ours): the keep/drop decision is a near-tie when the last kept and the first dropped row have
        # L1 norms within `rel` of each other -- float32 summation-order noise can then flip it
        srt = np.take_along_axis(mags, order, axis=-1)
        a, b = srt[:, n_keep - 1], srt[:, n_keep]
        # (rows that are exactly zero tie exactly on both sides: the stable sort settles them identically)
        _RANK_TIE_RECORDER["flags"].append((a > 0) & ((a - b) <= _RANK_TIE_RECORDER["rel"] * a))
    aff = np.take_along_axis(aff, order[:, :, None], axis=1)
    keep, drop = aff[:, :n_keep, :], aff[:, n_keep:, :]
    err = (err + np.abs(drop).sum(axis=1, dtype=F32)).astype(F32)
    return base, np.ascontiguousarray(keep), err


def _apply_linear_approx(ctx, base, aff, err, alpha, beta, delta):
    """affine.py:164-193 (modes interval / affine_fixed / affine_truncate / affine_all)."""
    base = (alpha * base + beta).astype(F32)
    aff = (alpha[:, None, :] * aff).astype(F32)
    delta = np.abs(delta)
    if ctx.mode in ("interval", "affine_fixed"):
        err = (alpha * err + delta).astype(F32)
    elif ctx.mode == "affine_append":
        err = (alpha * err).astype(F32)
        aff, err = _append_top_k(ctx, aff, err, delta)
    else:
        err = (alpha * err).astype(F32)
        n, w = delta.shape
        new_aff = np.zeros((n, w, w), F32)
        idx = np.arange(w)
        new_aff[:, idx, idx] = delta
        aff = np.concatenate((aff, new_aff), axis=1)
        base, aff, err = _truncate(ctx, base, aff, err)
    return base, aff, err


def _append_top_k(ctx, aff, err, delta):
    """affine.py:183-191 (mode affine_append): the n_append largest deltas (jax.lax.top_k: descending, lower index
    first among equals) become new single-entry rows; the rest is added to err -- as ONE scalar on every neuron,
    exactly like the reference's `err + (jnp.sum(delta) - jnp.sum(keep_vals))`."""
    n, w = delta.shape
    na = ctx.n_append
    order = np.argsort(-delta, axis=-1, kind="stable")[:, :na]                  # (N, na)
    keep = np.take_along_axis(delta, order, axis=-1).astype(F32)               # (N, na)
    new_aff = np.zeros((n, na, w), F32)
    new_aff[np.arange(n)[:, None], np.arange(na)[None, :], order] = keep
    rest = (delta.sum(axis=-1, dtype=F32) - keep.sum(axis=-1, dtype=F32)).astype(F32)
    return np.concatenate((aff, new_aff), axis=1), (err + rest[:, None]).astype(F32)


def _relu_coeffs(base, aff, err):
    """affine_layers.py:34-56 -> (alpha, beta, delta)"""
    rad = _radius(aff, err)
    lower, upper = base - rad, base + rad
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = (np.maximum(upper, F32(0)) - np.maximum(lower, F32(0))) / (upper - lower)
    alpha = np.where(lower >= 0, F32(1), alpha)
    alpha = np.where(upper < 0, F32(0), alpha)
    alpha = np.nan_to_num(alpha, nan=0.0).astype(F32)
    alpha = np.clip(alpha, F32(0), F32(1))
    beta = ((np.maximum(lower, F32(0)) - alpha * lower) / F32(2)).astype(F32)
    return alpha, beta, beta


def _relu_rule(ctx, base, aff, err):
    alpha, beta, delta = _relu_coeffs(base, aff, err)
    return _apply_linear_approx(ctx, base, aff, err, alpha, beta, delta)


def _elu_coeffs(base, aff, err):
    """affine_layers.py:59-97 -> (alpha, beta, delta)"""
    rad = _radius(aff, err)
    lower, upper = base - rad, base + rad
    lowerF, upperF = _elu(lower), _elu(upper)
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        lowerS = np.minimum(np.exp(lower), F32(1))
        upperS = np.minimum(np.exp(upper), F32(1))
        alpha = (upperF - lowerF) / (upper - lower)
        alpha = np.where(lower >= 0, F32(1), alpha)
        alpha = np.nan_to_num(alpha, nan=0.0).astype(F32)
        alpha = np.minimum(np.maximum(alpha, lowerS), upperS)       # jnp.clip(a_min, a_max)
        r_upper = lowerF - alpha * lower
        x_lower = np.minimum(np.maximum(np.log(alpha), lower), upper)
        r_lower = (alpha - F32(1)) - alpha * x_lower
    beta = F32(0.5) * (r_upper + r_lower)
    delta = F32(0.5) * np.abs(r_upper - r_lower)
    pos = lower >= 0
    alpha = np.where(pos, F32(1), alpha).astype(F32)
    beta = np.where(pos, F32(0), beta).astype(F32)
    delta = np.where(pos, F32(0), delta).astype(F32)
    return alpha, beta, delta


def _elu_rule(ctx, base, aff, err):
    alpha, beta, delta = _elu_coeffs(base, aff, err)
    return _apply_linear_approx(ctx, base, aff, err, alpha, beta, delta)


def _sin_coeffs(base, aff, err):
    """affine_layers.py:100-137 -> (alpha, beta, delta): a not-quite-Chebyshev linearisation of sin on [lower, upper]."""
    rad = _radius(aff, err)
    lower, upper = (base - rad).astype(F32), (base + rad).astype(F32)
    slope_lower, slope_upper = _cos_bound(lower, upper)
    alpha = (F32(0.5) * (slope_lower + slope_upper)).astype(F32)
    alpha = np.clip(alpha, F32(-1.), F32(1.))
    intA = np.arccos(alpha)
    intB = -intA
    two_pi = F32(2. * PI)

    def first(x):
        return (two_pi * np.ceil((lower + x) / two_pi) - x).astype(F32)

    def last(x):
        return (two_pi * np.floor((upper - x) / two_pi) + x).astype(F32)

    locs = [lower, upper, first(intA), last(intA), first(intB), last(intB)]
    locs = [np.minimum(np.maximum(x, lower), upper) for x in locs]
    vals = [(np.sin(x) - alpha * x).astype(F32) for x in locs]
    r_lower, r_upper = vals[0], vals[0]
    for v in vals[1:]:
        r_lower, r_upper = np.minimum(r_lower, v), np.maximum(r_upper, v)
    beta = (F32(0.5) * (r_upper + r_lower)).astype(F32)
    delta = (r_upper - beta).astype(F32)
    return alpha.astype(F32), beta, delta


def _sin_rule(ctx, base, aff, err):
    alpha, beta, delta = _sin_coeffs(base, aff, err)
    return _apply_linear_approx(ctx, base, aff, err, alpha, beta, delta)


def _tanh_coeffs(base, aff, err, dtype=F32):
    """PARITY UNPINNED -- the reference registers no tanh rule (README.md:24 names TanH, the code has only
    relu / elu / sin), so this restates nothing: it is the Chebyshev-style linearisation of the paper's construction, written
    here and in csrc/niq_engine.cuh tanh_lin from the same formulas.  On [l, u], w = u - l:
      alpha = the secant slope (tanh u - tanh l) / w (the minimax slope of a convex or concave piece), evaluated WITHOUT the
              cancelling difference through tanh u - tanh l = (1 - tanh u tanh l) tanh(u - l):
              alpha = (1 - tanh u tanh l) * g(w),  g(w) = tanh(w) / w  (1 - w^2/3 below w = 1e-3);
      r(x)  = tanh(x) - alpha x attains its extrema over [l, u] at l, u or where tanh'(x) = alpha, x* = +-atanh(sqrt(1-alpha));
      beta  = (r_max + r_min) / 2,  delta = (r_max - r_min) / 2, both formed from d(x) = r(x) - r(l) (see below).
    Sound for ANY alpha in [0, 1] (the residual bounds are exact for the alpha actually used), tight for the secant."""
    one = dtype(1)
    rad = _radius(aff, err) if dtype is F32 else np.abs(aff).sum(axis=1) + err
    lower, upper = (base - rad).astype(dtype), (base + rad).astype(dtype)
    tl, tu = np.tanh(lower).astype(dtype), np.tanh(upper).astype(dtype)
    width = (upper - lower).astype(dtype)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        g = np.where(width < dtype(1e-3), one - width * width / dtype(3), np.tanh(width) / width).astype(dtype)
        alpha = ((one - tu * tl) * g).astype(dtype)
        alpha = np.nan_to_num(alpha, nan=0.0).astype(dtype)
        alpha = np.clip(alpha, dtype(0), one)
        zero = alpha == 0
        # residual r(x) = tanh(x) - alpha x RELATIVE to its value at l: d(x) = (tanh x - tanh l) - alpha (x - l), with the tanh
        # difference again in product form -- the cancellation then costs ~1 ulp of w, not of |tanh l| (delta of a narrow
        # interval is O(w^3): formed from r(x) directly it would drown in the rounding noise of r, which the following
        # layers amplify like any other error term)
        def d_of(x):
            dx = (x - lower).astype(dtype)
            return ((one - np.tanh(x).astype(dtype) * tl) * np.tanh(dx).astype(dtype) - alpha * dx).astype(dtype)
        du = np.where(zero, tu - tl, d_of(upper)).astype(dtype)
        d_lo, d_hi = np.minimum(du, dtype(0)), np.maximum(du, dtype(0))
        xs = np.arctanh(np.sqrt(np.maximum(one - alpha, dtype(0)))).astype(dtype)           # alpha = 0 -> inf: skipped below
        for sgn in (one, dtype(-1)):
            x = np.minimum(np.maximum(sgn * xs, lower), upper).astype(dtype)                # clipped: an end point, already covered
            v = np.where(zero, dtype(0), d_of(x)).astype(dtype)
            d_lo, d_hi = np.minimum(d_lo, v), np.maximum(d_hi, v)
        rl = np.where(zero, tl, tl - alpha * lower).astype(dtype)                           # r(l)
        # narrow intervals (w < 1e-2): the true half-range is O(w^2 |f''|) and even d(x) cannot resolve it in float32; there the
        # residual is bounded by the secant-error theorem instead, |f(x) - secant(x)| <= max|f''| w^2 / 8, one-sided where f'' keeps
        # its sign (tanh is convex below 0, concave above).  f'' = -2 t (1 - t^2), |f''| <= 4 / (3 sqrt 3) with equality at t^2 = 1/3
        t_hi = np.maximum(np.abs(tl), np.abs(tu))
        t_lo = np.where((lower <= 0) & (upper >= 0), dtype(0), np.minimum(np.abs(tl), np.abs(tu)))
        f2 = lambda t: (dtype(2) * t * (one - t * t)).astype(dtype)
        peak = dtype(0.5773502691896258)
        M = np.where((t_lo <= peak) & (t_hi >= peak), dtype(0.7698003589195010), np.maximum(f2(t_lo), f2(t_hi))).astype(dtype)
        bnd = (M * width * width / dtype(8)).astype(dtype)
        narrow = (width < dtype(1e-2)) & ~zero
        d_hi = np.where(narrow, np.where(lower >= 0, bnd, np.where(upper <= 0, dtype(0), bnd)), d_hi).astype(dtype)
        d_lo = np.where(narrow, np.where(lower >= 0, dtype(0), np.where(upper <= 0, -bnd, -bnd)), d_lo).astype(dtype)
    beta = (rl + dtype(0.5) * (d_hi + d_lo)).astype(dtype)
    delta = (dtype(0.5) * (d_hi - d_lo)).astype(dtype)
    return alpha, beta, delta


def _tanh_rule(ctx, base, aff, err):
    alpha, beta, delta = _tanh_coeffs(base, aff, err)
    return _apply_linear_approx(ctx, base, aff, err, alpha, beta, delta)


def _dense_rule(base, aff, err, A, b):
    """affine_layers.py:11-31 -- base@A+b, every aff row @A, err@|A|."""
    A = np.asarray(A, F32)
    base = base @ A
    aff = np.matmul(aff, A)
    err = err @ np.abs(A)
    if b is not None:
        base = base + np.asarray(b, F32)
    return base.astype(F32), aff.astype(F32), err.astype(F32)


def affine_forward(params, ctx, center, vecs):
    """Propagate the affine form of a general box through the op list.

    center (N,3), vecs (N,v,3)  ->  base (N,), aff (N,k), err (N,) of the scalar output, plus `scale`
    (N,) = sum_j |base_j A_j| + |b| of the last dense layer (ours: the magnitude of what the output was
    summed from, used as the yardstick of tolerances and near-tie bands).
    affine.py:109-117 (input form), mlp.py:99-111 with the 'affine' rules.
    """
    center = np.ascontiguousarray(center, F32)
    vecs = np.ascontiguousarray(vecs, F32)
    n = center.shape[0]
    base = center
    if ctx.mode == "interval":
        aff = np.zeros((n, 0, center.shape[-1]), F32)
        err = np.abs(vecs).sum(axis=1, dtype=F32)
    else:
        aff = vecs
        err = np.zeros_like(center)
    ops = op_list(params)
    last_dense = max(i for i, (nm, _) in enumerate(ops) if nm == "dense")
    scale = None
    for i_op, (name, args) in enumerate(ops):
        if name == "dense":
            if i_op == last_dense:
                A = np.asarray(args["A"], F32)
                scale = (np.abs(base)[:, :, None] * np.abs(A)[None, :, :]).sum(axis=1, dtype=F32)
                if args.get("b") is not None:
                    scale = scale + np.abs(np.asarray(args["b"], F32))
                scale = scale.max(axis=-1).astype(F32)
            base, aff, err = _dense_rule(base, aff, err, args["A"], args.get("b"))
        elif name == "spatial_transformation":
            A, b = _spatial_as_dense(args["R"], args["t"])
            base, aff, err = _dense_rule(base, aff, err, A, b)
        elif name in ("relu", "elu", "sin", "tanh"):
            base, aff, err = {"relu": _relu_rule, "elu": _elu_rule, "sin": _sin_rule, "tanh": _tanh_rule}[name](ctx, base, aff, err)
        elif name == "pow2_frequency_encode":                      # affine_layers.py:140-161
            base = _pow2_encode(base, args["coefs"], args.get("shift"), True)
            aff = _pow2_encode(aff, args["coefs"], None, False)
            err = _pow2_encode(err, args["coefs"], None, False)
        elif name == "squeeze_last":
            assert base.shape[-1] == 1
            base, aff, err = base[:, 0], aff[:, :, 0], err[:, 0]
        else:
            raise ValueError(f"oracle: unsupported op '{name}'")
    return base, aff, err, scale


def sdf_center_value_and_reach(params, ctx, center, vecs):
    """sdf.py:31-43 (WeakSDFImplicitFunction.classify_general_box), batched: f(centre) and lipschitz * radius with
    radius = sqrt(sum_v ||vec_v||^2)."""
    center = np.ascontiguousarray(center, F32)
    vecs = np.ascontiguousarray(vecs, F32)
    nv = np.sqrt((vecs * vecs).sum(axis=-1, dtype=F32)).astype(F32)          # jnp.linalg.norm(box_vecs, axis=-1)
    rad = np.sqrt((nv * nv).sum(axis=-1, dtype=F32)).astype(F32)
    val = eval_points(params, center)
    return val, (rad * F32(ctx.sdf_lipschitz)).astype(F32)


def _slope_activation(name, primal, sc, sw):
    """slope_interval_layers.py:35-58 (relu) / :60-83 (elu), batched: primal (N,w), sc / sw (N,v,w)."""
    sl, su = sc - sw, sc + sw                                              # slope_interval.py:196-199 slope_bounds
    prad = np.maximum(su, -sl).sum(axis=1, dtype=F32)                      # :201-206 primal_may_contain_bounds
    pl, pu = primal - prad, primal + prad
    if name == "sin":                                               # slope_interval_layers.py:85-110
        dfl, dfu = _cos_bound(pl.astype(F32), pu.astype(F32))
        cands = [sl * dfl[:, None, :], sl * dfu[:, None, :], su * dfl[:, None, :], su * dfu[:, None, :]]
        nl, nu = cands[0], cands[0]
        for cnd in cands[1:]:
            nl, nu = np.minimum(nl, cnd), np.maximum(nu, cnd)
        nc = (F32(0.5) * (nl + nu)).astype(F32)
        return np.sin(primal).astype(F32), nc, (nu - nc).astype(F32)
    if name == "relu":
        dfl = np.where(pl > 0, F32(1), F32(0))
        dfu = np.where(pu < 0, F32(0), F32(1))
        new_primal = np.maximum(primal, F32(0))
    elif name == "tanh":            # OURS (parity unpinned): tanh' = 1 - tanh^2 is largest nearest to 0, smallest farthest from it
        far = np.tanh(np.maximum(np.abs(pl), np.abs(pu))).astype(F32)
        near = np.tanh(np.minimum(np.abs(pl), np.abs(pu))).astype(F32)
        dfl = (F32(1) - far * far).astype(F32)
        dfu = np.where((pl <= 0) & (pu >= 0), F32(1), F32(1) - near * near).astype(F32)
        new_primal = np.tanh(primal)
    else:
        with np.errstate(over="ignore"):
            dfl = np.minimum(np.exp(pl), F32(1))
            dfu = np.minimum(np.exp(pu), F32(1))
        new_primal = _elu(primal)
    nl = np.minimum(sl * dfl[:, None, :], sl * dfu[:, None, :])
    nu = np.maximum(su * dfl[:, None, :], su * dfu[:, None, :])
    nc = (F32(0.5) * (nl + nu)).astype(F32)
    return new_primal.astype(F32), nc, (nu - nc).astype(F32)


def slope_forward(params, center, vecs):
    """slope_interval.py:172-194 (input form) + the 'slope_interval' rules of slope_interval_layers.py through the op
    list: -> primal (N,), slope centre (N,v), slope width (N,v) of the scalar output, scale (N,) as in affine_forward."""
    center = np.ascontiguousarray(center, F32)
    vecs = np.ascontiguousarray(vecs, F32)
    primal, sc, sw = center, vecs, np.zeros_like(vecs)
    ops = op_list(params)
    last_dense = max(i for i, (nm, _) in enumerate(ops) if nm == "dense")
    scale = None
    for i_op, (name, args) in enumerate(ops):
        if name in ("dense", "spatial_transformation"):
            if name == "dense":
                A, b = np.asarray(args["A"], F32), args.get("b")
            else:
                A, b = _spatial_as_dense(args["R"], args["t"])
            if i_op == last_dense:
                scale = (np.abs(primal)[:, :, None] * np.abs(A)[None, :, :]).sum(axis=1, dtype=F32)
                if b is not None:
                    scale = scale + np.abs(np.asarray(b, F32))
                scale = scale.max(axis=-1).astype(F32)
            primal = primal @ A
            if b is not None:
                primal = primal + np.asarray(b, F32)
            primal = primal.astype(F32)
            sc = np.matmul(sc, A).astype(F32)
            sw = np.matmul(sw, np.abs(A)).astype(F32)
        elif name in ("relu", "elu", "sin", "tanh"):
            primal, sc, sw = _slope_activation(name, primal, sc, sw)
        elif name == "pow2_frequency_encode":                      # slope_interval_layers.py:112-126
            primal = _pow2_encode(primal, args["coefs"], args.get("shift"), True)
            sc = _pow2_encode(sc, args["coefs"], None, False)
            sw = _pow2_encode(sw, args["coefs"], None, False)
        elif name == "squeeze_last":
            assert primal.shape[-1] == 1
            primal, sc, sw = primal[:, 0], sc[:, :, 0], sw[:, :, 0]
        else:
            raise ValueError(f"oracle: unsupported op '{name}'")
    return primal, sc, sw, scale


def bound_general_box(params, ctx, center, vecs, chunk=None, return_scale=False):
    """-> (lower, upper[, scale]) float32 (N,): affine.py:119-125 applied to the propagated output."""
    center = np.ascontiguousarray(center, F32)
    vecs = np.ascontiguousarray(vecs, F32)
    if ctx.mode == "slope_interval":
        # slope_interval.py:37-44: may-contain bounds of the output
        primal, sc, sw, scale = slope_forward(params, center, vecs)
        prad = np.maximum(sc + sw, -(sc - sw)).sum(axis=1, dtype=F32)
        lower, upper = (primal - prad).astype(F32), (primal + prad).astype(F32)
        return (lower, upper, scale) if return_scale else (lower, upper)
    if ctx.mode == "sdf":
        val, reach = sdf_center_value_and_reach(params, ctx, center, vecs)
        lower, upper = (val - reach).astype(F32), (val + reach).astype(F32)
        if return_scale:
            from . import rays
            return lower, upper, (rays.point_scale(params, center) + reach).astype(F32)
        return lower, upper
    n = center.shape[0]
    if chunk is None:
        chunk = 65536 if ctx.mode in ("interval", "affine_fixed") else 1024
    lower = np.empty(n, F32)
    upper = np.empty(n, F32)
    scale = np.empty(n, F32)
    for s in range(0, n, chunk):
        base, aff, err, sc = affine_forward(params, ctx, center[s:s + chunk], vecs[s:s + chunk])
        rad = (np.abs(aff).sum(axis=1, dtype=F32) + err).astype(F32)
        lower[s:s + chunk] = base - rad
        upper[s:s + chunk] = base + rad
        scale[s:s + chunk] = sc
    if return_scale:
        return lower, upper, scale
    return lower, upper


def slope_min_distance_to_zero(params, box_center, box_axis_vec):
    """slope_interval.py:52-78, batched over N axis-aligned boxes (centre, half extents): -> (raw_primal, distance) (N,)."""
    c = np.ascontiguousarray(box_center, F32).reshape(-1, 3)
    a = np.ascontiguousarray(box_axis_vec, F32).reshape(-1, 3)
    center, vecs = box_to_general(c - a, c + a)                     # coordinates_in_box (:177-181)
    raw_primal, sc, sw, _ = slope_forward(params, center, vecs)
    sl, su = (sc - sw).astype(F32), (sc + sw).astype(F32)           # slope_bounds (:196-199)
    with np.errstate(divide="ignore", invalid="ignore"):
        primal = np.where(raw_primal >= 0, raw_primal, -raw_primal)
        dec = np.maximum(np.abs(sl), np.abs(su))
        vec_len = np.abs(a)
        min_len = vec_len.min(axis=-1)
        dec = np.maximum((dec / vec_len).astype(F32), F32(0))
        axis_decrease = ((dec[:, 0] + dec[:, 1]) + dec[:, 2]).astype(F32)
        dist = np.minimum((primal / axis_decrease).astype(F32), min_len)
        dist = np.where(dist == 0, F32(0), dist).astype(F32)
    return raw_primal, dist


def slope_min_distance_to_zero_in_direction(params, source_point, bound_vec, source_range=None):
    """slope_interval.py:81-163, batched over N: -> (source_val, distance), or with source_range (N,k,3):
    (source_lower, source_upper, distance)."""
    s = np.ascontiguousarray(source_point, F32).reshape(-1, 3)
    b = np.ascontiguousarray(bound_vec, F32).reshape(-1, 3)
    fwd = (b * F32(0.5)).astype(F32)
    center = (s + fwd).astype(F32)
    rng = None if source_range is None else np.ascontiguousarray(source_range, F32).reshape(s.shape[0], -1, 3)
    vecs = fwd[:, None, :] if rng is None else np.concatenate((fwd[:, None, :], rng), axis=1)
    _, sc, sw, _ = slope_forward(params, center, vecs)
    sl, su = (sc - sw).astype(F32), (sc + sw).astype(F32)
    blen = np.sqrt(((b[:, 0] * b[:, 0] + b[:, 1] * b[:, 1]) + b[:, 2] * b[:, 2]).astype(F32)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        if rng is not None:
            sp, ssc, ssw, _ = slope_forward(params, s, rng)
            prad = np.maximum(ssc + ssw, -(ssc - ssw)).sum(axis=1, dtype=F32)
            s_lo, s_up = (sp - prad).astype(F32), (sp + prad).astype(F32)
            is_pos = s_lo >= 0
            val = np.where(is_pos, s_lo, -s_up)
            slope = (F32(2) * np.where(is_pos, sl[:, 0], -su[:, 0]) / blen).astype(F32)
            dist = np.minimum((val / np.maximum(-slope, F32(0))).astype(F32), blen)
            dist = np.where((s_lo <= 0) & (s_up >= 0), F32(0), dist).astype(F32)
            return s_lo, s_up, dist
        sval = eval_points(params, s).astype(F32)
        is_pos = sval >= 0
        slope = (F32(2) * np.where(is_pos, sl[:, 0], -su[:, 0]) / blen).astype(F32)
        dist = np.minimum((np.abs(sval) / np.maximum(-slope, F32(0))).astype(F32), blen)
        dist = np.where(sval == 0, F32(0), dist).astype(F32)
    return sval, dist


def labels_from_bounds(lower, upper, offset=0.0):
    """affine.py:49-53: POSITIVE if lower > offset, then NEGATIVE if upper < -offset (wins)."""
    offset = F32(offset)
    out = np.full(lower.shape, SIGN_UNKNOWN, np.int32)
    out = np.where(lower > offset, SIGN_POSITIVE, out)
    out = np.where(upper < -offset, SIGN_NEGATIVE, out)
    return out.astype(np.int32)


def classify_general_box(params, ctx, center, vecs, offset=0.0, return_bounds=False, return_scale=False):
    """affine.py:34-55, batched over N boxes."""
    lower, upper, scale = bound_general_box(params, ctx, center, vecs, return_scale=True)
    if ctx.mode == "sdf":
        # sdf.py:42-48: the offset is tested on f(centre) itself, not on the bounds
        val, reach = sdf_center_value_and_reach(params, ctx, center, vecs)
        can_change = (np.abs(val) - reach) < 0
        lab = np.full(val.shape, SIGN_UNKNOWN, np.int32)
        lab = np.where(~can_change & (val > F32(offset)), SIGN_POSITIVE, lab)
        lab = np.where(~can_change & (val < -F32(offset)), SIGN_NEGATIVE, lab).astype(np.int32)
    else:
        lab = labels_from_bounds(lower, upper, offset)
    if return_scale:
        return lab, lower, upper, scale
    return (lab, lower, upper) if return_bounds else lab


def box_to_general(lo, hi):
    """implicit_function.py:28-37: centre = 0.5*(lo+hi), vecs = diag(hi - centre)."""
    lo = np.asarray(lo, F32)
    hi = np.asarray(hi, F32)
    center = (F32(0.5) * (lo + hi)).astype(F32)
    half = (hi - center).astype(F32)
    vecs = np.zeros(lo.shape[:-1] + (3, 3), F32)
    for i in range(3):
        vecs[..., i, i] = half[..., i]
    return center, vecs


def classify_box(params, ctx, lo, hi, offset=0.0, return_bounds=False, return_scale=False):
    center, vecs = box_to_general(lo, hi)
    return classify_general_box(params, ctx, center, vecs, offset, return_bounds, return_scale)


# ----------------------------------------------------------------------------------------------
# near-tie bands (ours, not the reference's): which decisions are within float32 summation noise
# ----------------------------------------------------------------------------------------------

NEAR_TIE_REL = 1e-5
NEAR_TIE_REL_ELU = 2e-4
NEAR_TIE_REL_TANH = 1e-4   # nets with a tanh layer (ours, unpinned): the secant slope and atanh(sqrt(1 - alpha)) amplify 1-ulp differences of tanhf
NEAR_TIE_REL_SIN = 5e-5    # nets with a sin layer: float32 vs float64 of the same formulas differ by up to 1.6e-5 (tests)


def tie_rel(params):
    """Relative tolerance / near-tie band for bounds of this MLP: 1e-5 (BASELINE north_star) for relu-only
    nets; 2e-4 for nets with elu, whose rule (affine_layers.py:59-97) is ill-conditioned in float32 --
    delta = |r_upper - r_lower|/2 cancels O(1) terms, so two IEEE-correct implementations that differ by
    1 ulp in exp/log/expm1 disagree by ~1e-7 ABSOLUTE per neuron, amplified by the following layers
    (tests/test_oracle_golden.py::test_elu_rule_conditioning measures it against float64)."""
    names = {nm for nm, _ in op_list(params)}
    if "elu" in names:
        return NEAR_TIE_REL_ELU
    if "tanh" in names:
        return NEAR_TIE_REL_TANH
    return NEAR_TIE_REL_SIN if "sin" in names else NEAR_TIE_REL


def mode_rel(params, ctx):
    """Relative tolerance / near-tie band of a mode on a net: tie_rel(params) for the four hot-path modes.
    affine_append is 100x wider: its `err + (sum(delta) - sum(kept))` (src/affine.py:191) subtracts two nearly equal
    float32 sums and broadcasts the noisy scalar to every neuron, where the remaining layers amplify it.  Measured
    on 1,200 random boxes per sample net, float32 vs float64 evaluation of the SAME formulas deviates by up to
    2e-4 (fox) / 6e-4 (birdcage) / 7e-3 (bunny) of the yardstick in affine_append against <= 8e-6 / 4e-5 in
    affine_fixed (tests/test_oracle_golden.py::test_append_is_ill_conditioned_in_float32): no float32 implementation
    can agree with another more closely than that."""
    return (100.0 if ctx.mode == "affine_append" else 1.0) * tie_rel(params)


def tol_scale(lower, upper, scale=None):
    """Yardstick of a bound: |base| + rad (= max(|lower|,|upper|)) plus, when given, the magnitude the
    output was summed from (sum_j |base_j A_j| + |b| of the last layer).  float32 summation-order noise
    is a few ulp of THIS, not of the (possibly cancelled) bound itself."""
    s = np.maximum(np.abs(lower.astype(np.float64)), np.abs(upper.astype(np.float64)))
    if scale is not None:
        s = s + scale.astype(np.float64)
    return s


def bound_near_tie(lower, upper, offset=0.0, scale=None, rel=NEAR_TIE_REL):
    """A box is 'near-tie' when either threshold test sits within rel*tol_scale of flipping
    (BASELINE.json north_star: such boxes are counted and excluded from the bit-exact label check)."""
    s = tol_scale(lower, upper, scale)
    lower = lower.astype(np.float64)
    upper = upper.astype(np.float64)
    return (np.abs(lower - offset) <= rel * s) | (np.abs(upper + offset) <= rel * s)
