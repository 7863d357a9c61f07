"""CPU oracle, part 2: bucketing helpers, camera-ray generator and `cast_rays`.

TEST INFRASTRUCTURE ONLY (see net.py header).  Follows /root/reference/src:
  bucketing.py:7-36, queries.py:23-175, render.py:17-50,229-242, geometry.py:6-22.
The host-level control flow (one pass per step iteration, power-of-two bucket compaction) is kept
because `N_evals` counts padded lanes (queries.py:137,164).
"""
import numpy as np

from . import net

F32 = np.float32

BUCKET_SIZES = [2 ** s for s in range(7, 31)]            # bucketing.py:7-9


def get_next_bucket_size(s):
    """bucketing.py:10-14"""
    for b in BUCKET_SIZES:
        if s <= b:
            return b
    raise ValueError("max bucket size exceeded")


def fits_in_smaller_bucket(size, curr_bucket_size):
    """bucketing.py:35-36"""
    return get_next_bucket_size(size) < curr_bucket_size


def compactify_and_rebucket(mask, bucket_size, *arrs):
    """bucketing.py:16-32: order-preserving compaction of the masked rows, padded to bucket_size.
    Padding rows are unspecified in the reference (out-of-range gather); we zero them."""
    idx = np.nonzero(mask)[0]
    n_in = idx.shape[0]
    out_mask = np.arange(bucket_size) < n_in
    outs = []
    for a in arrs:
        o = np.zeros((bucket_size,) + a.shape[1:], a.dtype)
        o[:min(n_in, bucket_size)] = a[idx[:bucket_size]]
        outs.append(o)
    return (out_mask, n_in, *outs)


def get_default_cast_opts():
    """queries.py:23-36"""
    return {
        "hit_eps": 0.001, "max_dist": 10.0, "n_max_step": 512, "n_substeps": 1,
        "safety_factor": 0.98, "interval_grow_fac": 1.5, "interval_shrink_fac": 0.5,
        "interval_init_size": 0.1, "refine_width_fac": 2.0, "n_side_init": 16,
    }


# ----------------------------------------------------------------------------------------------
# camera (render.py:17-50, 229-242; geometry.py:6-22), generalised from res x res to res_x x res_y
# ----------------------------------------------------------------------------------------------

def _normalize(v):
    v = np.asarray(v, F32)
    return (v / np.sqrt((v * v).sum(axis=-1, keepdims=True, dtype=F32))).astype(F32)


def look_at(eye_pos, target=(0.0, 0.0, 0.0), up_dir=(0.0, 1.0, 0.0)):
    eye_pos = np.asarray(eye_pos, F32)
    look = _normalize(np.asarray(target, F32) - eye_pos)
    up = np.asarray(up_dir, F32)
    up = _normalize(up - np.dot(up, look).astype(F32) * look)
    left = np.cross(look, up).astype(F32)
    return look, up, left


def generate_camera_rays(eye_pos, look_dir, up_dir, res=1024, fov_deg=30.0, res_y=None):
    """Pinhole rays, pixel order y-major (meshgrid default 'xy' then flatten), render.py:26-50.
    `res_y` (ours) generalises the square image: the same fov is used on both axes as in the
    reference, with res_y samples on the vertical axis."""
    res_x = res
    res_y = res if res_y is None else res_y
    eye_pos = np.asarray(eye_pos, F32)
    look_dir = np.asarray(look_dir, F32)
    up_dir = np.asarray(up_dir, F32)
    ax_x = np.linspace(-1.0, 1.0, res_x, dtype=F32)
    ax_y = np.linspace(-1.0, 1.0, res_y, dtype=F32)
    cam_x, cam_y = np.meshgrid(ax_x, ax_y)
    cam_x = cam_x.reshape(-1)
    cam_y = cam_y.reshape(-1)
    up_dir = up_dir - np.dot(look_dir, up_dir).astype(F32) * look_dir
    up_dir = _normalize(up_dir)
    left_dir = np.cross(look_dir, up_dir).astype(F32)
    tan_half = np.tan(np.deg2rad(F32(fov_deg)) / F32(2)).astype(F32)
    plane = (look_dir[None, :]
             + left_dir[None, :] * (cam_x * tan_half)[:, None]
             + up_dir[None, :] * (cam_y * tan_half)[:, None]).astype(F32)
    dirs = _normalize(plane)
    roots = np.tile(eye_pos[None, :], (dirs.shape[0], 1)).astype(F32)
    return roots, dirs


# ----------------------------------------------------------------------------------------------
# cast_rays (queries.py:39-175)
# ----------------------------------------------------------------------------------------------

def _sign(x):
    return np.sign(x)


def _take_steps(funcs, params_tuple, opts, roots, dirs, t, step_size, n_substeps, tie=None):
    """queries.py:44-107 for all lanes at once.  funcs = tuple of AffineContext."""
    n = roots.shape[0]
    hit_eps = F32(opts["hit_eps"])
    is_hit = np.zeros(n, bool)
    hit_id = np.zeros(n, np.int32)
    step_count = np.zeros(n, np.int32)
    for _ in range(n_substeps):
        can_step = ~is_hit
        step_count = step_count + (~is_hit)
        func_id = 1
        for ctx, params in zip(funcs, params_tuple):
            pos_start = (roots + t[:, None] * dirs).astype(F32)
            half_vec = (F32(0.5) * step_size[:, None] * dirs).astype(F32)
            pos_mid = (pos_start + half_vec).astype(F32)
            lo, up, bsc = net.bound_general_box(params, ctx, pos_mid, half_vec[:, None, :], return_scale=True)
            box_type = net.labels_from_bounds(lo, up, 0.0)
            can_step = can_step & ((box_type == net.SIGN_POSITIVE) | (box_type == net.SIGN_NEGATIVE))
            pos_eps = (roots + (t + hit_eps)[:, None] * dirs).astype(F32)
            val_start = net.eval_points(params, pos_start)
            val_eps = net.eval_points(params, pos_eps)
            if tie is not None:
                tie |= net.bound_near_tie(lo, up, 0.0, bsc, rel=net.tie_rel(params))
                tie |= _point_sign_near_tie(params, pos_start, val_start)
                tie |= _point_sign_near_tie(params, pos_eps, val_eps)
            this_is_hit = _sign(val_start) != _sign(val_eps)
            hit_id = np.where(this_is_hit, func_id, hit_id).astype(np.int32)
            is_hit = is_hit | this_is_hit
            func_id += 1
        this_step = np.where(can_step, step_size, hit_eps).astype(F32)
        t = np.where(is_hit, t, t + this_step * F32(opts["safety_factor"])).astype(F32)
        step_size = np.where(can_step, step_size * F32(opts["interval_grow_fac"]),
                             step_size * F32(opts["interval_shrink_fac"])).astype(F32)
        step_size = np.maximum(step_size, hit_eps)                      # jnp.clip(a_min=hit_eps)
    return t, step_size, is_hit, hit_id, step_count


def point_scale(params, x):
    """Magnitude scale of the last dot product of a point evaluation: sum_j |h_j * A_j| + |b|.
    Used only for the near-tie band of sign tests (ours, not the reference's)."""
    ops = net.op_list(params)
    # find the last dense
    last = max(i for i, (nm, _) in enumerate(ops) if nm == "dense")
    h = np.ascontiguousarray(x, dtype=F32)
    for i, (name, args) in enumerate(ops):
        if i == last:
            A = np.asarray(args["A"], F32)
            s = np.abs(h[:, :, None] * A[None, :, :]).sum(axis=1, dtype=F32)
            if "b" in args:
                s = s + np.abs(np.asarray(args["b"], F32))
            return s[:, 0] if s.shape[-1] == 1 else s.max(axis=-1)
        if name == "dense":
            h = h @ np.asarray(args["A"], F32)
            if "b" in args:
                h = h + np.asarray(args["b"], F32)
        elif name == "spatial_transformation":
            A, b = net._spatial_as_dense(args["R"], args["t"])
            h = h @ A + b
        elif name == "relu":
            h = np.maximum(h, F32(0))
        elif name == "elu":
            h = net._elu(h)
        elif name == "sin":
            h = np.sin(h)
        elif name == "pow2_frequency_encode":
            h = net._pow2_encode(h, args["coefs"], args.get("shift"), True)
        h = h.astype(F32, copy=False)
    raise ValueError("no dense layer")


def _point_sign_near_tie(params, x, f, rel=net.NEAR_TIE_REL):
    return np.abs(f.astype(np.float64)) <= rel * point_scale(params, x).astype(np.float64)


def cast_rays(funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie=False):
    """queries.py:134-175.  funcs_tuple holds AffineContext objects (the oracle's stand-in for the
    reference's ImplicitFunction); returns (out_t, out_hit_id, out_count, N_evals[, near_tie])."""
    roots = np.ascontiguousarray(roots, F32)
    dirs = np.ascontiguousarray(dirs, F32)
    N = roots.shape[0]
    n_substeps = int(opts["n_substeps"])
    N_evals = 0
    out_t = np.zeros(N, F32)
    out_hit_id = np.zeros(N, np.int32)
    out_count = np.zeros(N, np.int32)
    near_tie_out = np.zeros(N, bool)

    curr_roots, curr_dirs = roots, dirs
    curr_t = np.zeros(N, F32)
    curr_int_size = (np.ones(N, F32) * F32(opts["interval_init_size"]) * F32(opts["max_dist"])).astype(F32)
    curr_inds = np.arange(N, dtype=np.int32)
    curr_valid = np.ones(N, bool)
    curr_count = np.zeros(N, np.int32)
    curr_tie = np.zeros(N, bool)

    while True:
        tie = curr_tie if return_near_tie else None
        curr_t, curr_int_size, is_hit, hit_id, n_inner = _take_steps(
            funcs_tuple, params_tuple, opts, curr_roots, curr_dirs, curr_t, curr_int_size, n_substeps, tie)
        curr_count = curr_count + curr_valid * n_inner
        is_miss = curr_t > F32(opts["max_dist"])
        is_count_terminate = curr_count >= opts["n_max_step"]
        terminated = (is_hit | is_miss | is_count_terminate) & curr_valid
        w = curr_inds[terminated]
        out_t[w] = curr_t[terminated]
        out_hit_id[w] = hit_id[terminated]
        out_count[w] = curr_count[terminated]
        near_tie_out[w] = curr_tie[terminated]
        curr_valid = curr_valid & ~terminated
        N_evals += curr_t.shape[0] * n_substeps
        N_valid = int(curr_valid.sum())
        if N_valid == 0:
            break
        if fits_in_smaller_bucket(N_valid, curr_valid.shape[0]):
            nb = get_next_bucket_size(N_valid)
            (curr_valid, _, curr_roots, curr_dirs, curr_t, curr_int_size, curr_inds, curr_count, curr_tie) = \
                compactify_and_rebucket(curr_valid, nb, curr_roots, curr_dirs, curr_t, curr_int_size,
                                        curr_inds, curr_count, curr_tie)
    if return_near_tie:
        return out_t, out_hit_id, out_count, N_evals, near_tie_out
    return out_t, out_hit_id, out_count, N_evals


# ----------------------------------------------------------------------------------------------
# the direct caller of cast_rays: normals + 'normal' shading + image assembly (render.py:53-165)
# ----------------------------------------------------------------------------------------------

def outward_normals(params_tuple, hit_pos, hit_ids, eps):
    """render.py:53-90, method 'finite_differences'."""
    hit_pos = np.ascontiguousarray(hit_pos, F32)
    eps = F32(eps)
    offsets = np.array(((+eps, -eps, -eps), (-eps, -eps, +eps), (-eps, +eps, -eps), (+eps, +eps, +eps)), F32)
    x_pts = (hit_pos[:, None, :] + offsets[None, :, :]).astype(F32)
    out = np.zeros_like(hit_pos)
    for i_func, params in enumerate(params_tuple, start=1):
        samples = net.eval_points(params, x_pts.reshape(-1, 3)).reshape(-1, 4)
        grad = (offsets[None, :, :] * samples[:, :, None]).sum(axis=1, dtype=F32)
        with np.errstate(invalid="ignore", divide="ignore"):
            grad = (grad / np.linalg.norm(grad, axis=-1, keepdims=True)).astype(F32)
        out = np.where((np.asarray(hit_ids) == i_func)[:, None], grad, out).astype(F32)
    return out


def render_image(ctx_tuple, params_tuple, eye_pos, look_dir, up_dir, res, fov_deg, opts):
    """render.py:94-150 with frustum=False, shading='normal', no tonemap."""
    roots, dirs = generate_camera_rays(eye_pos, look_dir, up_dir, res=res, fov_deg=fov_deg)
    t, hit, cnt, n_eval = cast_rays(ctx_tuple, params_tuple, roots, dirs, opts)
    hit_pos = (roots + t[:, None] * dirs).astype(F32)
    nrm = outward_normals(params_tuple, hit_pos, hit, opts["hit_eps"])
    color = ((nrm + F32(1.)) / F32(2.)).astype(F32)
    img = np.where((hit != 0)[:, None], color, np.ones((res * res, 3), F32)).astype(F32)
    return img.reshape(res, res, 3), t.reshape(res, res), cnt.reshape(res, res), hit.reshape(res, res), n_eval
