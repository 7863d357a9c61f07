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


def render_image(ctx_tuple, params_tuple, eye_pos, look_dir, up_dir, res, fov_deg, opts, frustum=False, left_dir=None):
    """render.py:94-150 with shading='normal', no tonemap; frustum=True takes the cast_rays_frustum branch (:116-126), whose
    (res_x, res_y) images are transposed into the ray order of generate_camera_rays."""
    roots, dirs = generate_camera_rays(eye_pos, look_dir, up_dir, res=res, fov_deg=fov_deg)
    if frustum:
        cam = (eye_pos, look_dir, up_dir, left_dir, fov_deg, fov_deg, res, res)
        t, hit, cnt, n_eval = cast_rays_frustum(ctx_tuple, params_tuple, cam, opts)
        t, hit, cnt = t.transpose().flatten(), hit.transpose().flatten(), cnt.transpose().flatten()
    else:
        t, hit, cnt, n_eval = cast_rays(ctx_tuple, params_tuple, roots, dirs, opts)
    hit_pos = (roots + t[:, None] * dirs).astype(F32)
    nrm = outward_normals(params_tuple, hit_pos, hit, opts["hit_eps"])
    color = ((nrm + F32(1.)) / F32(2.)).astype(F32)
    img = np.where((hit != 0)[:, None], color, np.ones((res * res, 3), F32)).astype(F32)
    return img.reshape(res, res, 3), t.reshape(res, res), cnt.reshape(res, res), hit.reshape(res, res), n_eval


# ----------------------------------------------------------------------------------------------
# cast_rays_frustum (queries.py:178-587): frusta of pixels marched together, split when too wide
# ----------------------------------------------------------------------------------------------

def camera_ray(look_dir, up_dir, left_dir, fov_deg_x, fov_deg_y, theta_x, theta_y):
    """render.py:17-24 for arrays of image-plane coordinates theta in [-1, 1]."""
    tan_x = np.tan(np.deg2rad(F32(fov_deg_x)) / F32(2)).astype(F32)
    tan_y = np.tan(np.deg2rad(F32(fov_deg_y)) / F32(2)).astype(F32)
    plane = (np.asarray(look_dir, F32)[None, :]
             + np.asarray(left_dir, F32)[None, :] * (theta_x * tan_x).astype(F32)[:, None]
             + np.asarray(up_dir, F32)[None, :] * (theta_y * tan_y).astype(F32)[:, None]).astype(F32)
    return _normalize(plane)


def _frustum_steps(funcs, params_tuple, cam_params, opts, n_substeps, frust_range, t, step_size, tie):
    """queries.py:202-337 (take_step / take_several_steps) for all frusta at once."""
    root_pos, look_dir, up_dir, left_dir, fov_x, fov_y, res_x, res_y = cam_params
    root_pos = np.asarray(root_pos, F32)
    n = t.shape[0]
    hit_eps = F32(opts["hit_eps"])
    x_lower, y_lower, x_upper, y_upper = (frust_range[:, k] for k in range(4))
    is_single_pixel = (x_lower + 1 == x_upper) & (y_lower + 1 == y_upper)
    # pixel = point sample: the -1 on the upper coordinates (queries.py:283-293)
    xc_lower = ((F32(2) * x_lower.astype(F32)) / F32(res_x + 1.0) - F32(1)).astype(F32)
    xc_upper = ((F32(2) * (x_upper - 1).astype(F32)) / F32(res_x + 1.0) - F32(1)).astype(F32)
    yc_lower = ((F32(2) * y_lower.astype(F32)) / F32(res_y + 1.0) - F32(1)).astype(F32)
    yc_upper = ((F32(2) * (y_upper - 1).astype(F32)) / F32(res_y + 1.0) - F32(1)).astype(F32)
    gen = lambda tx, ty: camera_ray(look_dir, up_dir, left_dir, fov_x, fov_y, tx, ty)
    ray_xu_yu, ray_xl_yu = gen(xc_upper, yc_upper), gen(xc_lower, yc_upper)
    ray_xu_yl, ray_xl_yl = gen(xc_upper, yc_lower), gen(xc_lower, yc_lower)
    mid_ray = (F32(0.5) * (ray_xu_yu + ray_xl_yl)).astype(F32)
    mid_ray_len = np.sqrt((mid_ray * mid_ray).sum(axis=-1, dtype=F32)).astype(F32)
    mid_ray = (mid_ray / mid_ray_len[:, None]).astype(F32)
    expand_fac = (F32(1) / mid_ray_len).astype(F32)

    is_hit = np.zeros(n, bool)
    hit_id = np.zeros(n, np.int32)
    step_count = np.zeros(n, np.int32)
    step_demands_subd = np.zeros(n, bool)
    for _ in range(n_substeps):
        t_upper = (t + step_size).astype(F32)
        t_upper_adj = (t_upper * expand_fac).astype(F32)
        right_front = ((ray_xu_yu - ray_xl_yu) * t_upper_adj[:, None] / F32(2)).astype(F32)
        up_front = ((ray_xu_yu - ray_xu_yl) * t_upper_adj[:, None] / F32(2)).astype(F32)
        can_step = ~is_hit
        step_count = step_count + (~is_hit)
        center_mid = (root_pos[None, :] + (F32(0.5) * (t + t_upper_adj))[:, None] * mid_ray).astype(F32)
        center_vec = ((F32(0.5) * (t_upper_adj - t))[:, None] * mid_ray).astype(F32)
        box_vecs = np.stack((center_vec, right_front, up_front), axis=1)
        func_id = 1
        for ctx, params in zip(funcs, params_tuple):
            lo, up, bsc = net.bound_general_box(params, ctx, center_mid, box_vecs, return_scale=True)
            box_type = net.labels_from_bounds(lo, up, 0.0)
            can_step = can_step & ((box_type == net.SIGN_POSITIVE) | (box_type == net.SIGN_NEGATIVE))
            pos_start = (root_pos[None, :] + t[:, None] * mid_ray).astype(F32)
            pos_eps = (root_pos[None, :] + (t + hit_eps)[:, None] * mid_ray).astype(F32)
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
        # a failed step inches forward only for single-pixel frusta (queries.py:249-254)
        this_step = np.where(can_step, step_size, hit_eps * is_single_pixel.astype(F32)).astype(F32)
        t = np.where(is_hit, t, t + this_step * F32(opts["safety_factor"])).astype(F32)
        step_size = np.where(can_step, step_size * F32(opts["interval_grow_fac"]),
                             step_size * F32(opts["interval_shrink_fac"])).astype(F32)
        step_demands_subd = step_demands_subd | (step_size < hit_eps) | is_hit
        step_size = np.maximum(step_size, hit_eps)
    return t, step_size, is_hit, hit_id, step_demands_subd, step_count


def subdivide_frusta(sub_mask, empty_start_ind, valid_mask, frust_range, arrs):
    """queries.py:371-432: halve the longer pixel axis (x on ties); child A replaces the entry, the B children are
    appended from empty_start_ind in index order.  The midpoint is an integer store of (lo+hi)/2: truncation."""
    sub_mask = sub_mask & valid_mask
    x_gap = frust_range[:, 2] - frust_range[:, 0]
    y_gap = frust_range[:, 3] - frust_range[:, 1]
    subd_x = x_gap >= y_gap
    x_mid = ((frust_range[:, 0] + frust_range[:, 2]) / 2).astype(F32).astype(np.int32)
    y_mid = ((frust_range[:, 1] + frust_range[:, 3]) / 2).astype(F32).astype(np.int32)
    range_A = frust_range.copy()
    range_A[:, 2] = np.where(subd_x, x_mid, frust_range[:, 2])
    range_A[:, 3] = np.where(~subd_x, y_mid, frust_range[:, 3])
    range_B = frust_range.copy()
    range_B[:, 0] = np.where(subd_x, x_mid, frust_range[:, 0])
    range_B[:, 1] = np.where(~subd_x, y_mid, frust_range[:, 1])
    idx = np.nonzero(sub_mask)[0]
    out_range = np.where(sub_mask[:, None], range_A, frust_range).astype(np.int32)
    dst = empty_start_ind + np.arange(idx.shape[0])
    assert idx.shape[0] == 0 or dst[-1] < frust_range.shape[0], "no room to subdivide"
    out_range[dst] = range_B[idx]
    outs = []
    for a in arrs:
        o = a.copy()
        o[dst] = a[idx]
        outs.append(o)
    valid_mask = valid_mask.copy()
    valid_mask[dst] = True
    return valid_mask, out_range, outs


def cast_rays_frustum(funcs_tuple, params_tuple, cam_params, opts, return_near_tie=False, iter_counts=None, init_ranges=None):
    """queries.py:465-587 -> (out_t (res_x,res_y) f32, out_hit_id i32, out_count i32, N_evals[, near_tie]).
    N_evals counts the padded array length of every marching iteration (queries.py:523), NOT times n_substeps.
    iter_counts (ours): a list that receives (terminated, split) per iteration; init_ranges (ours): replaces the initial
    tiles (a rank's share of a sharded image; the other pixels stay zero)."""
    root_pos, look_dir, up_dir, left_dir, fov_x, fov_y, res_x, res_y = cam_params
    n_substeps = int(opts["n_substeps"])
    N_out = res_x * res_y
    N_side_init = int(opts["n_side_init"])
    N_init = N_side_init ** 2
    N_evals = 0
    # initial tiles (queries.py:495-501); jnp.linspace(dtype=int) floors
    x_ticks = np.floor(np.linspace(0, res_x, N_side_init + 1, dtype=F32)).astype(np.int32)
    y_ticks = np.floor(np.linspace(0, res_y, N_side_init + 1, dtype=F32)).astype(np.int32)
    x_start, x_end = np.tile(x_ticks[:-1], N_side_init), np.tile(x_ticks[1:], N_side_init)
    y_start, y_end = np.repeat(y_ticks[:-1], N_side_init), np.repeat(y_ticks[1:], N_side_init)
    cur_range = np.stack((x_start, y_start, x_end, y_end), axis=-1).astype(np.int32)
    if init_ranges is not None:
        cur_range = np.array(init_ranges, np.int32).reshape(-1, 4)
        N_init = cur_range.shape[0]
    cur_t = np.zeros(N_init, F32)
    cur_size = (np.ones(N_init, F32) * F32(opts["interval_init_size"]) * F32(opts["max_dist"])).astype(F32)
    cur_count = np.zeros(N_init, F32)
    cur_valid = np.ones(N_init, bool)
    cur_tie = np.zeros(N_init, bool)
    empty_start_ind = N_init

    fin_range = np.zeros((N_out, 4), np.int32)
    fin_t = np.zeros(N_out, F32)
    fin_hit = np.zeros(N_out, np.int32)
    fin_count = np.zeros(N_out, F32)
    fin_tie = np.zeros(N_out, bool)
    fin_start = 0

    it = 0
    while True:
        N_evals += cur_t.shape[0]
        v = np.nonzero(cur_valid)[0]                 # padding entries are evaluated by the reference but never read
        tie = cur_tie[v].copy() if return_near_tie else None
        t_v, size_v, is_hit, hit_id, demands, n_inner = _frustum_steps(
            funcs_tuple, params_tuple, cam_params, opts, n_substeps, cur_range[v], cur_t[v], cur_size[v], tie)
        cur_t[v], cur_size[v] = t_v, size_v
        if return_near_tie:
            cur_tie[v] = tie
        r = cur_range[v]
        area = (r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1])
        cur_count[v] = (cur_count[v] + (n_inner.astype(F32) * (F32(1.0) / area.astype(F32))).astype(F32)).astype(F32)
        is_hit = is_hit & (area == 1)                # only single-pixel frusta get to hit
        is_miss = t_v > F32(opts["max_dist"])
        terminated = is_hit | is_miss | (it >= opts["n_max_step"])
        w = v[terminated]
        k = w.shape[0]
        fin_range[fin_start:fin_start + k] = cur_range[w]
        fin_t[fin_start:fin_start + k] = cur_t[w]
        fin_hit[fin_start:fin_start + k] = hit_id[terminated]     # whatever the substeps left, hit or not
        fin_count[fin_start:fin_start + k] = cur_count[w]
        fin_tie[fin_start:fin_start + k] = cur_tie[w]
        cur_valid[w] = False
        fin_start += k
        # who needs to be split (queries.py:350-360)
        half_fov_x = (np.deg2rad(F32(fov_x)) / F32(2)).astype(F32)
        half_fov_y = (np.deg2rad(F32(fov_y)) / F32(2)).astype(F32)
        width_x = (F32(2) * np.sin((half_fov_x * (r[:, 2] - r[:, 0]).astype(F32) / F32(res_x)).astype(F32)) * t_v).astype(F32)
        width_y = (F32(2) * np.sin((half_fov_y * (r[:, 3] - r[:, 1]).astype(F32) / F32(res_y)).astype(F32)) * t_v).astype(F32)
        can_subd = (r[:, 2] > r[:, 0] + 1) | (r[:, 3] > r[:, 1] + 1)
        lim = (F32(opts["refine_width_fac"]) * size_v).astype(F32)
        refine_v = ((width_x > lim) | (width_y > lim) | demands) & can_subd & ~terminated
        needs_refine = np.zeros(cur_valid.shape[0], bool)
        needs_refine[v] = refine_v
        it += n_substeps
        N_valid = int(cur_valid.sum())
        N_needs_refine = int(needs_refine.sum())
        if iter_counts is not None:
            iter_counts.append((k, N_needs_refine))
        if N_valid == 0:
            break
        new_bucket = get_next_bucket_size(N_valid + N_needs_refine)
        cur_bucket = cur_valid.shape[0]
        if empty_start_ind + N_needs_refine > cur_bucket or new_bucket < cur_bucket:
            (cur_valid, empty_start_ind, cur_range, cur_t, cur_size, cur_count, needs_refine, cur_tie) = \
                compactify_and_rebucket(cur_valid, new_bucket, cur_range, cur_t, cur_size, cur_count, needs_refine, cur_tie)
        cur_valid, cur_range, (cur_t, cur_size, cur_count, cur_tie) = subdivide_frusta(
            needs_refine, empty_start_ind, cur_valid, cur_range, [cur_t, cur_size, cur_count, cur_tie])
        empty_start_ind += N_needs_refine

    # (2) split the finished frusta down to single pixels, children inherit everything (queries.py:558-577)
    fin_valid = np.arange(N_out) < fin_start
    while True:
        single = (fin_range[:, 0] + 1 == fin_range[:, 2]) & (fin_range[:, 1] + 1 == fin_range[:, 3])
        needs = fin_valid & ~single
        if not needs.any():
            break
        fin_valid, fin_range, (fin_t, fin_hit, fin_count, fin_tie) = subdivide_frusta(
            needs, fin_start, fin_valid, fin_range, [fin_t, fin_hit, fin_count, fin_tie])
        fin_start += int(needs.sum())
    # (3) one pixel per frustum; the float count lands in an int image: truncation (queries.py:442-456)
    assert fin_start == N_out or init_ranges is not None
    fin_range, fin_t, fin_hit, fin_count, fin_tie = (a[:fin_start] for a in (fin_range, fin_t, fin_hit, fin_count, fin_tie))
    out_t = np.zeros((res_x, res_y), F32)
    out_hit = np.zeros((res_x, res_y), np.int32)
    out_count = np.zeros((res_x, res_y), np.int32)
    out_tie = np.zeros((res_x, res_y), bool)
    xs, ys = fin_range[:, 0], fin_range[:, 1]
    out_t[xs, ys] = fin_t
    out_hit[xs, ys] = fin_hit
    out_count[xs, ys] = fin_count.astype(np.int32)
    out_tie[xs, ys] = fin_tie
    if return_near_tie:
        return out_t, out_hit, out_count, N_evals, out_tie
    return out_t, out_hit, out_count, N_evals
