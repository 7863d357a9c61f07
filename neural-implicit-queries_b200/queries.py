"""`queries` module surface of the reference for the hot path: get_default_cast_opts, cast_rays
(/root/reference/src/queries.py:23-36, :39-175) and cast_rays_frustum (:178-587).

interval / affine_fixed / slope_interval: ONE persistent CUDA kernel (csrc/niq_kernels.cuh k_cast_rays) marches every ray
to termination with an in-kernel work queue -- no per-iteration host round trip, no bucket padding.
affine_all / affine_truncate / affine_append: the same, one CTA per ray in flight (csrc/niq_rays_grow.cuh k_cast_rays_grow).
sdf (and NIQ_RAYS_HOST_LOOP=1, the A/B switch of the tests): the reference's host-level iteration (one pass per step,
order-preserving compaction) with the bound / point evaluations batched on the GPU."""
import ctypes as C
import os

import numpy as np

import _niq
from bucketing import fits_in_smaller_bucket, get_next_bucket_size


# every mode of the reference's factory except 'sdf' marches its rays in ONE persistent kernel (k_cast_rays for the fixed-row modes,
# k_cast_rays_grow for the growing affine forms); NIQ_RAYS_HOST_LOOP=1 (A/B tests) forces the host-level iteration below
_PERSISTENT_RAY_MODES = {"interval", "affine_fixed", "slope_interval", "affine_truncate", "affine_all", "affine_append"}


def get_default_cast_opts():
    return {
        'hit_eps': 0.001,
        'max_dist': 10.,
        'n_max_step': 512,
        'n_substeps': 1,
        'safety_factor': 0.98,
        'interval_grow_fac': 1.5,
        'interval_shrink_fac': 0.5,
        'interval_init_size': 0.1,   # relative, as a factor of max_dist
        'refine_width_fac': 2.,
        'n_side_init': 16,
    }


def _opts_struct(opts):
    o = _niq.CastOpts()
    o.hit_eps = opts['hit_eps']
    o.max_dist = opts['max_dist']
    o.n_max_step = int(opts['n_max_step'])
    o.n_substeps = int(opts['n_substeps'])
    o.safety_factor = opts['safety_factor']
    o.interval_grow_fac = opts['interval_grow_fac']
    o.interval_shrink_fac = opts['interval_shrink_fac']
    o.interval_init_size = opts['interval_init_size']
    return o


def cast_rays(funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie=False, ctx=None):
    """-> (out_t (N,) f32, out_hit_id (N,) i32, out_count (N,) i32, N_evals int[, near_tie (N,) bool])."""
    ctx = ctx or _niq.default_context()
    if len(funcs_tuple) != len(params_tuple) or len(funcs_tuple) < 1:
        raise ValueError("funcs_tuple and params_tuple must have the same (non-zero) length")
    roots = np.ascontiguousarray(roots, np.float32)
    dirs = np.ascontiguousarray(dirs, np.float32)
    if roots.ndim != 2 or roots.shape[1] != 3 or roots.shape != dirs.shape:
        raise ValueError("roots and dirs must both have shape (N,3)")
    modes = {f.ctx.mode for f in funcs_tuple}
    if len(modes) == 1 and modes <= _PERSISTENT_RAY_MODES and not os.environ.get("NIQ_RAYS_HOST_LOOP"):
        return _cast_rays_persistent(ctx, funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie)
    return _cast_rays_host_loop(ctx, funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie)


def _cast_rays_persistent(ctx, funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie):
    n = roots.shape[0]
    nf = len(funcs_tuple)
    mlps = [ctx.mlp(p) for p in params_tuple]
    handles = (C.c_void_p * nf)(*[m.handle for m in mlps])
    cfgs = (_niq.ModeCfg * nf)(*[_niq.mode_cfg(f.ctx) for f in funcs_tuple])
    o = _opts_struct(opts)
    t = np.zeros(n, np.float32)
    hit = np.zeros(n, np.int32)
    cnt = np.zeros(n, np.int32)
    tie = np.zeros(n, np.uint8) if return_near_tie else None      # the kernel skips the band bookkeeping when not asked
    n_evals = C.c_int64(0)
    _niq.check(_niq.lib().niq_cast_rays(ctx.handle, C.c_int32(nf), handles, cfgs, C.byref(o), C.c_int64(n),
                                        _niq.ptr(roots), _niq.ptr(dirs), _niq.ptr(t), _niq.ptr(hit), _niq.ptr(cnt),
                                        C.byref(n_evals), _niq.ptr(tie), C.c_int(_niq.MEM_HOST)))
    if return_near_tie:
        return t, hit, cnt, int(n_evals.value), tie.astype(bool)
    return t, hit, cnt, int(n_evals.value)


def _cast_rays_host_loop(ctx, funcs_tuple, params_tuple, roots, dirs, opts, return_near_tie):
    """src/queries.py:134-175 with cast_rays_iter (:39-132) evaluated through the batched GPU primitives."""
    import mlp
    f32 = np.float32
    N = roots.shape[0]
    n_substeps = int(opts['n_substeps'])
    hit_eps = f32(opts['hit_eps'])
    N_evals = 0
    out_t = np.zeros(N, f32)
    out_hit_id = np.zeros(N, np.int32)
    out_count = np.zeros(N, np.int32)
    out_tie = np.zeros(N, bool)
    cur = dict(roots=roots, dirs=dirs, t=np.zeros(N, f32),
               size=(np.ones(N, f32) * f32(opts['interval_init_size']) * f32(opts['max_dist'])).astype(f32),
               inds=np.arange(N, dtype=np.int64), count=np.zeros(N, np.int32), tie=np.zeros(N, bool))
    bucket = N                                   # the padded lane count the reference would be evaluating
    while cur['t'].shape[0] > 0:
        n = cur['t'].shape[0]
        is_hit = np.zeros(n, bool)
        hit_id = np.zeros(n, np.int32)
        n_inner = np.zeros(n, np.int32)
        t, size = cur['t'], cur['size']
        for _ in range(n_substeps):
            can_step = ~is_hit
            n_inner = n_inner + (~is_hit)
            for fid, (func, params) in enumerate(zip(funcs_tuple, params_tuple), start=1):
                pos_start = (cur['roots'] + t[:, None] * cur['dirs']).astype(f32)
                half_vec = (f32(0.5) * size[:, None] * cur['dirs']).astype(f32)
                pos_mid = (pos_start + half_vec).astype(f32)
                lab, _, _, btie = func.bound_general_box(params, pos_mid, half_vec[:, None, :], ctx=ctx)
                can_step &= (lab == 1) | (lab == 2)
                pos_eps = (cur['roots'] + (t + hit_eps)[:, None] * cur['dirs']).astype(f32)
                v, s = mlp.eval_points(params, np.concatenate((pos_start, pos_eps)), return_scale=True, ctx=ctx)
                v0, v1 = v[:n], v[n:]
                cur['tie'] |= btie | (np.abs(v0) <= f32(1e-5) * s[:n]) | (np.abs(v1) <= f32(1e-5) * s[n:])
                this_hit = np.sign(v0) != np.sign(v1)
                hit_id = np.where(this_hit, fid, hit_id).astype(np.int32)
                is_hit |= this_hit
            this_step = np.where(can_step, size, hit_eps).astype(f32)
            t = np.where(is_hit, t, t + this_step * f32(opts['safety_factor'])).astype(f32)
            size = np.where(can_step, size * f32(opts['interval_grow_fac']),
                            size * f32(opts['interval_shrink_fac'])).astype(f32)
            size = np.maximum(size, hit_eps)
        cur['t'], cur['size'] = t, size
        cur['count'] = cur['count'] + n_inner
        done = is_hit | (t > f32(opts['max_dist'])) | (cur['count'] >= opts['n_max_step'])
        w = cur['inds'][done]
        out_t[w] = t[done]
        out_hit_id[w] = hit_id[done]
        out_count[w] = cur['count'][done]
        out_tie[w] = cur['tie'][done]
        N_evals += bucket * n_substeps
        keep = ~done
        n_valid = int(keep.sum())
        if n_valid == 0:
            break
        if fits_in_smaller_bucket(n_valid, bucket):
            bucket = get_next_bucket_size(n_valid)
        cur = {k: a[keep] for k, a in cur.items()}
    if return_near_tie:
        return out_t, out_hit_id, out_count, N_evals, out_tie
    return out_t, out_hit_id, out_count, N_evals


def cast_rays_device(funcs_tuple, params_tuple, n, roots_ptr, dirs_ptr, t_ptr, hit_ptr, count_ptr, opts,
                     tie_ptr=None, want_n_evals=True, ctx=None):
    """cast_rays on DEVICE-resident buffers (raw device pointers as ints, e.g. torch `tensor.data_ptr()`):
    roots/dirs (n,3) f32 in, t (n) f32 / hit_id (n) i32 / count (n) i32 [/ near_tie (n) u8] out.
    interval / affine_fixed only.  Synchronous on return.  -> N_evals (0 when not requested)."""
    ctx = ctx or _niq.default_context()
    nf = len(funcs_tuple)
    mlps = [ctx.mlp(p) for p in params_tuple]
    handles = (C.c_void_p * nf)(*[m.handle for m in mlps])
    cfgs = (_niq.ModeCfg * nf)(*[_niq.mode_cfg(f.ctx) for f in funcs_tuple])
    o = _opts_struct(opts)
    n_evals = C.c_int64(0)
    vp = lambda p: None if p is None else C.c_void_p(int(p))
    _niq.check(_niq.lib().niq_cast_rays(ctx.handle, C.c_int32(nf), handles, cfgs, C.byref(o), C.c_int64(int(n)),
                                        vp(roots_ptr), vp(dirs_ptr), vp(t_ptr), vp(hit_ptr), vp(count_ptr),
                                        C.byref(n_evals) if want_n_evals else None, vp(tie_ptr),
                                        C.c_int(_niq.MEM_DEVICE)))
    return int(n_evals.value)


# ----------------------------------------------------------------------------------------------------
# cast_rays_frustum (src/queries.py:178-587)
# ----------------------------------------------------------------------------------------------------

def _frustum_cam(cam_params):
    root_pos, look_dir, up_dir, left_dir, fov_x, fov_y, res_x, res_y = cam_params
    f32 = np.float32
    res_x, res_y = int(res_x), int(res_y)
    if res_x < 1 or res_y < 1:
        raise ValueError("cast_rays_frustum: image resolution must be positive")
    cam = _niq.Camera()
    for k in range(3):
        cam.root[k] = f32(np.asarray(root_pos, f32)[k]); cam.look[k] = f32(np.asarray(look_dir, f32)[k])
        cam.up[k] = f32(np.asarray(up_dir, f32)[k]); cam.left[k] = f32(np.asarray(left_dir, f32)[k])
    # the transcendental constants of src/render.py:17-24 and src/queries.py:352-353, evaluated once in float32
    cam.tan_half_fov_x = np.tan(np.deg2rad(f32(fov_x)) / f32(2)).astype(f32)
    cam.tan_half_fov_y = np.tan(np.deg2rad(f32(fov_y)) / f32(2)).astype(f32)
    cam.half_fov_x = (np.deg2rad(f32(fov_x)) / f32(2)).astype(f32)
    cam.half_fov_y = (np.deg2rad(f32(fov_y)) / f32(2)).astype(f32)
    cam.res_x, cam.res_y = res_x, res_y
    return cam


def _initial_frusta(res_x, res_y, n_side):
    """src/queries.py:495-501: n_side x n_side tiles of pixels [x0, y0, x1, y1) (jnp.linspace(dtype=int) floors)."""
    def ticks(res):
        s = (np.arange(n_side, dtype=np.float32) / np.float32(n_side)).astype(np.float32)
        body = (np.float32(0) * (np.float32(1) - s)).astype(np.float32) + (np.float32(res) * s).astype(np.float32)
        return np.floor(np.concatenate((body, [np.float32(res)]))).astype(np.int32)
    xt, yt = ticks(res_x), ticks(res_y)
    return np.ascontiguousarray(np.stack((np.tile(xt[:-1], n_side), np.repeat(yt[:-1], n_side),
                                          np.tile(xt[1:], n_side), np.repeat(yt[1:], n_side)), axis=-1), np.int32)


def _frustum_n_evals(n_init, n_term, n_refine):
    """N_evals of the reference (src/queries.py:523-548): the padded array length of every marching iteration, replayed
    from the number of frusta that terminated / were split in each iteration."""
    size = empty_start = alive = n_init
    n_evals = 0
    for k in range(len(n_term)):
        n_evals += size
        n_valid = alive - int(n_term[k])
        if n_valid == 0:
            break
        n_ref = int(n_refine[k])
        new_bucket = get_next_bucket_size(n_valid + n_ref)
        if empty_start + n_ref > size or new_bucket < size:
            size, empty_start = new_bucket, n_valid
        empty_start += n_ref
        alive = n_valid + n_ref
    return n_evals


def cast_rays_frustum(funcs_tuple, params_tuple, cam_params, in_opts, return_near_tie=False, ctx=None, init_ranges=None,
                      iter_counts=None):
    """src/queries.py:465-587.  cam_params = (root_pos, look_dir, up_dir, left_dir, fov_x, fov_y, res_x, res_y)
    -> (out_t (res_x,res_y) f32, out_hit_id i32, out_count i32, N_evals[, near_tie bool]).
    Extras for the multi-GPU partition (sharding.cast_rays_frustum_sharded): init_ranges (k,4) int32 replaces the initial
    tiles (pixels outside them stay zero); iter_counts, a list, receives (terminated, split) of every iteration."""
    ctx = ctx or _niq.default_context()
    if isinstance(funcs_tuple, list): funcs_tuple = tuple(funcs_tuple)
    if isinstance(params_tuple, list): params_tuple = tuple(params_tuple)
    if len(funcs_tuple) != len(params_tuple) or len(funcs_tuple) < 1:
        raise ValueError("funcs_tuple and params_tuple must have the same (non-zero) length")
    n_side = int(in_opts['n_side_init'])
    res_x, res_y = int(cam_params[6]), int(cam_params[7])
    if n_side < 1 or n_side > min(res_x, res_y):
        raise ValueError("cast_rays_frustum: n_side_init must be in 1..min(res_x, res_y)")
    if init_ranges is None:
        init_ranges = _initial_frusta(res_x, res_y, n_side)
    init_ranges = np.ascontiguousarray(init_ranges, np.int32).reshape(-1, 4)
    modes = {f.ctx.mode for f in funcs_tuple}
    persistent = len(modes) == 1 and modes <= _PERSISTENT_RAY_MODES and not os.environ.get("NIQ_RAYS_HOST_LOOP")
    impl = _cast_rays_frustum_persistent if persistent else _cast_rays_frustum_host_loop
    return impl(ctx, funcs_tuple, params_tuple, cam_params, in_opts, return_near_tie, init_ranges, iter_counts)


def _cast_rays_frustum_persistent(ctx, funcs_tuple, params_tuple, cam_params, opts, return_near_tie, init=None, iter_counts=None):
    nf = len(funcs_tuple)
    mlps = [ctx.mlp(p) for p in params_tuple]
    handles = (C.c_void_p * nf)(*[m.handle for m in mlps])
    cfgs = (_niq.ModeCfg * nf)(*[_niq.mode_cfg(f.ctx) for f in funcs_tuple])
    o = _opts_struct(opts)
    cam = _frustum_cam(cam_params)
    if init is None:
        init = _initial_frusta(cam.res_x, cam.res_y, int(opts['n_side_init']))
    n = cam.res_x * cam.res_y
    n_bins = int(opts['n_max_step']) // int(opts['n_substeps']) + 3
    iters = np.zeros(2 * n_bins, np.int64)
    if init.shape[0] == 0:                                   # a rank without tiles (more ranks than tiles)
        z = np.zeros((cam.res_x, cam.res_y), np.float32)
        return (z, z.astype(np.int32), z.astype(np.int32), 0) + ((z.astype(bool),) if return_near_tie else ())
    t = np.zeros(n, np.float32)
    hit = np.zeros(n, np.int32)
    cnt = np.zeros(n, np.int32)
    tie = np.zeros(n, np.uint8)
    n_evals = C.c_int64(0)
    _niq.check(_niq.lib().niq_cast_rays_frustum(ctx.handle, C.c_int32(nf), handles, cfgs, C.byref(o), C.byref(cam),
                                                C.c_float(opts['refine_width_fac']), C.c_int64(init.shape[0]), _niq.ptr(init),
                                                _niq.ptr(t), _niq.ptr(hit), _niq.ptr(cnt), C.byref(n_evals), _niq.ptr(iters),
                                                _niq.ptr(tie), C.c_int(_niq.MEM_HOST)))
    if iter_counts is not None:
        last = int(np.nonzero(iters[:n_bins] + iters[n_bins:])[0].max()) + 1 if iters.any() else 0
        iter_counts.extend(zip(iters[:last].tolist(), iters[n_bins:n_bins + last].tolist()))
    shp = (cam.res_x, cam.res_y)
    if return_near_tie:
        return t.reshape(shp), hit.reshape(shp), cnt.reshape(shp), int(n_evals.value), tie.reshape(shp).astype(bool)
    return t.reshape(shp), hit.reshape(shp), cnt.reshape(shp), int(n_evals.value)


def _cast_rays_frustum_host_loop(ctx, funcs_tuple, params_tuple, cam_params, opts, return_near_tie, init=None, iter_counts=None):
    """The reference's host-level iteration over a compact list of live frusta (no padding entries; N_evals is replayed
    from the per-iteration counts), bounds and point values on the GPU.  Serves the modes without a persistent kernel
    (affine_all / affine_truncate / affine_append / sdf)."""
    import mlp
    f32 = np.float32
    cam = _frustum_cam(cam_params)
    res_x, res_y = cam.res_x, cam.res_y
    root = np.array(cam.root[:], f32); look = np.array(cam.look[:], f32)
    up = np.array(cam.up[:], f32); left = np.array(cam.left[:], f32)
    tan_x, tan_y = f32(cam.tan_half_fov_x), f32(cam.tan_half_fov_y)
    n_substeps = int(opts['n_substeps'])
    hit_eps = f32(opts['hit_eps'])

    def cam_ray(tx, ty):
        plane = (look[None, :] + left[None, :] * (tx * tan_x).astype(f32)[:, None]
                 + up[None, :] * (ty * tan_y).astype(f32)[:, None]).astype(f32)
        return (plane / np.sqrt((plane * plane).sum(axis=-1, keepdims=True, dtype=f32))).astype(f32)

    rng = _initial_frusta(res_x, res_y, int(opts['n_side_init'])) if init is None else np.array(init, np.int32).reshape(-1, 4)
    n_init = rng.shape[0]
    t = np.zeros(n_init, f32)
    size = (np.ones(n_init, f32) * f32(opts['interval_init_size']) * f32(opts['max_dist'])).astype(f32)
    count = np.zeros(n_init, f32)
    tie = np.zeros(n_init, bool)
    out_t = np.zeros((res_x, res_y), f32)
    out_hit = np.zeros((res_x, res_y), np.int32)
    out_count = np.zeros((res_x, res_y), np.int32)
    out_tie = np.zeros((res_x, res_y), bool)
    n_term, n_refine = [], []
    it = 0
    while rng.shape[0] > 0:
        n = rng.shape[0]
        x0, y0, x1, y1 = (rng[:, k] for k in range(4))
        single = (x0 + 1 == x1) & (y0 + 1 == y1)
        xc0 = ((f32(2) * x0.astype(f32)) / f32(res_x + 1.0) - f32(1)).astype(f32)
        xc1 = ((f32(2) * (x1 - 1).astype(f32)) / f32(res_x + 1.0) - f32(1)).astype(f32)
        yc0 = ((f32(2) * y0.astype(f32)) / f32(res_y + 1.0) - f32(1)).astype(f32)
        yc1 = ((f32(2) * (y1 - 1).astype(f32)) / f32(res_y + 1.0) - f32(1)).astype(f32)
        r_uu, r_lu, r_ul, r_ll = cam_ray(xc1, yc1), cam_ray(xc0, yc1), cam_ray(xc1, yc0), cam_ray(xc0, yc0)
        mid = (f32(0.5) * (r_uu + r_ll)).astype(f32)
        mid_len = np.sqrt((mid * mid).sum(axis=-1, dtype=f32)).astype(f32)
        mid = (mid / mid_len[:, None]).astype(f32)
        expand = (f32(1) / mid_len).astype(f32)
        is_hit = np.zeros(n, bool)
        hit_id = np.zeros(n, np.int32)
        n_inner = np.zeros(n, np.int32)
        demands = np.zeros(n, bool)
        for _ in range(n_substeps):
            t_adj = ((t + size).astype(f32) * expand).astype(f32)
            right_front = ((r_uu - r_lu) * t_adj[:, None] / f32(2)).astype(f32)
            up_front = ((r_uu - r_ul) * t_adj[:, None] / f32(2)).astype(f32)
            can_step = ~is_hit
            n_inner = n_inner + (~is_hit)
            center = (root[None, :] + (f32(0.5) * (t + t_adj))[:, None] * mid).astype(f32)
            cvec = ((f32(0.5) * (t_adj - t))[:, None] * mid).astype(f32)
            vecs = np.stack((cvec, right_front, up_front), axis=1)
            pos_start = (root[None, :] + t[:, None] * mid).astype(f32)
            pos_eps = (root[None, :] + (t + hit_eps)[:, None] * mid).astype(f32)
            for fid, (func, params) in enumerate(zip(funcs_tuple, params_tuple), start=1):
                lab, _, _, btie = func.bound_general_box(params, center, vecs, ctx=ctx)
                can_step &= (lab == 1) | (lab == 2)
                v, s = mlp.eval_points(params, np.concatenate((pos_start, pos_eps)), return_scale=True, ctx=ctx)
                v0, v1 = v[:n], v[n:]
                tie |= btie | (np.abs(v0) <= f32(1e-5) * s[:n]) | (np.abs(v1) <= f32(1e-5) * s[n:])
                this_hit = np.sign(v0) != np.sign(v1)
                hit_id = np.where(this_hit, fid, hit_id).astype(np.int32)
                is_hit |= this_hit
            this_step = np.where(can_step, size, hit_eps * single.astype(f32)).astype(f32)
            t = np.where(is_hit, t, t + this_step * f32(opts['safety_factor'])).astype(f32)
            size = np.where(can_step, size * f32(opts['interval_grow_fac']), size * f32(opts['interval_shrink_fac'])).astype(f32)
            demands |= (size < hit_eps) | is_hit
            size = np.maximum(size, hit_eps)
        area = (x1 - x0) * (y1 - y0)
        count = (count + (n_inner.astype(f32) * (f32(1.0) / area.astype(f32))).astype(f32)).astype(f32)
        done = (is_hit & (area == 1)) | (t > f32(opts['max_dist'])) | (it >= opts['n_max_step'])
        for i in np.nonzero(done)[0]:                      # a finished frustum fills its pixels (src/queries.py:558-577, 442-456)
            sl = (slice(x0[i], x1[i]), slice(y0[i], y1[i]))
            out_t[sl] = t[i]; out_hit[sl] = hit_id[i]; out_count[sl] = np.int32(count[i]); out_tie[sl] = tie[i]
        wx = (f32(2) * np.sin((f32(cam.half_fov_x) * (x1 - x0).astype(f32) / f32(res_x)).astype(f32)) * t).astype(f32)
        wy = (f32(2) * np.sin((f32(cam.half_fov_y) * (y1 - y0).astype(f32) / f32(res_y)).astype(f32)) * t).astype(f32)
        lim = (f32(opts['refine_width_fac']) * size).astype(f32)
        refine = ((wx > lim) | (wy > lim) | demands) & ((x1 > x0 + 1) | (y1 > y0 + 1)) & ~done
        n_term.append(int(done.sum()))
        n_refine.append(int(refine.sum()))
        it += n_substeps
        # split along the longer pixel axis, x on ties; the integer midpoint truncates (src/queries.py:371-432)
        sx = (x1 - x0) >= (y1 - y0)
        xm, ym = (x0 + x1) // 2, (y0 + y1) // 2
        a = rng.copy()
        a[:, 2] = np.where(refine & sx, xm, x1); a[:, 3] = np.where(refine & ~sx, ym, y1)
        b = rng[refine].copy()
        b[:, 0] = np.where(sx[refine], xm[refine], b[:, 0]); b[:, 1] = np.where(~sx[refine], ym[refine], b[:, 1])
        keep = ~done
        rng = np.concatenate((a[keep], b)).astype(np.int32)
        t = np.concatenate((t[keep], t[refine])); size = np.concatenate((size[keep], size[refine]))
        count = np.concatenate((count[keep], count[refine])); tie = np.concatenate((tie[keep], tie[refine]))
    N_evals = _frustum_n_evals(n_init, n_term, n_refine)
    if iter_counts is not None:
        iter_counts.extend(zip(n_term, n_refine))
    if return_near_tie:
        return out_t, out_hit, out_count, N_evals, out_tie
    return out_t, out_hit, out_count, N_evals
