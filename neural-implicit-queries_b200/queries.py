"""`queries` module surface of the reference for the hot path: get_default_cast_opts and cast_rays
(/root/reference/src/queries.py:23-36, :39-175).

interval / affine_fixed / slope_interval: ONE persistent CUDA kernel (csrc/niq_kernels.cuh k_cast_rays) marches every ray
to termination with an in-kernel work queue -- no per-iteration host round trip, no bucket padding.
affine_all / affine_truncate: the reference's host-level iteration (one pass per step, order-preserving
compaction) with the bound / point evaluations on the GPU (one CTA per ray segment, niq_grow.cuh)."""
import ctypes as C

import numpy as np

import _niq
from bucketing import fits_in_smaller_bucket, get_next_bucket_size


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
    if len(modes) == 1 and modes <= {"interval", "affine_fixed", "slope_interval"}:
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
    tie = np.zeros(n, np.uint8)
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
