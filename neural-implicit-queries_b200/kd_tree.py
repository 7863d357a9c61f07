"""`kd_tree` module surface of the reference for the hot path (/root/reference/src/kd_tree.py):
  construct_uniform_unknown_levelset_tree :102-218, hierarchical_marching_cubes :357-399,
  find_any_intersection :567-655, closest_point :765-802.
Same names, arguments, return shapes / dtypes and error behaviour; the jitted *_iter bodies and the
Python-level batch loops are replaced by the CUDA drivers behind include/niq.h."""
import ctypes as C

import numpy as np

import _niq
import extract_cell
from bucketing import bucket_sizes, get_next_bucket_size
from implicit_function import SIGN_NEGATIVE, SIGN_POSITIVE, SIGN_UNKNOWN  # noqa: F401

INVALID_IND = 2 ** 30


def _vec3(x, name):
    a = np.ascontiguousarray(x, np.float32)
    if a.shape != (3,):
        raise ValueError(f"{name} must have shape (3,)")
    return a


class _Tree:
    """Owns a niq_tree handle (device-resident node lists)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def close(self):
        if self.handle:
            _niq.lib().niq_tree_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, which):
        n = C.c_int64()
        _niq.check(_niq.lib().niq_tree_count(self.handle, C.c_int(which), C.byref(n)))
        return n.value

    def nodes(self, which):
        n = self.count(which)
        lo = np.zeros((n, 3), np.float32)
        hi = np.zeros((n, 3), np.float32)
        if n:
            _niq.check(_niq.lib().niq_tree_copy(self.handle, C.c_int(which), _niq.ptr(lo), _niq.ptr(hi), C.c_int64(n),
                                                C.c_int(_niq.MEM_HOST)))
        return lo, hi

    def stats(self):
        s = (C.c_int64 * 4)()
        _niq.check(_niq.lib().niq_tree_stats(self.handle, s))
        return {"n_evals": s[0], "n_near_tie": s[1], "n_levels": s[2], "max_frontier": s[3]}

    def level_info(self):
        out = []
        for lv in range(self.stats()["n_levels"]):
            info = (C.c_int64 * 4)()
            _niq.check(_niq.lib().niq_tree_level_info(self.handle, C.c_int32(lv), info))
            out.append(tuple(info))
        return out


def build_tree(func, params, lower, upper, node_terminate_thresh=None, split_depth=None, with_interior_nodes=False,
               with_exterior_nodes=False, offset=0., batch_process_size=2048, ctx=None):
    """The device-resident tree (ours): what construct_uniform_unknown_levelset_tree wraps.  lower / upper may be
    (3,) -- the reference's single root box -- or (n,3): n root boxes refined in one level-synchronous build (the unit
    of the multi-GPU subtree partition, sharding.tree_sharded)."""
    ctx = ctx or _niq.default_context()
    lower = np.ascontiguousarray(lower, np.float32)
    upper = np.ascontiguousarray(upper, np.float32)
    if lower.ndim == 1:
        lower, upper = _vec3(lower, "lower")[None], _vec3(upper, "upper")[None]
    if lower.ndim != 2 or lower.shape[1] != 3 or lower.shape != upper.shape or lower.shape[0] < 1:
        raise ValueError("lower / upper must have shape (3,) or (n,3)")
    flags = (_niq.TREE_INTERIOR if with_interior_nodes else 0) | (_niq.TREE_EXTERIOR if with_exterior_nodes else 0)
    cfg = _niq.mode_cfg(func.ctx)
    m = ctx.mlp(params)
    h = C.c_void_p()
    _niq.check(_niq.lib().niq_tree_build_roots(
        ctx.handle, m.handle, C.byref(cfg), C.c_int64(lower.shape[0]), _niq.ptr(lower), _niq.ptr(upper),
        C.c_int32(-1 if split_depth is None else int(split_depth)),
        C.c_int64(0 if node_terminate_thresh is None else int(node_terminate_thresh)), C.c_float(offset),
        C.c_int32(flags), C.c_int32(int(batch_process_size)), C.byref(h)))
    return _Tree(ctx, h)


def build_tree_dealt(func, params, lower, upper, split_depth, deal_depth, rank, world, offset=0., batch_process_size=2048, ctx=None):
    """Ours: this rank's share of the tree at `split_depth` when the subtrees below `deal_depth` are partitioned round-robin
    over `world` ranks (niq_tree_build_dealt: replicated top + own subtrees in ONE persistent launch).  Fixed-row modes and
    slope_interval; raises NiqError(NIQ_EUNSUPPORTED) otherwise (sharding.tree_sharded then deals on the host)."""
    ctx = ctx or _niq.default_context()
    lower, upper = _vec3(lower, "lower"), _vec3(upper, "upper")
    cfg = _niq.mode_cfg(func.ctx)
    m = ctx.mlp(params)
    h = C.c_void_p()
    _niq.check(_niq.lib().niq_tree_build_dealt(
        ctx.handle, m.handle, C.byref(cfg), _niq.ptr(lower), _niq.ptr(upper), C.c_int32(int(split_depth)), C.c_float(offset),
        C.c_int32(int(batch_process_size)), C.c_int32(int(deal_depth)), C.c_int32(int(rank)), C.c_int32(int(world)), C.byref(h)))
    return _Tree(ctx, h)


def _padded(lo, hi, size):
    n = lo.shape[0]
    out_lo = np.zeros((size, 3), np.float32)
    out_hi = np.zeros((size, 3), np.float32)
    out_lo[:n], out_hi[:n] = lo, hi
    return np.arange(size) < n, out_lo, out_hi


def construct_uniform_unknown_levelset_tree(func, params, lower, upper, node_terminate_thresh=None, split_depth=None,
                                            compress_after=False, with_childern=False, with_interior_nodes=False,
                                            with_exterior_nodes=False, offset=0., batch_process_size=2048,
                                            stats=None, ctx=None):
    """src/kd_tree.py:102-218.  Returns the same dict of padded arrays (`*_valid` masks are the contract;
    padding rows are zero).  Node ORDER follows the reference.  `stats` (ours): optional dict receiving
    n_evals / n_near_tie / n_levels / max_frontier."""
    for b in bucket_sizes:
        if b > batch_process_size and (b // batch_process_size) * batch_process_size != b:
            raise ValueError(f"batch_process_size must be a factor of our bucket sizes, is not a factor of {b} (try a power of 2)")
    if node_terminate_thresh is None and split_depth is None:
        raise ValueError("must specify at least one of node_terminate_thresh or split_depth as a terminating condition")
    tree = build_tree(func, params, _vec3(lower, "lower"), _vec3(upper, "upper"), node_terminate_thresh, split_depth,
                      with_interior_nodes, with_exterior_nodes, offset, batch_process_size, ctx)
    try:
        lo, hi = tree.nodes(0)
        valid, plo, phi = _padded(lo, hi, get_next_bucket_size(lo.shape[0]))
        out = {'unknown_node_valid': valid, 'unknown_node_lower': plo, 'unknown_node_upper': phi}
        levels = tree.level_info()
        for tag, which, col, want in (("interior", 1, 2, with_interior_nodes), ("exterior", 2, 3, with_exterior_nodes)):
            if not want:
                continue
            # replay the reference's lazy doubling (:156-164): before each level the array must have room
            # for every node entering the level
            size, n_fin = batch_process_size, 0
            for info in levels:
                while size - n_fin < info[0]:
                    size *= 2
                n_fin += info[col]
            flo, fhi = tree.nodes(which)
            v, a, b = _padded(flo, fhi, size)
            out[f'{tag}_node_valid'], out[f'{tag}_node_lower'], out[f'{tag}_node_upper'] = v, a, b
        if stats is not None:
            stats.update(tree.stats())
            stats['level_sizes'] = [lv[0] for lv in levels]
        return out
    finally:
        tree.close()


def hierarchical_marching_cubes(func, params, lower, upper, depth, n_subcell_depth=2, extract_batch_max_tri_out=1000000,
                                ctx=None):
    """src/kd_tree.py:357-399 -> (T,3,3) float32.  Tree leaves never leave the GPU: the MC kernels read the
    tree's device-resident leaf list.  `extract_batch_max_tri_out` only sized the reference's batches."""
    ctx = ctx or _niq.default_context()
    tree = build_tree(func, params, lower, upper, split_depth=3 * (depth - n_subcell_depth), ctx=ctx)
    try:
        mesh = C.c_void_p()
        m = ctx.mlp(params)
        _niq.check(_niq.lib().niq_marching_cubes_tree(ctx.handle, m.handle, tree.handle, C.c_int32(n_subcell_depth),
                                                      C.byref(mesh)))
        try:
            return extract_cell._mesh_to_numpy(mesh)
        finally:
            _niq.lib().niq_mesh_destroy(mesh)
    finally:
        tree.close()


def find_any_intersection(func_tuple, params_tuple, lower, upper, eps, viz_nodes=False, stats=None, ctx=None):
    """src/kd_tree.py:567-655 -> (found, 1|0, 2|0, loc (3,)).  `viz_nodes` is GUI-only and unsupported."""
    if len(func_tuple) != 2 or len(params_tuple) != 2:
        raise ValueError("intersection supports pairwise only as written")
    if viz_nodes:
        raise _niq.NiqError(_niq.NIQ_EUNSUPPORTED, "viz_nodes=True is a GUI debugging aid outside this backend")
    ctx = ctx or _niq.default_context()
    lower, upper = _vec3(lower, "lower"), _vec3(upper, "upper")
    mA, mB = ctx.mlp(params_tuple[0]), ctx.mlp(params_tuple[1])
    cA, cB = _niq.mode_cfg(func_tuple[0].ctx), _niq.mode_cfg(func_tuple[1].ctx)
    found = C.c_int32(0)
    loc = np.zeros(3, np.float32)
    st = (C.c_int64 * 3)()
    _niq.check(_niq.lib().niq_find_any_intersection(ctx.handle, mA.handle, C.byref(cA), mB.handle, C.byref(cB),
                                                    _niq.ptr(lower), _niq.ptr(upper), C.c_float(eps), C.byref(found),
                                                    _niq.ptr(loc), st))
    if stats is not None:
        stats.update(n_nodes=st[0], n_rounds=st[1], n_near_tie=st[2])
    if found.value:
        return True, 1, 2, loc
    return False, 0, 0, np.array((-777., -777., -777.), np.float32)


def find_any_intersection_batch(func_tuple, params_tuple, lower, upper, eps, R_B=None, t_B=None, R_A=None, t_A=None, stats=None,
                                ctx=None):
    """Ours: a BATCH of find_any_intersection queries that differ in the rigid transforms of the shapes -- the values the
    reference's GUI writes into params["0000.spatial_transformation.R" / ".t"] before every call
    (src/main_intersection.py:171-183).  R_* (n,3,3), t_* (n,3); a shape that receives transforms must have a
    spatial_transformation as its first op (mlp.prepend_op).  The growing-form modes run all queries in ONE persistent kernel
    (niq_find_any_intersection_batch); other modes run one call per query.  -> (found (n,) bool, loc (n,3) f32); query i
    equals find_any_intersection with R[i], t[i] stored in the params."""
    if len(func_tuple) != 2 or len(params_tuple) != 2:
        raise ValueError("intersection supports pairwise only as written")
    ctx = ctx or _niq.default_context()
    lower, upper = _vec3(lower, "lower"), _vec3(upper, "upper")

    def pack(R, t):
        if R is None and t is None:
            return None
        R = np.ascontiguousarray(R, np.float32).reshape(-1, 9)
        t = np.ascontiguousarray(t, np.float32).reshape(-1, 3)
        if R.shape[0] != t.shape[0]:
            raise ValueError("R and t must hold one transform per query")
        return np.ascontiguousarray(np.concatenate((R, t), axis=1))
    xA, xB = pack(R_A, t_A), pack(R_B, t_B)
    if xA is None and xB is None:
        raise ValueError("give the transforms of at least one shape")
    n = (xA if xA is not None else xB).shape[0]
    if xA is not None and xB is not None and xA.shape[0] != xB.shape[0]:
        raise ValueError("both shapes need the same number of transforms")
    modes = {f.ctx.mode for f in func_tuple}
    found = np.zeros(n, np.int32)
    loc = np.full((n, 3), -777., np.float32)
    st = np.zeros((n, 3), np.int64)
    if modes <= {"affine_truncate", "affine_all", "affine_append"}:
        mA, mB = ctx.mlp(params_tuple[0]), ctx.mlp(params_tuple[1])
        cA, cB = _niq.mode_cfg(func_tuple[0].ctx), _niq.mode_cfg(func_tuple[1].ctx)
        _niq.check(_niq.lib().niq_find_any_intersection_batch(ctx.handle, mA.handle, C.byref(cA), mB.handle, C.byref(cB), C.c_int64(n),
                                                              _niq.ptr(xA), _niq.ptr(xB), _niq.ptr(lower), _niq.ptr(upper),
                                                              C.c_float(eps), _niq.ptr(found), _niq.ptr(loc), _niq.ptr(st)))
    else:
        pA, pB = dict(params_tuple[0]), dict(params_tuple[1])
        for i in range(n):
            for p, x in ((pA, xA), (pB, xB)):
                if x is not None:
                    p["0000.spatial_transformation.R"] = x[i, :9].reshape(3, 3)
                    p["0000.spatial_transformation.t"] = x[i, 9:]
            s1 = {}
            f, _, _, l = find_any_intersection(func_tuple, (pA, pB), lower, upper, eps, stats=s1, ctx=ctx)
            found[i], loc[i] = int(bool(f)), l
            st[i] = (s1["n_nodes"], s1["n_rounds"], s1["n_near_tie"])
    if stats is not None:
        stats.update(n_nodes=st[:, 0].copy(), n_rounds=st[:, 1].copy(), n_near_tie=st[:, 2].copy())
    return found.astype(bool), loc


def closest_point(func, params, lower, upper, query_points, eps=0.001, batch_process_size=2048, stats=None, ctx=None):
    """src/kd_tree.py:765-802 -> (query_min_dist (Q,), query_min_loc (Q,3)).  Results depend on
    `batch_process_size` exactly as in the reference (global LIFO window)."""
    ctx = ctx or _niq.default_context()
    lower, upper = _vec3(lower, "lower"), _vec3(upper, "upper")
    q = np.ascontiguousarray(query_points, np.float32)
    if q.ndim != 2 or q.shape[1] != 3:
        raise ValueError("query_points must have shape (Q,3)")
    Q = q.shape[0]
    dist = np.full(Q, np.inf, np.float32)
    loc = np.full((Q, 3), -777., np.float32)
    st = (C.c_int64 * 4)()
    cfg = _niq.mode_cfg(func.ctx)
    m = ctx.mlp(params)
    _niq.check(_niq.lib().niq_closest_point(ctx.handle, m.handle, C.byref(cfg), _niq.ptr(lower), _niq.ptr(upper),
                                            C.c_int64(Q), _niq.ptr(q), C.c_float(eps), C.c_int64(int(batch_process_size)),
                                            _niq.ptr(dist), _niq.ptr(loc), st, C.c_int(_niq.MEM_HOST)))
    if stats is not None:
        stats.update(n_rounds=st[0], n_visits=st[1], max_stack=st[2], n_near_tie=st[3])
    return dist, loc


# ---------------------------------------------------------------------------------------------------
# tree consumers: src/kd_tree.py:220-292 sample_surface, :804-863 bulk_properties.
# The tree and every function evaluation run on the GPU; the draws come from a numpy.random.Generator (the reference's
# jax.random / threefry streams are not reproducible without JAX: same quantities in the same order -- a node per
# sample, then a point inside it -- hence the same distribution, not the same bits).  `rngkey` may be an int seed or a
# Generator.
# ---------------------------------------------------------------------------------------------------

def _rng(rngkey):
    return rngkey if isinstance(rngkey, np.random.Generator) else np.random.default_rng(rngkey)


def _draw_in_nodes(rng, node_lower, node_upper, n):
    node_ind = rng.integers(0, node_lower.shape[0], size=n)
    u = rng.random((n, 3), dtype=np.float32)
    lo, hi = node_lower[node_ind], node_upper[node_ind]
    return (lo + u * (hi - lo)).astype(np.float32)


def sample_surface(func, params, lower, upper, n_samples, width, rngkey, n_node_thresh=4096, ctx=None):
    """src/kd_tree.py:253-292 -> (n_samples,3) points with |f| < width, drawn uniformly from the unknown leaves of a tree
    built with offset=width (so the band is covered)."""
    rng = _rng(rngkey)
    out = construct_uniform_unknown_levelset_tree(func, params, lower, upper, node_terminate_thresh=n_node_thresh,
                                                  offset=width, ctx=ctx)
    v = out['unknown_node_valid']
    nl, nu = out['unknown_node_lower'][v], out['unknown_node_upper'][v]
    if nl.shape[0] == 0:
        raise ValueError("no unknown node: the level set does not cross the domain")
    per_round = min(3 * n_samples, 100000)
    found = np.zeros((n_samples, 3), np.float32)
    n_found = 0
    while n_found < n_samples:
        pos = _draw_in_nodes(rng, nl, nu, per_round)
        ok = np.abs(func(params, pos)) < np.float32(width)
        take = pos[ok][:n_samples - n_found]
        found[n_found:n_found + take.shape[0]] = take
        n_found += take.shape[0]
    return found


def sample_surface_uniform(func, params, lower, upper, n_samples, width, rngkey, ctx=None):
    """src/kd_tree.py:296-336: the tree-free baseline of sample_surface -- uniform draws in [lower, upper], keep |f| < width.
    Point evaluations on the GPU, draws from a numpy Generator (jax.random streams are not reproducible without JAX)."""
    rng = _rng(rngkey)
    lower, upper = _vec3(lower, "lower"), _vec3(upper, "upper")
    per_round = min(10 * n_samples, 100000)
    found = np.zeros((n_samples, 3), np.float32)
    n_found = 0
    while n_found < n_samples:
        pos = rng.uniform(lower, upper, (per_round, 3)).astype(np.float32)
        import mlp
        ok = np.abs(mlp.eval_points(params, pos, ctx=ctx)) < np.float32(width)
        take = pos[ok][:n_samples - n_found]
        found[n_found:n_found + take.shape[0]] = take
        n_found += take.shape[0]
    return found


def bulk_properties(func, params, lower, upper, rngkey, n_expand=int(1e4), n_sample=int(1e6), ctx=None):
    """src/kd_tree.py:837-863 -> (mass, centroid (3,)) of {f < 0}: exact over the interior nodes, Monte Carlo over the
    unknown leaves (:804-835)."""
    rng = _rng(rngkey)
    f32 = np.float32
    out = construct_uniform_unknown_levelset_tree(func, params, lower, upper, with_interior_nodes=True,
                                                  node_terminate_thresh=n_expand, ctx=ctx)
    v, iv = out['unknown_node_valid'], out['interior_node_valid']
    nl, nu = out['unknown_node_lower'][v], out['unknown_node_upper'][v]
    il, iu = out['interior_node_lower'][iv], out['interior_node_upper'][iv]
    m_int = np.prod(iu - il, axis=-1, dtype=f32)
    mass_interior = m_int.sum(dtype=f32)
    centroid_interior = (m_int[:, None] * (f32(0.5) * (il + iu))).sum(axis=0, dtype=f32)
    pos = _draw_in_nodes(rng, nl, nu, n_sample)
    inside = func(params, pos) < 0
    vol_per_sample = np.prod(nu - nl, axis=-1, dtype=f32).sum(dtype=f32) / f32(n_sample)
    mass_boundary = vol_per_sample * f32(inside.sum())
    centroid_boundary = vol_per_sample * np.where(inside[:, None], pos, f32(0)).sum(axis=0, dtype=f32)
    mass = mass_interior + mass_boundary
    return f32(mass), ((centroid_interior + centroid_boundary) / mass).astype(f32)
