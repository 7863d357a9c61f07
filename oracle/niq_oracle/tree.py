"""CPU oracle, part 3: the kd-tree queries.

TEST INFRASTRUCTURE ONLY (see net.py header).  Follows /root/reference/src/kd_tree.py:
  :19-218  construct_uniform_unknown_levelset_tree[_iter]
  :338-399 hierarchical_marching_cubes[_extract_iter]   (cell extraction itself is in mc.py)
  :402-655 find_any_intersection[_iter]
  :659-802 closest_point[_iter]
The reference's host-level control flow is kept verbatim (bucket sizes, 2048-node batches, the
[A...,B...] vs interleaved child orders, the global LIFO window) because output ORDER -- and for
closest_point the VALUES -- depend on it.
"""
import math

import numpy as np

from . import mc, net
from .rays import BUCKET_SIZES, compactify_and_rebucket, get_next_bucket_size

F32 = np.float32


def _argmax_first(x):
    return np.argmax(x, axis=-1)            # first index on ties, as jnp.argmax


def _split_children(lower, upper, split_dim):
    """kd_tree.py:69-78 (same code at :545-558 and :737-749)."""
    mid = (F32(0.5) * (lower + upper)).astype(F32)
    coord_mask = np.arange(3)[None, :] == split_dim[:, None]
    a_lo = lower
    a_hi = np.where(coord_mask, mid, upper).astype(F32)
    b_lo = np.where(coord_mask, mid, lower).astype(F32)
    b_hi = upper
    return a_lo, a_hi, b_lo, b_hi


def _interleave(a, b):
    """utils.py:173-176 for two arrays."""
    s = list(a.shape)
    s[0] *= 2
    return np.stack((a, b), axis=1).reshape(s)


# ----------------------------------------------------------------------------------------------
# level-set tree
# ----------------------------------------------------------------------------------------------

def construct_uniform_unknown_levelset_tree(ctx, params, lower, upper, node_terminate_thresh=None,
                                            split_depth=None, with_interior_nodes=False,
                                            with_exterior_nodes=False, offset=0.0, batch_process_size=2048,
                                            stats=None):
    """kd_tree.py:102-218.  `ctx` is a net.AffineContext.  Returns the same dict of padded arrays.
    `stats` (ours): optional dict receiving n_evals (valid nodes classified), n_near_tie, level sizes."""
    for b in BUCKET_SIZES:
        if b > batch_process_size and (b // batch_process_size) * batch_process_size != b:
            raise ValueError(f"batch_process_size must be a factor of our bucket sizes, is not a factor of {b} (try a power of 2)")
    if node_terminate_thresh is None and split_depth is None:
        raise ValueError("must specify at least one of node_terminate_thresh or split_depth as a terminating condition")
    if node_terminate_thresh is None:
        node_terminate_thresh = 9999999999

    B = batch_process_size
    lower = np.asarray(lower, F32)
    upper = np.asarray(upper, F32)
    node_lower = lower[None, :].copy()
    node_upper = upper[None, :].copy()
    node_valid = np.ones((1,), bool)
    N_curr = 1
    fin = {}
    for tag, want in (("interior", with_interior_nodes), ("exterior", with_exterior_nodes)):
        if want:
            fin[tag] = [np.zeros((B, 3), F32), np.zeros((B, 3), F32), 0]
    n_evals = 0
    n_near_tie = 0
    level_sizes = []

    n_splits = 99999999 if split_depth is None else split_depth + 1
    for i_split in range(n_splits):
        init_bucket = node_lower.shape[0]
        this_b = min(B, init_bucket)
        nb = init_bucket // this_b
        n_occ = int(math.ceil(N_curr / this_b))
        quit_next = (N_curr >= node_terminate_thresh) or (i_split + 1 == n_splits)
        cont = not quit_next
        level_sizes.append(N_curr)

        for tag in fin:                                                   # :156-164
            while fin[tag][0].shape[0] - fin[tag][2] < N_curr:
                for j in (0, 1):
                    a = fin[tag][j]
                    g = np.zeros((2 * a.shape[0], 3), F32)
                    g[:a.shape[0]] = a
                    fin[tag][j] = g

        v3 = node_valid.reshape(nb, this_b)
        l3 = node_lower.reshape(nb, this_b, 3)
        u3 = node_upper.reshape(nb, this_b, 3)
        out_valid = np.zeros((nb, 2 * this_b), bool)
        out_lower = np.zeros((nb, 2 * this_b, 3), F32)
        out_upper = np.zeros((nb, 2 * this_b, 3), F32)
        total_valid = 0
        for ib in range(n_occ):
            bv, bl, bu = v3[ib], l3[ib], u3[ib]
            lab, lo_b, up_b, sc_b = net.classify_box(params, ctx, bl, bu, offset, return_scale=True)
            n_evals += int(bv.sum())
            n_near_tie += int((net.bound_near_tie(lo_b, up_b, offset, sc_b, rel=net.tie_rel(params)) & bv).sum())
            split_dim = _argmax_first(bu - bl)
            for tag, sign in (("interior", net.SIGN_NEGATIVE), ("exterior", net.SIGN_POSITIVE)):
                if tag in fin:
                    m = bv & (lab == sign)
                    k = int(m.sum())
                    s = fin[tag][2]
                    fin[tag][0][s:s + k] = bl[m]
                    fin[tag][1][s:s + k] = bu[m]
                    fin[tag][2] = s + k
            split_mask = bv & (lab == net.SIGN_UNKNOWN)
            if cont:
                a_lo, a_hi, b_lo, b_hi = _split_children(bl, bu, split_dim)
                out_valid[ib, :] = np.concatenate((split_mask, split_mask))      # [A..., B...] :80-82
                out_lower[ib, :, :] = np.concatenate((a_lo, b_lo))
                out_upper[ib, :, :] = np.concatenate((a_hi, b_hi))
                total_valid += 2 * int(split_mask.sum())
            else:
                out_valid[ib, :this_b] = split_mask
                out_lower[ib, :this_b, :] = bl
                out_upper[ib, :this_b, :] = bu
                total_valid += int(split_mask.sum())

        node_valid = out_valid.reshape(-1)
        node_lower = out_lower.reshape(-1, 3)
        node_upper = out_upper.reshape(-1, 3)
        target = get_next_bucket_size(total_valid)
        node_valid, N_curr, node_lower, node_upper = compactify_and_rebucket(node_valid, target, node_lower, node_upper)
        if quit_next:
            break

    out = {"unknown_node_valid": node_valid, "unknown_node_lower": node_lower, "unknown_node_upper": node_upper}
    for tag in fin:
        out[f"{tag}_node_valid"] = np.arange(fin[tag][0].shape[0]) < fin[tag][2]
        out[f"{tag}_node_lower"] = fin[tag][0]
        out[f"{tag}_node_upper"] = fin[tag][1]
    if stats is not None:
        stats.update(n_evals=n_evals, n_near_tie=n_near_tie, level_sizes=level_sizes)
    return out


# ----------------------------------------------------------------------------------------------
# hierarchical marching cubes
# ----------------------------------------------------------------------------------------------

def hierarchical_marching_cubes(ctx, params, lower, upper, depth, n_subcell_depth=2,
                                extract_batch_max_tri_out=1000000):
    """kd_tree.py:357-399 -> (T,3,3) float32 triangle soup, order = node, subcell, slot."""
    tree = construct_uniform_unknown_levelset_tree(ctx, params, lower, upper,
                                                   split_depth=3 * (depth - n_subcell_depth))
    valid = tree["unknown_node_valid"]
    lo = tree["unknown_node_lower"][valid]
    hi = tree["unknown_node_upper"][valid]
    return mc.extract_mesh_from_leaves(params, lo, hi, n_subcell_depth)


# ----------------------------------------------------------------------------------------------
# find_any_intersection
# ----------------------------------------------------------------------------------------------

_SAMPLE_OFFSETS = np.concatenate((np.zeros((1, 3), F32), np.eye(3, dtype=F32), -np.eye(3, dtype=F32)), axis=0)


def _all_same_sign(vals):
    """utils.py:146-150 along the last axis."""
    return np.all(vals < 0, axis=-1) | np.all(vals > 0, axis=-1)


def _first_true(mask):
    """index of the first True along axis -1 (0 if none) and whether any -- jnp.nonzero(size=1, fill_value=0)."""
    idx = np.argmax(mask, axis=-1)
    has = np.take_along_axis(mask, idx[:, None], axis=-1)[:, 0]
    return idx, has


def find_any_intersection(ctx_tuple, params_tuple, lower, upper, eps, stats=None):
    """kd_tree.py:402-655 -> (found, idA, idB, loc).  Whole frontier each round, children interleaved."""
    if len(ctx_tuple) != 2:
        raise ValueError("intersection supports pairwise only as written")
    ctxA, ctxB = ctx_tuple
    pA, pB = params_tuple
    eps_cube_width = (F32(eps) / np.sqrt(F32(3))).astype(F32)
    node_lower = np.asarray(lower, F32)[None, :].copy()
    node_upper = np.asarray(upper, F32)[None, :].copy()
    N_curr = 1
    n_nodes = 0
    n_rounds = 0
    n_tie = 0
    while True:
        n_nodes += N_curr
        n_rounds += 1
        nb = node_lower.shape[0]
        valid = np.arange(nb) < N_curr
        lo, hi = node_lower, node_upper
        width = (hi - lo).max(axis=-1)
        split_dim = _argmax_first(hi - lo)
        is_small = width < eps_cube_width
        center = (F32(0.5) * (lo + hi)).astype(F32)
        pts = (center[:, None, :] + eps_cube_width * _SAMPLE_OFFSETS[None, :, :]).astype(F32)   # (nb,7,3)

        tA, loA, upA, scA = net.classify_box(pA, ctxA, lo, hi, return_scale=True)
        tB, loB, upB, scB = net.classify_box(pB, ctxB, lo, hi, return_scale=True)
        vA = net.eval_points(pA, pts.reshape(-1, 3)).reshape(nb, 7)
        vB = net.eval_points(pB, pts.reshape(-1, 3)).reshape(nb, 7)
        n_tie += int(((net.bound_near_tie(loA, upA, 0.0, scA, rel=net.tie_rel(pA)) | net.bound_near_tie(loB, upB, 0.0, scB, rel=net.tie_rel(pB))) & valid).sum())

        near_A = is_small & ~_all_same_sign(vA)
        near_B = is_small & ~_all_same_sign(vB)

        iA, anyA = _first_true(vA < 0)
        iB, anyB = _first_true(vB < 0)
        locA = np.take_along_axis(pts, iA[:, None, None], axis=1)[:, 0, :]
        locB = np.take_along_axis(pts, iB[:, None, None], axis=1)[:, 0, :]
        have_near = is_small & anyA & anyB
        found = have_near.copy()
        loc = np.full((nb, 3), F32(-777.0), F32)
        loc = np.where(have_near[:, None], (F32(0.5) * (locA + locB)).astype(F32), loc)

        both = (vA < 0) & (vB < 0)
        iT, anyT = _first_true(both)
        locT = np.take_along_axis(pts, iT[:, None, None], axis=1)[:, 0, :]
        found = found | anyT
        loc = np.where(anyT[:, None], locT, loc).astype(F32)

        insideA = (tA == net.SIGN_NEGATIVE) | ((tA == net.SIGN_UNKNOWN) & ~near_A)
        insideB = (tB == net.SIGN_NEGATIVE) | ((tB == net.SIGN_UNKNOWN) & ~near_B)
        needs = insideA & insideB & valid
        found = found & valid

        if found.any():
            i = int(np.argmax(found))
            if stats is not None:
                stats.update(n_nodes=n_nodes, n_rounds=n_rounds, n_near_tie=n_tie)
            return True, 1, 2, loc[i].copy()

        idx = np.nonzero(needs)[0]
        n_new = idx.shape[0]
        if n_new == 0:
            if stats is not None:
                stats.update(n_nodes=n_nodes, n_rounds=n_rounds, n_near_tie=n_tie)
            return False, 0, 0, np.array((-777.0, -777.0, -777.0), F32)
        a_lo, a_hi, b_lo, b_hi = _split_children(lo[idx], hi[idx], split_dim[idx])
        new_lo = _interleave(a_lo, b_lo)
        new_hi = _interleave(a_hi, b_hi)
        N_curr = 2 * n_new
        size = 2 * nb                                         # arrays double each round (:561-562)
        new_bucket = get_next_bucket_size(N_curr)
        if new_bucket < size:
            size = new_bucket                                 # :649-653
        node_lower = np.full((size, 3), F32(-777.0), F32)
        node_upper = np.full((size, 3), F32(-777.0), F32)
        node_lower[:N_curr] = new_lo
        node_upper[:N_curr] = new_hi


# ----------------------------------------------------------------------------------------------
# closest_point
# ----------------------------------------------------------------------------------------------

def _norm(x):
    return np.sqrt((x * x).sum(axis=-1, dtype=F32)).astype(F32)


def closest_point(ctx, params, lower, upper, query_points, eps=0.001, batch_process_size=2048, stats=None):
    """kd_tree.py:659-802 -> (query_min_dist (Q,), query_min_loc (Q,3)).

    Keeps the global LIFO stack and the window of `batch_process_size` entries popped per round;
    duplicate scatter targets resolve as XLA-CPU does (updates applied in index order: last wins)."""
    B = int(batch_process_size)
    query_points = np.ascontiguousarray(query_points, F32)
    Q = query_points.shape[0]
    lower = np.asarray(lower, F32)
    upper = np.asarray(upper, F32)
    w_lo = np.repeat(lower[None, :], Q, axis=0)
    w_hi = np.repeat(upper[None, :], Q, axis=0)
    w_id = np.arange(Q, dtype=np.int64)
    min_dist = np.full((Q,), np.inf, F32)
    min_loc = np.full((Q, 3), F32(-777.0), F32)
    top = Q
    eps_cube_width = (F32(eps) / np.sqrt(F32(3))).astype(F32)
    n_rounds = 0
    n_visits = 0
    max_top = top
    n_tie = 0
    tie_query = np.zeros((Q,), bool)          # ours: queries that visited a near-tie box (diagnostic, results unaffected)

    def grow(a, n_new):
        g = np.zeros((n_new,) + a.shape[1:], a.dtype)
        g[:a.shape[0]] = a
        return g

    while top > 0:
        while w_lo.shape[0] < top + B:                                     # :783-788
            n_new = max(2 * w_lo.shape[0], 8 * B)
            w_lo, w_hi, w_id = grow(w_lo, n_new), grow(w_hi, n_new), grow(w_id, n_new)

        pop = max(top - B, 0)
        b_id = w_id[pop:pop + B]
        lo = w_lo[pop:pop + B]
        hi = w_hi[pop:pop + B]
        q = query_points[b_id]
        q_min = min_dist[b_id]                                              # snapshot (:684)
        valid = np.arange(B) < top
        n_visits += int(valid.sum())
        top = pop

        ext = (hi - lo).astype(F32)
        width = ext.max(axis=-1)
        center = (F32(0.5) * (lo + hi)).astype(F32)
        center_off = np.sqrt((ext * ext).sum(axis=-1, dtype=F32)).astype(F32)
        d_center = _norm(q - center)
        max_dist_in_node = (d_center + center_off).astype(F32)
        split_dim = _argmax_first(ext)
        is_small = width < eps_cube_width
        pts = (center[:, None, :] + ext[:, None, :] * _SAMPLE_OFFSETS[None, :, :]).astype(F32)

        # The reference pushes all B lanes through the net; every use of a lane's label / sample values is masked by `valid`
        # (= the first n_valid lanes), so only those are evaluated here (B = 2^21 would otherwise cost 2 M boxes per round)
        n_valid = int(valid.sum())
        lab = np.zeros((B,), np.int32)
        vals = np.ones((B, 7), F32)
        lab_v, lo_b, up_b, sc_b = net.classify_box(params, ctx, lo[:n_valid], hi[:n_valid], return_scale=True)
        lab[:n_valid] = lab_v
        tie_b = np.zeros((B,), bool)
        tie_b[:n_valid] = net.bound_near_tie(lo_b, up_b, 0.0, sc_b, rel=net.tie_rel(params))
        n_tie += int(tie_b.sum())
        tie_query[b_id[tie_b]] = True
        is_outside = (lab == net.SIGN_NEGATIVE) | (lab == net.SIGN_POSITIVE)
        for s0 in range(0, n_valid, 1 << 16):
            s1 = min(s0 + (1 << 16), n_valid)
            vals[s0:s1] = net.eval_points(params, pts[s0:s1].reshape(-1, 3)).reshape(-1, 7)
        spans = ~_all_same_sign(vals) & valid
        this_dist = np.where(spans, max_dist_in_node, F32(np.inf)).astype(F32)
        needs = valid & ~is_outside & ~is_small & (d_center < q_min)

        np.minimum.at(min_dist, b_id, this_dist)                            # :725
        new_min = min_dist[b_id]
        has_new = this_dist == new_min
        tgt = b_id[has_new]
        min_loc[tgt] = center[has_new]                                      # last write wins (:727-729)

        idx = np.nonzero(needs)[0]
        n_new = idx.shape[0]
        a_lo, a_hi, b_lo, b_hi = _split_children(lo[idx], hi[idx], split_dim[idx])
        w_lo[pop:pop + 2 * n_new] = _interleave(a_lo, b_lo)
        w_hi[pop:pop + 2 * n_new] = _interleave(a_hi, b_hi)
        w_id[pop:pop + 2 * n_new] = _interleave(b_id[idx], b_id[idx])
        top = pop + 2 * n_new
        max_top = max(max_top, top)
        n_rounds += 1

    if stats is not None:
        stats.update(n_rounds=n_rounds, n_visits=n_visits, max_stack=max_top, n_near_tie=n_tie, tie_query=tie_query)
    return min_dist, min_loc


# ----------------------------------------------------------------------------------------------
# tree consumers: surface sampling and bulk properties (kd_tree.py:220-292, 804-863)
# The reference draws from jax.random (threefry), which is not available here; these restatements take a
# numpy.random.Generator and draw the SAME quantities in the SAME order (node index, then position), so the product
# and the oracle agree sample by sample, while agreement with the reference is distributional.
# ----------------------------------------------------------------------------------------------

def _draw_in_nodes(rng, node_lower, node_upper, n):
    """kd_tree.py:230-240 / 810-820: a node per sample (uniform over the valid nodes), then a point inside it."""
    node_ind = rng.integers(0, node_lower.shape[0], size=n)
    u = rng.random((n, 3), dtype=np.float32)
    lo, hi = node_lower[node_ind], node_upper[node_ind]
    return (lo + u * (hi - lo)).astype(F32)


def sample_surface(ctx, params, lower, upper, n_samples, width, rng, n_node_thresh=4096):
    """kd_tree.py:253-292: tree with offset=width, then rejection sampling |f| < width inside the unknown leaves."""
    out = construct_uniform_unknown_levelset_tree(ctx, params, lower, upper, node_terminate_thresh=n_node_thresh, offset=width)
    v = out["unknown_node_valid"]
    nl, nu = out["unknown_node_lower"][v], out["unknown_node_upper"][v]
    per_round = min(3 * n_samples, 100000)
    found = np.zeros((n_samples, 3), F32)
    n_found = 0
    while n_found < n_samples:
        pos = _draw_in_nodes(rng, nl, nu, per_round)
        ok = np.abs(net.eval_points(params, pos)) < F32(width)
        take = pos[ok][: n_samples - n_found]
        found[n_found:n_found + take.shape[0]] = take
        n_found += take.shape[0]
    return found


def bulk_properties(ctx, params, lower, upper, rng, n_expand=int(1e4), n_sample=int(1e6)):
    """kd_tree.py:837-863: mass and centroid of {f < 0}: exact over the interior (NEGATIVE) nodes, Monte Carlo over the
    unknown leaves (:804-835)."""
    out = construct_uniform_unknown_levelset_tree(ctx, params, lower, upper, with_interior_nodes=True, node_terminate_thresh=n_expand)
    v, iv = out["unknown_node_valid"], out["interior_node_valid"]
    nl, nu = out["unknown_node_lower"][v], out["unknown_node_upper"][v]
    il, iu = out["interior_node_lower"][iv], out["interior_node_upper"][iv]
    m_int = np.prod(iu - il, axis=-1, dtype=F32)
    mass_interior = m_int.sum(dtype=F32)
    centroid_interior = (m_int[:, None] * (F32(0.5) * (il + iu))).sum(axis=0, dtype=F32)
    pos = _draw_in_nodes(rng, nl, nu, n_sample)
    inside = net.eval_points(params, pos) < 0
    vol_per_sample = np.prod(nu - nl, axis=-1, dtype=F32).sum(dtype=F32) / F32(n_sample)
    mass_boundary = vol_per_sample * F32(inside.sum())
    centroid_boundary = vol_per_sample * np.where(inside[:, None], pos, F32(0)).sum(axis=0, dtype=F32)
    mass = mass_interior + mass_boundary
    return F32(mass), ((centroid_interior + centroid_boundary) / mass).astype(F32)
