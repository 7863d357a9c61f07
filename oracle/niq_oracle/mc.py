"""CPU oracle, part 4: marching-cubes extraction over tree leaves.

TEST INFRASTRUCTURE ONLY (see net.py header).  Follows /root/reference/src/extract_cell.py:306-421
(table lookup, edge interpolation, subcell lattice) and kd_tree.py:338-355 (compaction of the valid
triangles into a soup ordered by node, subcell ('ij' order, axis 0 slowest), slot).
"""
import numpy as np

from . import mc_tables, net

F32 = np.float32

_TRI_TABLE, _EDGE_VERTS, _VERT_COORDS = mc_tables.unpack()


def get_mc_data():
    """extract_cell.py:306-310"""
    return _TRI_TABLE, _EDGE_VERTS, _VERT_COORDS


def lattice_points(cell_lower, cell_upper, n_sub_depth):
    """extract_cell.py:372-381: (L,3)x2 -> (L, P, P, P, 3) lattice with P = 2^n + 1, linspace per axis."""
    n_pts = 1 + 2 ** n_sub_depth
    lo = np.asarray(cell_lower, F32)
    hi = np.asarray(cell_upper, F32)
    # jnp.linspace in float32 (jax/_src/numpy/lax_numpy.py, `linspace`): for i < num-1 the sample is
    # start*(1 - i/div) + stop*(i/div) with div = num-1, and the last sample is `stop` itself.
    div = F32(n_pts - 1)
    step = (np.arange(n_pts - 1, dtype=F32) / div).astype(F32)
    side = np.empty((lo.shape[0], n_pts, 3), F32)
    side[:, :-1, :] = (lo[:, None, :] * (F32(1) - step)[None, :, None]).astype(F32) \
        + (hi[:, None, :] * step[None, :, None]).astype(F32)
    side[:, -1, :] = hi
    L = lo.shape[0]
    grid = np.empty((L, n_pts, n_pts, n_pts, 3), F32)
    grid[..., 0] = side[:, :, None, None, 0]
    grid[..., 1] = side[:, None, :, None, 1]
    grid[..., 2] = side[:, None, None, :, 2]
    return grid


def triangles_from_cells(cell_lower, cell_upper, vert_vals):
    """extract_cell.py:314-364 for C cells: (C,3),(C,3),(C,8) -> tri_pos (C,5,3,3), tri_valid (C,5)."""
    lo = np.asarray(cell_lower, F32)
    hi = np.asarray(cell_upper, F32)
    vv = np.asarray(vert_vals, F32)
    vert_pos = np.where(_VERT_COORDS[None, :, :], hi[:, None, :], lo[:, None, :]).astype(F32)     # (C,8,3)
    ia, ib = _EDGE_VERTS[:, 0], _EDGE_VERTS[:, 1]
    valA, valB = vv[:, ia], vv[:, ib]                                                              # (C,12)
    posA, posB = vert_pos[:, ia, :], vert_pos[:, ib, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -valA / (valB - valA)
    t = np.nan_to_num(t).astype(F32)
    t = np.clip(t, F32(0), F32(1))
    cross = ((F32(1) - t)[:, :, None] * posA + t[:, :, None] * posB).astype(F32)                   # (C,12,3)
    case_id = ((vv < 0) * (2 ** np.arange(8))[None, :]).sum(axis=-1)
    tris = _TRI_TABLE[case_id, :15].reshape(-1, 5, 3)                                              # (C,5,3)
    valid = tris[:, :, 0] != -1
    idx = np.clip(tris, 0, None)
    tri_pos = cross[np.arange(cross.shape[0])[:, None, None], idx, :]                              # (C,5,3,3)
    return tri_pos.astype(F32), valid


def extract_triangles_from_subcells(params, n_sub_depth, cell_lower, cell_upper):
    """extract_cell.py:366-421 batched over L leaves -> tri_pos (L, S*5, 3, 3), tri_valid (L, S*5), S = 8^n."""
    lo = np.asarray(cell_lower, F32)
    hi = np.asarray(cell_upper, F32)
    L = lo.shape[0]
    n_side = 2 ** n_sub_depth
    n_pts = n_side + 1
    grid = lattice_points(lo, hi, n_sub_depth)
    vals = net.eval_points(params, grid.reshape(-1, 3)).reshape(L, n_pts, n_pts, n_pts)

    ii = np.arange(n_side)
    g0, g1, g2 = np.meshgrid(ii, ii, ii, indexing="ij")
    sub_inds = np.stack((g0, g1, g2), axis=-1).reshape(-1, 3)                        # (S,3)
    delta = ((hi - lo) / F32(n_side)).astype(F32)                                     # (L,3)
    sub_lo = (lo[:, None, :] + sub_inds[None, :, :].astype(F32) * delta[:, None, :]).astype(F32)
    sub_hi = (sub_lo + delta[:, None, :]).astype(F32)
    vinds = sub_inds[:, None, :] + _VERT_COORDS[None, :, :].astype(np.int64)          # (S,8,3)
    vvals = vals[:, vinds[:, :, 0], vinds[:, :, 1], vinds[:, :, 2]]                   # (L,S,8)
    S = sub_inds.shape[0]
    tri_pos, tri_valid = triangles_from_cells(sub_lo.reshape(-1, 3), sub_hi.reshape(-1, 3), vvals.reshape(-1, 8))
    return tri_pos.reshape(L, S * 5, 3, 3), tri_valid.reshape(L, S * 5)


def extract_mesh_from_leaves(params, leaf_lower, leaf_upper, n_sub_depth, chunk=256):
    """kd_tree.py:338-355,385-397 over the valid leaves -> (T,3,3)."""
    out = []
    for s in range(0, leaf_lower.shape[0], chunk):
        tp, tv = extract_triangles_from_subcells(params, n_sub_depth, leaf_lower[s:s + chunk], leaf_upper[s:s + chunk])
        out.append(tp.reshape(-1, 3, 3)[tv.reshape(-1)])
    if not out:
        return np.zeros((0, 3, 3), F32)
    return np.concatenate(out, axis=0)
