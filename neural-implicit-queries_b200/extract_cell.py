"""`extract_cell` surface of the reference (/root/reference/src/extract_cell.py): the marching-cubes case
tables (:12-302, get_mc_data :306-310) and per-cell extraction over a lattice of subcells (:366-421).
Tables and extraction live in the CUDA library (csrc/mc_tables.h, k_mc_count / k_mc_write)."""
import ctypes as C

import numpy as np

import _niq


def get_mc_data():
    """-> (tri_table (256,16) i32, edge_verts (12,2) i32, vert_logical_coords (8,3) bool)"""
    tri = np.empty((256, 16), np.int32)
    ev = np.empty((12, 2), np.int32)
    vc = np.empty((8, 3), np.uint8)
    _niq.check(_niq.lib().niq_mc_tables(_niq.ptr(tri), _niq.ptr(ev), _niq.ptr(vc)))
    return tri, ev, vc.astype(bool)


def extract_mesh_from_cells(func, params, cell_lower, cell_upper, n_sub_depth, ctx=None):
    """Triangles of all cells (lo/hi (L,3)), order = cell, subcell ('ij', axis 0 slowest), slot.  (T,3,3) f32."""
    ctx = ctx or _niq.default_context()
    lo = np.ascontiguousarray(cell_lower, np.float32).reshape(-1, 3)
    hi = np.ascontiguousarray(cell_upper, np.float32).reshape(-1, 3)
    mesh = C.c_void_p()
    m = ctx.mlp(params)
    _niq.check(_niq.lib().niq_marching_cubes(ctx.handle, m.handle, C.c_int64(lo.shape[0]), _niq.ptr(lo), _niq.ptr(hi),
                                             C.c_int32(n_sub_depth), C.c_int(_niq.MEM_HOST), C.byref(mesh)))
    try:
        return _mesh_to_numpy(mesh)
    finally:
        _niq.lib().niq_mesh_destroy(mesh)


def _mesh_to_numpy(mesh):
    n = C.c_int64()
    _niq.check(_niq.lib().niq_mesh_count(mesh, C.byref(n)))
    tri = np.empty((n.value, 3, 3), np.float32)
    if n.value:
        _niq.check(_niq.lib().niq_mesh_copy(mesh, _niq.ptr(tri), C.c_int64(n.value), C.c_int(_niq.MEM_HOST)))
    return tri


def extract_triangles_from_subcells(func, params, mc_data, n_sub_depth, cell_lower, cell_upper, batch_eval_size=4096):
    """:366-421 for ONE cell -> (tri_pos (5*(2^n)^3, 3, 3), tri_is_valid (5*(2^n)^3,)).
    The valid triangles come from the CUDA extraction; the padded (slot-indexed) layout of the reference is
    rebuilt on the host from the case table so that `tri_pos[tri_is_valid]` equals the reference's."""
    import mlp
    tri_table, edge_verts, vlc = mc_data
    lo = np.asarray(cell_lower, np.float32)
    hi = np.asarray(cell_upper, np.float32)
    n = 2 ** n_sub_depth
    P = n + 1
    tris = extract_mesh_from_cells(func, params, lo[None], hi[None], n_sub_depth)
    # validity pattern: evaluate the lattice once more (same kernel, same bits) and look the cases up
    ax = [np.linspace(lo[d], hi[d], P, dtype=np.float32) for d in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing='ij'), axis=-1).reshape(-1, 3)
    vals = mlp.eval_points(params, g).reshape(P, P, P)
    ii = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing='ij'), axis=-1).reshape(-1, 3)
    vi = ii[:, None, :] + np.asarray(vlc, np.int64)[None, :, :]
    vv = vals[vi[..., 0], vi[..., 1], vi[..., 2]]
    case = ((vv < 0) * (1 << np.arange(8))[None, :]).sum(axis=1)
    valid = (np.asarray(tri_table)[case][:, 0:15:3] >= 0).reshape(-1)
    out = np.zeros((valid.shape[0], 3, 3), np.float32)
    assert int(valid.sum()) == tris.shape[0]
    out[valid] = tris
    return out, valid
