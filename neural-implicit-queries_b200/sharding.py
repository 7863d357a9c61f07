"""Multi-GPU partition of the hot path (new: the reference is single-device).

One process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).  Rays are
independent, so the image is cut into square pixel tiles dealt round-robin to the ranks (balances hit /
miss regions); every rank casts its own rays with no data-path collective, and ONE all_gather at the end
returns (t, hit_id, count) = 12 B/ray to every rank.  kd-tree subtrees shard the same way: the top levels
are built replicated, the frontier boxes are dealt round-robin and each rank finishes its own subtrees.
"""
import numpy as np


def tile_ids(res_x, res_y, tile):
    """Row-major ids of the tile grid covering a res_x x res_y image."""
    return (res_x + tile - 1) // tile, (res_y + tile - 1) // tile


def rank_pixels(res_x, res_y, tile, rank, world, tile_stride=1):
    """Flat pixel indices (y-major, as render.generate_camera_rays orders rays) owned by `rank`:
    every `tile_stride`-th tile of the grid (a stated sub-sample when > 1), dealt round-robin."""
    ntx, nty = tile_ids(res_x, res_y, tile)
    chosen = np.arange(0, ntx * nty, tile_stride)
    mine = chosen[rank::world]
    ty, tx = np.divmod(mine, ntx)
    oy, ox = np.meshgrid(np.arange(tile), np.arange(tile), indexing="ij")
    py = (ty[:, None, None] * tile + oy[None]).reshape(len(mine), -1)
    px = (tx[:, None, None] * tile + ox[None]).reshape(len(mine), -1)
    ok = (py < res_y) & (px < res_x)
    return (py * res_x + px)[ok].astype(np.int64)


def cast_rays_sharded(funcs_tuple, params_tuple, roots, dirs, opts, res_x, res_y, tile=16, cast_fn=None, group=None,
                      pixels_of_rank=None):
    """cast_rays over the full image with the rays partitioned across the ranks of `group`; every rank
    returns the full (out_t, out_hit_id, out_count, N_evals) like the single-device call.  N_evals is the
    reference's count for the WHOLE image (src/queries.py:137,164-173: padded lanes per iteration), replayed from the
    gathered per-ray step counts -- not the sum of the per-shard counts, whose bucket padding differs.
    `cast_fn` defaults to queries.cast_rays (tests inject a CPU stand-in to exercise the plumbing).
    `pixels_of_rank(r, world)` (optional) -> the flat pixel indices rank r casts (default: all tiles of the image dealt
    round-robin); pixels nobody casts stay zero (bench.py casts a stated sub-sample of the tiles)."""
    import torch
    import torch.distributed as dist
    device_path = cast_fn is None         # the product path keeps the results on the device until after the gather
    if cast_fn is None:
        import queries
        cast_fn = queries.cast_rays
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = roots.shape[0]
    assert n == res_x * res_y, "rays must be the full image in generate_camera_rays order"
    if pixels_of_rank is None:
        pixels_of_rank = lambda r, w: rank_pixels(res_x, res_y, tile, r, w)
    mine = pixels_of_rank(rank, world)
    if world == 1:
        t, hit, cnt, _ = cast_fn(funcs_tuple, params_tuple, roots[mine], dirs[mine], opts)[:4]
        out_t = np.zeros(n, np.float32); out_h = np.zeros(n, np.int32); out_c = np.zeros(n, np.int32)
        out_t[mine], out_h[mine], out_c[mine] = t, hit, cnt
        return out_t, out_h, out_c, replay_n_evals(out_c, opts)
    backend = dist.get_backend(group)
    shards = [pixels_of_rank(r, world) for r in range(world)]
    sizes = [len(x) for x in shards]
    cap = max(sizes)
    if backend == "nccl" and device_path:
        # results stay on the device: the persistent kernel writes (t, hit_id, count) into the rows of ONE packed buffer,
        # which is all-gathered over NVLink as it is (12 B/ray, one NCCL call), then read back once
        import _niq
        dev = torch.device("cuda", torch.cuda.current_device())
        r_d = torch.from_numpy(np.ascontiguousarray(roots[mine])).to(dev)
        d_d = torch.from_numpy(np.ascontiguousarray(dirs[mine])).to(dev)
        pack = torch.zeros((3, cap), dtype=torch.int32, device=dev)
        torch.cuda.current_stream().synchronize()              # the library runs on its own stream
        queries.cast_rays_device(funcs_tuple, params_tuple, len(mine), r_d.data_ptr(), d_d.data_ptr(), pack[0].data_ptr(),
                                 pack[1].data_ptr(), pack[2].data_ptr(), opts, want_n_evals=False,
                                 ctx=_niq.default_context(torch.cuda.current_device()))
        out = torch.empty((world, 3, cap), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(out, pack, group=group)
        out = out.cpu().numpy()
    else:
        t, hit, cnt, _ = cast_fn(funcs_tuple, params_tuple, roots[mine], dirs[mine], opts)[:4]
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        pack = torch.zeros((3, cap), dtype=torch.int32)
        pack[0, :len(mine)] = torch.from_numpy(t.view(np.int32))
        pack[1, :len(mine)] = torch.from_numpy(hit)
        pack[2, :len(mine)] = torch.from_numpy(cnt)
        pack = pack.to(dev)
        out = torch.empty((world, 3, cap), dtype=torch.int32, device=dev)
        if backend == "nccl":
            dist.all_gather_into_tensor(out, pack, group=group)
        else:
            dist.all_gather(list(out.unbind(0)), pack, group=group)
        out = out.cpu().numpy()
    out_t = np.zeros(n, np.float32); out_h = np.zeros(n, np.int32); out_c = np.zeros(n, np.int32)
    for r in range(world):
        idx = shards[r]
        out_t[idx] = out[r, 0, :sizes[r]].view(np.float32)
        out_h[idx] = out[r, 1, :sizes[r]]
        out_c[idx] = out[r, 2, :sizes[r]]
    # the replay needs the step counts of the rays that were cast + how many pixels nobody cast (count 0): on a stated sub-sample
    # of the tiles that is 4 % of the image, and a histogram over all 2 M pixels would cost more than the gather
    cast_counts = np.concatenate([out[r, 2, :sizes[r]] for r in range(world)])
    return out_t, out_h, out_c, replay_n_evals(cast_counts, opts, n_zero=n - int(cast_counts.shape[0]))


def replay_n_evals(count, opts, n_zero=0):
    """N_evals of src/queries.py:137,164-173 from the per-ray step counts of the whole image: every iteration evaluates
    the current padded array (n_substeps lanes each), which shrinks to the next bucket size whenever the live rays fit
    (src/bucketing.py:7-14,35-36).  The same replay niq_cast_rays does for one device.  `n_zero`: further rays of the image
    with a step count of 0 that are not listed in `count`."""
    from bucketing import get_next_bucket_size
    n_sub = int(opts['n_substeps'])
    n_bins = int(opts['n_max_step']) // n_sub + 3
    it = np.minimum((np.asarray(count, np.int64) + n_sub - 1) // n_sub, n_bins - 1)
    hist = np.bincount(it, minlength=n_bins)
    hist[0] += int(n_zero)
    cur = valid = int(it.shape[0]) + int(n_zero)
    evals = 0
    for k in range(1, n_bins):
        if valid <= 0:
            break
        evals += cur * n_sub
        valid -= int(hist[k])
        if valid <= 0:
            break
        nb = get_next_bucket_size(valid)
        if nb < cur:
            cur = nb
    return evals


def cast_rays_frustum_sharded(funcs_tuple, params_tuple, cam_params, opts, cast_fn=None, group=None):
    """cast_rays_frustum with the INITIAL TILES (src/queries.py:495-501) dealt round-robin to the ranks of `group`: a frustum
    never leaves its initial tile, so the ranks share no state.  Every rank marches its own tiles (pixels outside them
    stay zero), one all_reduce(SUM) of the three (res_x, res_y) images = 12 B/pixel assembles the result on every rank, and
    one more of the per-iteration (terminated, split) counts lets each rank replay the reference's N_evals for the WHOLE
    image.  `cast_fn(funcs, params, cam, opts, init_ranges, iter_counts)` defaults to queries.cast_rays_frustum."""
    import torch
    import torch.distributed as dist
    import queries
    if cast_fn is None:
        cast_fn = lambda f, p, cam, o, init, it: queries.cast_rays_frustum(f, p, cam, o, init_ranges=init, iter_counts=it)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    res_x, res_y = int(cam_params[6]), int(cam_params[7])
    init = queries._initial_frusta(res_x, res_y, int(opts['n_side_init']))
    iters = []
    t, hit, cnt = cast_fn(funcs_tuple, params_tuple, cam_params, opts, init[rank::world], iters)[:3]
    n_bins = int(opts['n_max_step']) // int(opts['n_substeps']) + 3
    counts = np.zeros((2, n_bins), np.int64)
    for k, (a, b) in enumerate(iters):
        counts[0, k], counts[1, k] = a, b
    if world > 1:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        img = torch.from_numpy(np.stack((t.view(np.int32), hit, cnt)).astype(np.int32)).to(dev)   # disjoint pixels: x + 0 is exact on the bits
        dist.all_reduce(img, group=group)
        cts = torch.from_numpy(counts).to(dev)
        dist.all_reduce(cts, group=group)
        img = img.cpu().numpy()
        t, hit, cnt = img[0].view(np.float32), img[1], img[2]
        counts = cts.cpu().numpy()
    n_evals = queries._frustum_n_evals(init.shape[0], counts[0].tolist(), counts[1].tolist())
    return np.ascontiguousarray(t), np.ascontiguousarray(hit), np.ascontiguousarray(cnt), n_evals


def default_top_depth(split_depth, world):
    """Levels built replicated before the frontier is dealt.  A level of a few thousand boxes costs every GPU the latency of
    ONE pass through the net whether it holds all of them or an eighth (the top of the tree is latency-bound), so replicating
    down to depth 12 (<= 4,096 boxes) is free and leaves each rank hundreds of boxes: the round-robin deal then balances the
    subtrees statistically instead of box by box."""
    return min(split_depth, max(12, int(np.ceil(np.log2(max(8 * world, 2)))) + 3))


def deal_boxes(n_boxes, rank, world):
    """Indices of the frontier boxes (top-of-tree leaves) owned by `rank`: box i goes to (sum of the base-`world` digits of i)
    mod world -- one box of every aligned group of `world` to each rank, rotating with the higher digits.  (Plain i mod world
    is the worst deal for a kd-tree frontier: the low bits of a node's index are its FIRST splits, so i mod 8 is the octant of
    the domain; csrc/niq_tree.cuh deal_owner is the same rule on the device.)"""
    i = np.arange(n_boxes, dtype=np.int64)
    if world <= 1:
        return i
    s, b = i % world, i // world
    while np.any(b > 0):
        s += b % world
        b //= world
    return i[s % world == rank]


def tree_sharded(func, params, lower, upper, split_depth, top_depth=None, build_fn=None, group=None, to_host=True, **kw):
    """construct_uniform_unknown_levelset_tree with subtrees sharded across ranks: the top `top_depth`
    levels are built on every rank (replicated, tiny), the surviving frontier is dealt (deal_boxes), each rank
    refines its own boxes to `split_depth`, and the UNKNOWN leaves are all-gathered (24 B/leaf).
    Leaf ORDER differs from the single-device call (canonicalise before comparing).  -> (lower, upper) (L,3).
    to_host=False (NCCL only): return (gathered device tensor (world, 2, cap, 3), per-rank counts) without the read-back."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    backend = dist.get_backend(group) if world > 1 else None
    if world > 1 and backend == "nccl" and build_fn is None:
        return _tree_sharded_device(func, params, lower, upper, split_depth, top_depth, rank, world, group, kw, to_host)
    lo, hi = _own_leaves(func, params, lower, upper, split_depth, top_depth, build_fn, rank, world, kw)
    if world == 1:
        return lo, hi
    parts = _gather_rows(np.concatenate((lo, hi), axis=1), world, rank, group, 6, torch.float32)
    allp = np.concatenate(parts)
    return allp[:, :3].copy(), allp[:, 3:].copy()


def _top_frontier(func, params, lower, upper, split_depth, top_depth, world, kw):
    import kd_tree
    if top_depth is None:
        top_depth = default_top_depth(split_depth, world)
    top = kd_tree.construct_uniform_unknown_levelset_tree(func, params, lower, upper, split_depth=top_depth, **kw)
    v = top['unknown_node_valid']
    return top['unknown_node_lower'][v], top['unknown_node_upper'][v], top_depth


def _build_own_tree(func, params, lower, upper, split_depth, top_depth, rank, world, kw):
    """This rank's subtrees as a device-resident tree: one dealt persistent launch where the mode has one
    (niq_tree_build_dealt), else the top on every rank, the deal on the host and one multi-root build."""
    import _niq
    import kd_tree
    tkw = {k: v for k, v in kw.items() if k in ("offset", "batch_process_size", "ctx")}
    if top_depth is None:
        top_depth = default_top_depth(split_depth, world)
    try:
        return kd_tree.build_tree_dealt(func, params, lower, upper, split_depth, top_depth, rank, world, **tkw)
    except _niq.NiqError as e:
        if e.code != _niq.NIQ_EUNSUPPORTED:
            raise
    flo, fhi, top_depth = _top_frontier(func, params, lower, upper, split_depth, top_depth, world, kw)
    mine = deal_boxes(flo.shape[0], rank, world)
    if not len(mine):
        return None
    return kd_tree.build_tree(func, params, flo[mine], fhi[mine], split_depth=split_depth - top_depth, **tkw)


# (split_depth, top_depth, world) -> rows per rank of the last gather: the next one needs no count exchange.  Like every collective
# call this relies on all ranks making the SAME sequence of tree_sharded calls (SPMD): the entry then exists on all ranks or on none.
_GATHER_CAP = {}


def _tree_sharded_device(func, params, lower, upper, split_depth, top_depth, rank, world, group, kw, to_host=True):
    """NCCL path: this rank's leaves never visit the host before the gather -- the tree's device-resident leaf list is copied
    into a padded device buffer (niq_tree_copy, device to device) and ONE all_gather_into_tensor moves 24 B/leaf over NVLink;
    the gathered leaves are read back once.  The buffer's first row carries the rank's leaf count, so a call whose capacity is
    known from the previous one (same depths and world) needs no separate exchange of the counts; when a rank's leaves do not fit
    (every rank sees that in the gathered counts) the gather is repeated with the capacity they need."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import _niq
    dev = torch.device("cuda", torch.cuda.current_device())
    tree = _build_own_tree(func, params, lower, upper, split_depth, top_depth, rank, world, kw)
    n_mine = tree.count(0) if tree is not None else 0
    key = (int(split_depth), top_depth, world)
    try:
        cap = _GATHER_CAP.get(key)
        if cap is None:
            cnt = torch.tensor([n_mine], dtype=torch.int64, device=dev)
            counts = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(counts, cnt, group=group)
            cap = int(counts.max().item()) * 9 // 8 + 64
        while True:
            # rows 1..cap of each half hold the leaves, row 0 of the first half the count (as int32 bits)
            pack = torch.zeros((2, cap + 1, 3), dtype=torch.float32, device=dev)
            pack.view(torch.int32)[0, 0, 0] = n_mine
            torch.cuda.current_stream().synchronize()
            if 0 < n_mine <= cap:
                _niq.check(_niq.lib().niq_tree_copy(tree.handle, C.c_int(0), C.c_void_p(pack[0, 1:].data_ptr()),
                                                    C.c_void_p(pack[1, 1:].data_ptr()), C.c_int64(cap), C.c_int(_niq.MEM_DEVICE)))
            out = torch.empty((world, 2, cap + 1, 3), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(out, pack, group=group)
            counts = out.view(torch.int32)[:, 0, 0, 0].cpu().tolist()
            if max(counts) <= cap:
                break
            cap = max(counts) * 9 // 8 + 64
        _GATHER_CAP[key] = max(counts) * 9 // 8 + 64        # follows the tree actually built (another net under the same key)
        if not to_host:
            torch.cuda.current_stream().synchronize()
            return out[:, :, 1:], counts
        out = out.cpu().numpy()[:, :, 1:]             # one contiguous read-back; the count row is dropped on the host
    finally:
        if tree is not None:
            tree.close()
    lo = np.concatenate([out[r, 0, :counts[r]] for r in range(world)])
    hi = np.concatenate([out[r, 1, :counts[r]] for r in range(world)])
    return lo, hi


def _own_leaves(func, params, lower, upper, split_depth, top_depth, build_fn, rank, world, kw):
    """The UNKNOWN leaves at `split_depth` below this rank's share of the frontier (top levels replicated)."""
    if build_fn is None:
        # the CUDA backend: one dealt persistent launch (or top + host deal + one multi-root build, _build_own_tree)
        tree = _build_own_tree(func, params, lower, upper, split_depth, top_depth, rank, world, kw)
        if tree is None:
            return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
        try:
            lo, hi = tree.nodes(0)
        finally:
            tree.close()
        return np.ascontiguousarray(lo, np.float32).reshape(-1, 3), np.ascontiguousarray(hi, np.float32).reshape(-1, 3)
    # an injected builder (the CPU tests run the oracle through this): top tree, round-robin deal, one build per frontier box
    if top_depth is None:
        top_depth = default_top_depth(split_depth, world)
    top = build_fn(func, params, lower, upper, split_depth=top_depth, **kw)
    v = top['unknown_node_valid']
    flo, fhi = top['unknown_node_lower'][v], top['unknown_node_upper'][v]
    los, his = [], []
    for i in deal_boxes(flo.shape[0], rank, world):
        sub = build_fn(func, params, flo[i], fhi[i], split_depth=split_depth - top_depth, **kw)
        sv = sub['unknown_node_valid']
        los.append(sub['unknown_node_lower'][sv]); his.append(sub['unknown_node_upper'][sv])
    lo = np.concatenate(los) if los else np.zeros((0, 3), np.float32)
    hi = np.concatenate(his) if his else np.zeros((0, 3), np.float32)
    return np.ascontiguousarray(lo, np.float32).reshape(-1, 3), np.ascontiguousarray(hi, np.float32).reshape(-1, 3)


def hierarchical_marching_cubes_sharded(func, params, lower, upper, depth, n_subcell_depth=2, top_depth=None, build_fn=None,
                                        mc_fn=None, group=None):
    """hierarchical_marching_cubes (src/kd_tree.py:357-399) with the subtrees sharded: the top of the level-set tree is built
    replicated, every rank refines its share of the frontier boxes to split depth 3*(depth - n_subcell_depth) and extracts
    the triangles of ITS OWN leaves; one variable-length all_gather of 36 B/triangle returns the whole soup to every rank.
    Triangle ORDER is rank-major (canonicalise before comparing with the single-device call).
    `mc_fn(func, params, leaf_lower, leaf_upper, n_subcell_depth)` defaults to extract_cell.extract_mesh_from_cells."""
    import torch.distributed as dist
    if mc_fn is None:
        import extract_cell
        mc_fn = extract_cell.extract_mesh_from_cells
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    split_depth = 3 * (depth - n_subcell_depth)
    if world > 1 and build_fn is None and mc_fn is extract_cell_default() and dist.get_backend(group) == "nccl":
        return _hmc_sharded_device(func, params, lower, upper, split_depth, n_subcell_depth, top_depth, rank, world, group)
    lo, hi = _own_leaves(func, params, lower, upper, split_depth, top_depth, build_fn, rank, world, {})
    tri = mc_fn(func, params, lo, hi, n_subcell_depth) if lo.shape[0] > 0 else np.zeros((0, 3, 3), np.float32)
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    if world == 1:
        return tri.reshape(-1, 3, 3)
    import torch
    parts = _gather_rows(tri, world, rank, group, 9, torch.float32)
    return np.concatenate(parts).reshape(-1, 3, 3)


def extract_cell_default():
    import extract_cell
    return extract_cell.extract_mesh_from_cells


def _hmc_sharded_device(func, params, lower, upper, split_depth, n_subcell_depth, top_depth, rank, world, group):
    """NCCL path of the sharded marching cubes: nothing visits the host before the gather.  The rank's own subtrees are one
    dealt launch (device-resident leaf list), the marching-cubes kernels read that list (niq_marching_cubes_tree), the
    triangles are copied device to device into a padded buffer (niq_mesh_copy, NIQ_MEM_DEVICE) and ONE all_gather_into_tensor
    moves 36 B/triangle over NVLink after an 8-byte exchange of the counts; one read-back."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import _niq
    ctx = _niq.default_context(torch.cuda.current_device())
    dev = torch.device("cuda", torch.cuda.current_device())
    tree = _build_own_tree(func, params, lower, upper, split_depth, top_depth, rank, world, {"ctx": ctx})
    mesh = C.c_void_p()
    n_tri = 0
    try:
        if tree is not None and tree.count(0) > 0:
            m = ctx.mlp(params)
            _niq.check(_niq.lib().niq_marching_cubes_tree(ctx.handle, m.handle, tree.handle, C.c_int32(n_subcell_depth), C.byref(mesh)))
            n = C.c_int64()
            _niq.check(_niq.lib().niq_mesh_count(mesh, C.byref(n)))
            n_tri = int(n.value)
        cnt = torch.tensor([n_tri], dtype=torch.int64, device=dev)
        counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, cnt, group=group)
        counts = counts.cpu().tolist()
        cap = max(max(counts), 1)
        pack = torch.zeros((cap, 9), dtype=torch.float32, device=dev)
        torch.cuda.current_stream().synchronize()              # the library runs on its own stream
        if n_tri:
            _niq.check(_niq.lib().niq_mesh_copy(mesh, C.c_void_p(pack.data_ptr()), C.c_int64(cap), C.c_int(_niq.MEM_DEVICE)))
        out = torch.empty((world, cap, 9), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, pack, group=group)
        out = out.cpu().numpy()
    finally:
        if mesh:
            _niq.lib().niq_mesh_destroy(mesh)
        if tree is not None:
            tree.close()
    return np.concatenate([out[r, :c] for r, c in enumerate(counts)]).reshape(-1, 3, 3)


def _gather_rows(local, world, rank, group, width, dtype):
    """all_gather of per-rank (n_r, width) blocks of different lengths -> list of arrays: an 8-byte all_gather of the
    counts, then ONE padded all_gather of the rows (NCCL: all_gather_into_tensor on device buffers)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pack = torch.zeros((cap, width), dtype=dtype)
    pack[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local))
    pack = pack.to(dev)
    out = torch.empty((world, cap, width), dtype=dtype, device=dev)
    if backend == "nccl":
        dist.all_gather_into_tensor(out, pack, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), pack, group=group)
    out = out.cpu().numpy()
    return [out[r, :c] for r, c in enumerate(counts)]


def closest_point_sharded(func, params, lower, upper, query_points, eps=0.001, batch_process_size=2 ** 26, cp_fn=None,
                          group=None):
    """closest_point with the QUERIES partitioned across ranks (contiguous ranges) and one gather of (dist, loc) =
    16 B/query.  Exact only in the `batch_process_size >= stack` regime, where the reference's algorithm is
    level-synchronous per query and the queries do not interact (with the default global window of 2048 the LIFO stack
    couples them: SURVEY.md F6) -- hence the large default window here.  Every rank returns the full result."""
    import torch
    import torch.distributed as dist
    if cp_fn is None:
        import kd_tree
        cp_fn = kd_tree.closest_point
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    q = np.ascontiguousarray(query_points, np.float32)
    bounds = np.linspace(0, q.shape[0], world + 1).astype(np.int64)
    mine = q[bounds[rank]:bounds[rank + 1]]
    if mine.shape[0]:
        d, loc = cp_fn(func, params, lower, upper, mine, eps=eps, batch_process_size=batch_process_size)
    else:
        d, loc = np.zeros(0, np.float32), np.zeros((0, 3), np.float32)
    if world == 1:
        return d, loc
    parts = _gather_rows(np.concatenate((d[:, None], loc), axis=1).astype(np.float32), world, rank, group, 4, torch.float32)
    allp = np.concatenate(parts)
    return allp[:, 0].copy(), allp[:, 1:].copy()


def find_any_intersection_batch_sharded(func_tuple, params_of, n_queries, lower, upper, eps, isect_fn=None, group=None):
    """A BATCH of pairwise intersection queries (e.g. one per rigid transform) dealt round-robin to the ranks; a single
    query does not shard (<= a few thousand nodes, global early exit: SURVEY.md 8(e) "replicas only").
    `params_of(i)` -> params_tuple of query i.  -> (found (n,) bool, loc (n,3)) on every rank."""
    import torch
    import torch.distributed as dist
    if isect_fn is None:
        import kd_tree
        isect_fn = kd_tree.find_any_intersection
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = np.arange(rank, n_queries, world)
    rows = np.zeros((len(mine), 5), np.float32)
    for j, i in enumerate(mine):
        found, _, _, loc = isect_fn(func_tuple, params_of(int(i)), lower, upper, eps)
        rows[j] = (i, float(bool(found)), *np.asarray(loc, np.float32))
    if world > 1:
        rows = np.concatenate(_gather_rows(rows, world, rank, group, 5, torch.float32))
    order = np.argsort(rows[:, 0], kind="stable")
    rows = rows[order]
    return rows[:, 1] > 0.5, rows[:, 2:].copy()


def find_any_intersection_transforms_sharded(func_tuple, params_tuple, lower, upper, eps, R_B=None, t_B=None, R_A=None, t_A=None,
                                             batch_fn=None, group=None):
    """kd_tree.find_any_intersection_batch with the queries (one per rigid transform) cut into contiguous ranges, one range per
    rank: every rank runs ITS queries in one persistent kernel, one gather of (found, loc) = 16 B/query returns everything to
    every rank.  -> (found (n,) bool, loc (n,3))."""
    import torch
    import torch.distributed as dist
    if batch_fn is None:
        import kd_tree
        batch_fn = kd_tree.find_any_intersection_batch
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = (R_B if R_B is not None else R_A).shape[0]
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    a, b = int(bounds[rank]), int(bounds[rank + 1])
    cut = lambda x: None if x is None else np.ascontiguousarray(x[a:b])
    if b > a:
        found, loc = batch_fn(func_tuple, params_tuple, lower, upper, eps, R_B=cut(R_B), t_B=cut(t_B), R_A=cut(R_A), t_A=cut(t_A))
    else:
        found, loc = np.zeros(0, bool), np.zeros((0, 3), np.float32)
    if world == 1:
        return found, loc
    rows = np.concatenate((found.astype(np.float32)[:, None], np.asarray(loc, np.float32)), axis=1)
    allr = np.concatenate(_gather_rows(rows, world, rank, group, 4, torch.float32))
    return allr[:, 0] > 0.5, allr[:, 1:].copy()
