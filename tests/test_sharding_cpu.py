"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in sharding.py: tile partition, the single
all_gather of ray results and the subtree deal + leaf gather.  The compute function is injected (the CPU
oracle stands in for the CUDA call), so what is tested is exactly the plumbing the GPU path uses."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, sample_params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tile_partition_covers_image_once():
    import sharding
    for (rx, ry, tile, world) in ((64, 48, 16, 2), (50, 30, 16, 3), (1920, 1080, 16, 8)):
        allp = np.concatenate([sharding.rank_pixels(rx, ry, tile, r, world) for r in range(world)])
        assert allp.shape[0] == rx * ry and np.unique(allp).shape[0] == rx * ry
        sizes = [len(sharding.rank_pixels(rx, ry, tile, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 2 * tile * tile * max(1, (rx // tile + 1) // world) or world > 2
    sub = sharding.rank_pixels(1920, 1080, 16, 0, 1, tile_stride=110)
    assert 74 * 256 <= sub.shape[0] <= 76 * 256 and np.unique(sub).shape[0] == sub.shape[0]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sharding
    from niq_oracle import net, rays, tree
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        params = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("fox/")}
    octx = net.AffineContext("affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    rx, ry = 24, 16
    roots, dirs = rays.generate_camera_rays(eye, look, up, res=rx, fov_deg=30., res_y=ry)
    opts = rays.get_default_cast_opts()
    cast = lambda f, p, r, d_, o: rays.cast_rays(f, p, r, d_, o)
    t, h, c, n_ev = sharding.cast_rays_sharded((octx,), (params,), roots, dirs, opts, rx, ry, tile=8, cast_fn=cast)

    build = lambda f, p, lo, hi, split_depth: tree.construct_uniform_unknown_levelset_tree(f, p, lo, hi, split_depth=split_depth)
    lo, hi = sharding.tree_sharded(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), 9, top_depth=4, build_fn=build)
    # closest point sharded by query (window >= stack regime) and a batch of intersection queries dealt round-robin
    qp = np.random.default_rng(3).uniform(-1, 1, (5, 3)).astype(np.float32)
    cp = lambda f, p, lo_, hi_, pts, eps, batch_process_size: tree.closest_point(f, p, lo_, hi_, pts, eps=eps, batch_process_size=batch_process_size)
    cd, cl = sharding.closest_point_sharded(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), qp, eps=0.4,
                                            batch_process_size=4096, cp_fn=cp)
    shifts = [(0.0, 0.0, 0.0), (1.9, 0.0, 0.0), (0.3, 0.2, 0.0)]
    params_of = lambda i: (params, net.prepend_op(params, net.spatial_transformation(np.eye(3, dtype=np.float32), np.array(shifts[i], np.float32))))
    isect = lambda fs, ps, lo_, hi_, eps: tree.find_any_intersection(fs, ps, lo_, hi_, eps)
    fi, fl = sharding.find_any_intersection_batch_sharded((octx, octx), params_of, 3, np.full(3, -1, np.float32), np.full(3, 1, np.float32), 0.1, isect_fn=isect)
    # frustum casting: initial tiles dealt round-robin, images summed, N_evals replayed from the summed iteration counts
    fopts = rays.get_default_cast_opts()
    fopts["n_side_init"] = 3
    _, _, left = rays.look_at(eye)
    cam = (eye, look, up, left, 30., 30., 12, 12)
    fcast = lambda f, p, cam_, o, init, it: rays.cast_rays_frustum(f, p, cam_, o, init_ranges=init, iter_counts=it)
    fr = sharding.cast_rays_frustum_sharded((octx,), (params,), cam, fopts, cast_fn=fcast)
    # hierarchical marching cubes: subtrees sharded, every rank extracts its own leaves, triangles gathered
    from niq_oracle import mc
    mcf = lambda f, p, lo_, hi_, n: mc.extract_mesh_from_leaves(p, lo_, hi_, n)
    tris = sharding.hierarchical_marching_cubes_sharded(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), 4,
                                                        n_subcell_depth=2, top_depth=3, build_fn=build, mc_fn=mcf)
    if rank == 0:
        q.put((t, h, c, n_ev, lo, hi, cd, cl, fi, fl, fr, tris))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_rays_and_tree_world2_gloo():
    from niq_oracle import net, rays, tree
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, h, c, n_ev, lo, hi, cd, cl, fi, fl, fr, tris = q.get(timeout=500)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    params = sample_params("fox")
    octx = net.AffineContext("affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    roots, dirs = rays.generate_camera_rays(eye, look, up, res=24, fov_deg=30., res_y=16)
    rt, rh, rc, rn_ev = rays.cast_rays((octx,), (params,), roots, dirs, rays.get_default_cast_opts())
    np.testing.assert_array_equal(t, rt)          # same oracle arithmetic per ray -> bit-identical after the gather
    np.testing.assert_array_equal(h, rh)
    np.testing.assert_array_equal(c, rc)
    assert n_ev == rn_ev                          # the WHOLE image's padded-lane count, replayed from the gathered step counts
    ref = tree.construct_uniform_unknown_levelset_tree(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), split_depth=9)
    v = ref["unknown_node_valid"]
    canon = lambda a, b: np.unique(np.concatenate((a, b), axis=1), axis=0)
    np.testing.assert_array_equal(canon(lo, hi), canon(ref["unknown_node_lower"][v], ref["unknown_node_upper"][v]))
    assert lo.shape[0] == int(v.sum())
    # closest point: per-query results of the sharded call equal the single-process call in the same regime
    qp = np.random.default_rng(3).uniform(-1, 1, (5, 3)).astype(np.float32)
    rd, rl = tree.closest_point(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), qp, eps=0.4, batch_process_size=4096)
    np.testing.assert_array_equal(cd, rd)
    np.testing.assert_array_equal(cl, rl)
    # intersection batch: identity transform intersects, a shift of 1.9 does not
    assert fi.shape == (3,) and bool(fi[0]) and not bool(fi[1]) and fl.shape == (3, 3)
    assert np.all(fl[1] == -777.)
    # frustum casting: the sharded images and N_evals equal the single-process call bit for bit
    fopts = rays.get_default_cast_opts()
    fopts["n_side_init"] = 3
    _, _, left = rays.look_at(eye)
    rt, rh, rc, rn = rays.cast_rays_frustum((octx,), (params,), (eye, look, up, left, 30., 30., 12, 12), fopts)
    np.testing.assert_array_equal(fr[0], rt)
    np.testing.assert_array_equal(fr[1], rh)
    np.testing.assert_array_equal(fr[2], rc)
    assert fr[3] == rn and (rh != 0).any()
    # marching cubes: the gathered soup equals the single-process soup as a set of triangles
    ref_tris = tree.hierarchical_marching_cubes(octx, params, np.full(3, -1, np.float32), np.full(3, 1, np.float32), 4, n_subcell_depth=2)
    canon_t = lambda a: np.unique(np.asarray(a, np.float32).reshape(-1, 9), axis=0)
    assert tris.shape == ref_tris.shape and tris.shape[0] > 100
    np.testing.assert_array_equal(canon_t(tris), canon_t(ref_tris))


def test_deal_boxes_is_a_partition_and_mixes_the_low_bits():
    """sharding.deal_boxes: every box has exactly one owner, every aligned group of `world` boxes gives one box to every rank,
    and -- the point of the digit-sum rule -- a rank's boxes are spread over all residues i mod world (for a kd-tree frontier
    i mod 8 is the octant of the domain: plain round-robin would hand each rank one octant)."""
    import sharding
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 9, 4096, 5000):
            owned = [sharding.deal_boxes(n, r, world) for r in range(world)]
            allb = np.sort(np.concatenate(owned)) if n else np.zeros(0, np.int64)
            assert np.array_equal(allb, np.arange(n))
            for r in range(world):
                groups = owned[r] // world
                assert np.array_equal(groups, np.arange(len(groups)))           # one per group, in order
                assert abs(len(owned[r]) - n / world) < 1
    mine = sharding.deal_boxes(4096, 3, 8)
    hist = np.bincount(mine % 8, minlength=8)
    assert hist.min() == hist.max() == 64                                       # all eight "octants" equally


def test_replay_n_evals_over_the_cast_rays_only():
    """sharding.replay_n_evals(count, opts, n_zero): listing only the rays that were cast and COUNTING the pixels nobody cast
    gives the reference's N_evals of the whole image (src/queries.py:137,164-173); the sharded ray call of a sub-sampled
    image relies on it."""
    import queries
    import sharding
    rng = np.random.default_rng(5)
    for n, n_cast, n_sub in ((100000, 7000, 1), (5000, 5000, 1), (40000, 123, 3), (129, 1, 1)):
        opts = dict(queries.get_default_cast_opts(), n_substeps=n_sub)
        count = np.zeros(n, np.int32)
        idx = rng.choice(n, n_cast, replace=False)
        count[idx] = rng.integers(1, int(opts["n_max_step"]) + 1, n_cast)
        full = sharding.replay_n_evals(count, opts)
        assert full == sharding.replay_n_evals(count[idx], opts, n_zero=n - n_cast)
