#!/usr/bin/env python
"""Round-2 multi-GPU check (torchrun, NCCL): every sharded entry point against the single-device call of the same rank,
results gathered from device buffers.  Prints one JSON line from rank 0.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/r2_nccl_check.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import implicit_mlp_utils
    import kd_tree
    import mlp
    import queries
    import render
    import sharding
    lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        mlps = {nm: {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")} for nm in ("fox", "bunny", "hammer", "birdcage_occ")}
    out = {"world": world}
    canon = lambda a, b: np.unique(np.concatenate((a, b), axis=1), axis=0)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize(); dist.barrier()
        return (time.perf_counter() - t0) / reps * 1e3, r

    # rays: fox 256 x 256, sharded by tile, gathered from device buffers
    p = mlps["fox"]
    f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=256, fov_deg=30.)
    o = queries.get_default_cast_opts()
    ms, r = timed(lambda: sharding.cast_rays_sharded((f,), (p,), roots, dirs, o, 256, 256))
    ref = queries.cast_rays((f,), (p,), roots, dirs, o)
    out["cast_rays_sharded"] = {"ms": ms, "equal": bool(all(np.array_equal(a, b) for a, b in zip(r[:3], ref[:3])) and r[3] == ref[3])}
    # tree: bunny depth 18, subtrees sharded
    pb = mlps["bunny"]
    fb = implicit_mlp_utils.generate_implicit_from_params(pb, "affine_fixed")
    ms, (tl, th) = timed(lambda: sharding.tree_sharded(fb, pb, lo, hi, 18))
    full = kd_tree.construct_uniform_unknown_levelset_tree(fb, pb, lo, hi, split_depth=18)
    v = full["unknown_node_valid"]
    out["tree_sharded_d18"] = {"ms": ms, "leaves": int(tl.shape[0]), "equal_as_set": bool(np.array_equal(canon(tl, th), canon(full["unknown_node_lower"][v], full["unknown_node_upper"][v])))}
    # marching cubes: depth 8 (n_subcell_depth 3), every rank extracts its own leaves
    ms, tris = timed(lambda: sharding.hierarchical_marching_cubes_sharded(fb, pb, lo, hi, 8, n_subcell_depth=3), reps=2)
    ref_t = kd_tree.hierarchical_marching_cubes(fb, pb, lo, hi, 8, n_subcell_depth=3)
    ct = lambda a: np.unique(np.asarray(a, np.float32).reshape(-1, 9), axis=0)
    out["hmc_sharded_d8"] = {"ms": ms, "triangles": int(tris.shape[0]), "equal_as_set": bool(tris.shape == ref_t.shape and np.array_equal(ct(tris), ct(ref_t)))}
    t0 = time.perf_counter(); kd_tree.hierarchical_marching_cubes(fb, pb, lo, hi, 8, n_subcell_depth=3); out["hmc_sharded_d8"]["single_gpu_ms"] = (time.perf_counter() - t0) * 1e3
    # closest point sharded by query (window >= stack regime)
    pc = mlps["birdcage_occ"]
    fc = implicit_mlp_utils.generate_implicit_from_params(pc, "affine_fixed")
    q = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)[:512]
    ms, (cd, cl) = timed(lambda: sharding.closest_point_sharded(fc, pc, lo, hi, q, eps=1e-3), reps=3)
    rd, rl = kd_tree.closest_point(fc, pc, lo, hi, q, eps=1e-3, batch_process_size=2 ** 26)
    t0 = time.perf_counter(); kd_tree.closest_point(fc, pc, lo, hi, q, eps=1e-3, batch_process_size=2 ** 26); single = (time.perf_counter() - t0) * 1e3
    out["closest_point_sharded_512"] = {"ms": ms, "single_gpu_ms": single, "equal": bool(np.array_equal(cd, rd) and np.array_equal(cl, rl))}
    # intersection batch dealt round-robin
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_isect_trunc64_list.npz"))
    pA = mlps["hammer"]
    pB = mlp.prepend_op(mlps["bunny"], mlp.spatial_transformation())
    kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
    fA = implicit_mlp_utils.generate_implicit_from_params(pA, "affine_truncate", **kw)
    fB = implicit_mlp_utils.generate_implicit_from_params(pB, "affine_truncate", **kw)
    ms, (bf, bl) = timed(lambda: sharding.find_any_intersection_transforms_sharded((fA, fB), (pA, pB), lo, hi, 1e-3, R_B=g["R"], t_B=g["t"]), reps=2)
    out["isect_batch_sharded_64"] = {"ms": ms, "found": int(bf.sum()), "equal_fixture": bool(np.array_equal(bf, g["found"])), "queries_per_s": 64 / ms * 1e3}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
