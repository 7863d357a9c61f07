#!/usr/bin/env python
"""Generate the config-level parity fixtures (tests/golden/cfg*.npz) with the CPU ORACLE (oracle/niq_oracle), at the sizes
SURVEY.md 8(d) names for BASELINE configs 3, 4 and 5 -- minutes of CPU time each, so the GPU box replays the committed
outputs instead of paying for them.  The oracle itself is pinned to the unmodified reference by tests/golden/* (see
oracle/tools/gen_golden.py); these files pin the CUDA path to the oracle at config scale.

    python oracle/tools/gen_config_fixtures.py [cfg3 cfg4a cfg4b cfg5tree cfg5rays]   # default: all, one process each

Reads tests/golden/mlps.npz (the four sample MLPs as a derived fixture), never /root/reference.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

LO = np.full(3, -1, np.float32)
HI = np.full(3, 1, np.float32)


def sample_params(name):
    with np.load(os.path.join(GOLD, "mlps.npz")) as d:
        return {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}


def cfg3_transforms(n, seed=0):
    """rotation about z by U[0,2pi), translation U[-1.5,1.5]^3, np.random.default_rng(0)."""
    rng = np.random.default_rng(seed)
    Rs, ts = [], []
    for _ in range(n):
        th = rng.uniform(0, 2 * np.pi)
        Rs.append(np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32))
        ts.append(rng.uniform(-1.5, 1.5, 3).astype(np.float32))
    return np.stack(Rs), np.stack(ts)


def case_cfg3(n=64):
    """hammer x bunny, affine_truncate n_keep 64 'absolute', eps 1e-3 (src/main_intersection.py:86,95-99; src/kd_tree.py:567-655)."""
    from niq_oracle import net, tree
    pA = sample_params("hammer")
    R, t = cfg3_transforms(n)
    ctx = net.AffineContext("affine_truncate", truncate_count=64)
    out = dict(R=R, t=t, eps=np.float32(1e-3), n_trunc=np.int32(64), found=np.zeros(n, bool), loc=np.zeros((n, 3), np.float32),
               n_nodes=np.zeros(n, np.int64), n_rounds=np.zeros(n, np.int64), n_near_tie=np.zeros(n, np.int64))
    for i in range(n):
        pB = net.prepend_op(sample_params("bunny"), net.spatial_transformation(R[i], t[i]))
        st = {}
        found, _, _, loc = tree.find_any_intersection((ctx, ctx), (pA, pB), LO, HI, 1e-3, stats=st)
        out["found"][i], out["loc"][i] = bool(found), loc
        out["n_nodes"][i], out["n_rounds"][i], out["n_near_tie"][i] = st["n_nodes"], st["n_rounds"], st["n_near_tie"]
    return out


def _cfg4(B, Q=256):
    """birdcage_occ closest_point, affine_fixed, eps 1e-3, the first Q of rng(0).uniform(-1,1,(1e6,3)) (src/kd_tree.py:765-802)."""
    from niq_oracle import net, tree
    p = sample_params("birdcage_occ")
    q = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)[:Q]
    st = {}
    d, loc = tree.closest_point(net.AffineContext("affine_fixed"), p, LO, HI, q, eps=1e-3, batch_process_size=B, stats=st)
    return dict(query_points=q, eps=np.float32(1e-3), B=np.int64(B), dist=d, loc=loc, n_rounds=np.int64(st["n_rounds"]),
                n_visits=np.int64(st["n_visits"]), max_stack=np.int64(st["max_stack"]), n_near_tie=np.int64(st["n_near_tie"]),
                tie_query=st["tie_query"])


def case_cfg4a():
    return _cfg4(2048)


def case_cfg4b():
    return _cfg4(2 ** 21)


LAYERS5 = [3] + [256] * 8 + [1]


def case_cfg5tree():
    """config 5: depth-14 level-set tree of the random-init 3->256x8->1 ReLU MLP (NumPy seed 0), affine_fixed.  The leaf arrays
    are stored whole (order is part of the contract), plus the per-level sizes."""
    from niq_oracle import net, tree
    p = net.random_mlp(LAYERS5, "relu", seed=0)
    st = {}
    out = tree.construct_uniform_unknown_levelset_tree(net.AffineContext("affine_fixed"), p, LO, HI, split_depth=14, stats=st)
    v = out["unknown_node_valid"]
    return dict(n_valid=np.int64(v.sum()), padded_size=np.int64(v.shape[0]), lower=out["unknown_node_lower"][v],
                upper=out["unknown_node_upper"][v], n_evals=np.int64(st["n_evals"]), n_near_tie=np.int64(st["n_near_tie"]),
                level_sizes=np.asarray(st["level_sizes"], np.int64))


def case_cfg5rays(n_rays=256):
    """config 5: cast_rays of the same net at the FULL 512 steps on 256 rays of the 1920x1080 camera (a 16x16 tile at the
    image centre), default opts."""
    from niq_oracle import net, rays
    p = net.random_mlp(LAYERS5, "relu", seed=0)
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = rays.look_at(eye)
    roots, dirs = rays.generate_camera_rays(eye, look, up, res=1920, fov_deg=30., res_y=1080)
    yy, xx = np.meshgrid(np.arange(532, 548), np.arange(952, 968), indexing="ij")
    idx = (yy * 1920 + xx).reshape(-1)[:n_rays]
    opts = rays.get_default_cast_opts()
    t, hit, cnt, n_evals, tie = rays.cast_rays((net.AffineContext("affine_fixed"),), (p,), roots[idx], dirs[idx], opts,
                                                return_near_tie=True)
    return dict(roots=roots[idx], dirs=dirs[idx], t=t, hit=hit, count=cnt, n_evals=np.int64(n_evals), near_tie=tie)


CASES = {"cfg3": ("cfg3_isect_trunc64_list", case_cfg3), "cfg4a": ("cfg4_closest_birdcage_B2048", case_cfg4a),
         "cfg4b": ("cfg4_closest_birdcage_B2p21", case_cfg4b), "cfg5tree": ("cfg5_tree_d14", case_cfg5tree),
         "cfg5rays": ("cfg5_rays_256x512", case_cfg5rays)}


def _run(key):
    from threadpoolctl import threadpool_limits
    name, fn = CASES[key]
    t0 = time.time()
    with threadpool_limits(limits=2):
        out = fn()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: {time.time() - t0:.1f} s", flush=True)
    return name


if __name__ == "__main__":
    keys = sys.argv[1:] or list(CASES)
    with mp.get_context("fork").Pool(min(len(keys), 4)) as pool:
        pool.map(_run, keys, chunksize=1)
