#!/usr/bin/env python
"""Round-2 development probe (GPU box): times named workloads, or runs ONE of them once for an ncu capture.
usage: python tools/r2_probe.py time            -> timings of every workload (persistent vs level-loop tree, ...)
       python tools/r2_probe.py run <workload>  -> one invocation (wrap in ncu -k regex:<kernel>)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT):
    sys.path.insert(0, p)

import implicit_mlp_utils  # noqa: E402
import kd_tree  # noqa: E402
import mlp  # noqa: E402

LO, HI = np.full(3, -1, np.float32), np.full(3, 1, np.float32)


def sample(name):
    with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
        return {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}


def net5():
    return mlp.initialize_params(mlp.build_spec(mlp.quick_mlp_spec([3] + [256] * 8 + [1], "relu")), 0)


def tree(params, depth, mode="affine_fixed", **kw):
    f = implicit_mlp_utils.generate_implicit_from_params(params, mode)
    t = kd_tree.build_tree(f, params, LO, HI, split_depth=depth, **kw)
    st = t.stats()
    n = t.count(0)
    t.close()
    return st, n


def timed(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


def isect_setup():
    pA = sample("hammer")
    pB = mlp.prepend_op(sample("bunny"), mlp.spatial_transformation())
    kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
    fA = implicit_mlp_utils.generate_implicit_from_params(pA, "affine_truncate", **kw)
    fB = implicit_mlp_utils.generate_implicit_from_params(pB, "affine_truncate", **kw)
    pB["0000.spatial_transformation.R"] = np.eye(3, dtype=np.float32)
    pB["0000.spatial_transformation.t"] = np.array((1.45, 0., 0.), np.float32)
    return fA, fB, pA, pB


WORK = {
    "tree_bunny_d21": lambda: tree(sample("bunny"), 21),
    "tree_bunny_d12": lambda: tree(sample("bunny"), 12),
    "tree_net5_d14": lambda: tree(net5(), 14),
    "hmc_bunny_d9": lambda: kd_tree.hierarchical_marching_cubes(implicit_mlp_utils.generate_implicit_from_params(sample("bunny"), "affine_fixed"),
                                                                sample("bunny"), LO, HI, 9, n_subcell_depth=3).shape,
    "isect_trunc64_disjoint": lambda: (lambda s: kd_tree.find_any_intersection((s[0], s[1]), (s[2], s[3]), LO, HI, 1e-3))(isect_setup()),
}


def main():
    if sys.argv[1] == "run":
        print(WORK[sys.argv[2]]())
        return
    out = {}
    bunny, n5 = sample("bunny"), net5()
    for tag, p, d in (("bunny_d12", bunny, 12), ("bunny_d21", bunny, 21), ("net5_d14", n5, 14), ("fox_d18", sample("fox"), 18)):
        for legacy in ("0", "1"):
            os.environ["NIQ_TREE_LEGACY"] = legacy
            ms, (st, n) = timed(lambda: tree(p, d))
            out[f"tree_{tag}_{'level_loop' if legacy == '1' else 'persistent'}"] = {"ms": ms, "boxes": st["n_evals"], "leaves": n,
                                                                                 "boxes_per_s": st["n_evals"] / ms * 1e3}
    os.environ["NIQ_TREE_LEGACY"] = "0"
    s = isect_setup()
    st = {}
    ms, r = timed(lambda: kd_tree.find_any_intersection((s[0], s[1]), (s[2], s[3]), LO, HI, 1e-3, stats=st))
    out["isect_trunc64_disjoint"] = {"ms": ms, "found": bool(r[0]), **st}
    # the same query through the per-round host loop, and the 64 config-3 transforms: one by one and as ONE batch
    os.environ["NIQ_ISECT_LEGACY"] = "1"
    st = {}
    ms, r = timed(lambda: kd_tree.find_any_intersection((s[0], s[1]), (s[2], s[3]), LO, HI, 1e-3, stats=st))
    out["isect_trunc64_disjoint_round_loop"] = {"ms": ms, "found": bool(r[0]), **st}
    os.environ["NIQ_ISECT_LEGACY"] = "0"
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_isect_trunc64_list.npz"))
    def singles():
        n = 0
        for i in range(64):
            s[3]["0000.spatial_transformation.R"], s[3]["0000.spatial_transformation.t"] = g["R"][i], g["t"][i]
            n += bool(kd_tree.find_any_intersection((s[0], s[1]), (s[2], s[3]), LO, HI, 1e-3)[0])
        return n
    ms, nf = timed(singles, reps=2)
    out["cfg3_64_single_queries"] = {"ms": ms, "found": nf, "queries_per_s": 64 / ms * 1e3}
    for reps in (1, 4, 16):
        R, t = np.tile(g["R"], (reps, 1, 1)), np.tile(g["t"], (reps, 1))
        st = {}
        ms, r = timed(lambda: kd_tree.find_any_intersection_batch((s[0], s[1]), (s[2], s[3]), LO, HI, 1e-3, R_B=R, t_B=t, stats=st), reps=2)
        out[f"cfg3_batch_{64 * reps}"] = {"ms": ms, "found": int(r[0].sum()), "queries_per_s": 64 * reps / ms * 1e3, "nodes": int(st["n_nodes"].sum()),
                                          "nodes_per_s": int(st["n_nodes"].sum()) / ms * 1e3}
    # cast_rays in a growing-form mode: persistent kernel vs host-level loop (fox, 128 x 128)
    import queries
    import render
    pf = sample("fox")
    eye = np.array((2., 1., 2.), np.float32)
    look, up, _ = render.look_at(eye)
    roots, dirs = render.generate_camera_rays(eye, look, up, res=128, fov_deg=30.)
    for mode, kw in (("affine_truncate", dict(affine_n_truncate=8, affine_truncate_policy="absolute")), ("affine_all", {})):
        f = implicit_mlp_utils.generate_implicit_from_params(pf, mode, **kw)
        for host in ("", "1"):
            if host:
                os.environ["NIQ_RAYS_HOST_LOOP"] = "1"
            else:
                os.environ.pop("NIQ_RAYS_HOST_LOOP", None)
            ms, r = timed(lambda: queries.cast_rays((f,), (pf,), roots, dirs, queries.get_default_cast_opts()), reps=2)
            out[f"cast_rays_fox128_{mode}_{'host_loop' if host else 'persistent'}"] = {"ms": ms, "ray_steps": int(r[2].sum()), "ray_steps_per_s": int(r[2].sum()) / ms * 1e3}
    os.environ.pop("NIQ_RAYS_HOST_LOOP", None)
    # marching cubes (bank-skewed point tiles)
    fb = implicit_mlp_utils.generate_implicit_from_params(bunny, "affine_fixed")
    ms, tri = timed(lambda: kd_tree.hierarchical_marching_cubes(fb, bunny, LO, HI, 9, n_subcell_depth=3), reps=2)
    out["hmc_bunny_d9"] = {"ms": ms, "triangles": int(tri.shape[0])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
