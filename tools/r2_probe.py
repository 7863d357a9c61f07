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
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
