"""hierarchical_marching_cubes(bunny, depth 9, n_subcell_depth 3) with and without the shared-face dedup of the lattice."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT]
import _niq  # noqa: E402
import implicit_mlp_utils  # noqa: E402
import kd_tree  # noqa: E402

with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("bunny/")}
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
ctx = _niq.default_context()
ref = None
for tag, env in (("dedup", None), ("one evaluation per leaf and point", "1"), ("dedup", None)):
    if env is None:
        os.environ.pop("NIQ_MC_NO_DEDUP", None)
    else:
        os.environ["NIQ_MC_NO_DEDUP"] = env
    kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 7, n_subcell_depth=3)
    ctx.mc_points(reset=True)
    t0 = time.perf_counter()
    tri = kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 9, n_subcell_depth=3)
    dt = time.perf_counter() - t0
    ev, lat = ctx.mc_points(reset=True)
    same = "" if ref is None else f", bit-identical to the first run: {np.array_equal(tri, ref)}"
    ref = tri if ref is None else ref
    print(f"{tag}: {dt * 1e3:.1f} ms, {tri.shape[0]} triangles, {lat // 729} leaves, evaluated {ev} of {lat} lattice points ({ev / lat:.3f}){same}", flush=True)
