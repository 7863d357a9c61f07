#!/usr/bin/env python
"""CPU-port timings of configs 1-4 on bounded samples (no GPU needed): the oracle (NumPy float32 restatement of the
reference, its batching structure kept) on this machine's cores.  Context for bench.py --extra; prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
from niq_oracle import net, rays, tree as otree  # noqa: E402

with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    mlps = {nm: {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")} for nm in ("fox", "bunny", "hammer", "birdcage_occ")}
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
out = {"cores": os.cpu_count(), "blas": "numpy default threads"}
fixed = net.AffineContext("affine_fixed")

# cfg 1: fox cast_rays, 128x128 of the 512x512 image (every 4th pixel of the same camera)
eye = np.array((2., 1., 2.), np.float32)
look, up, _ = rays.look_at(eye)
roots, dirs = rays.generate_camera_rays(eye, look, up, res=128, fov_deg=30.)
t0 = time.perf_counter()
t, hit, cnt, ne = rays.cast_rays((fixed,), (mlps["fox"],), roots, dirs, rays.get_default_cast_opts())
dt = time.perf_counter() - t0
out["cfg1_fox_cast_rays_128x128"] = {"rays_per_s": roots.shape[0] / dt, "ray_steps_per_s": int(cnt.sum()) / dt, "s": dt}
print(json.dumps(out), flush=True)

# cfg 2: bunny tree depth 12 and depth 17 (deeper levels prune)
for depth in (12, 17):
    st = {}
    t0 = time.perf_counter()
    tr = otree.construct_uniform_unknown_levelset_tree(fixed, mlps["bunny"], lo, hi, split_depth=depth, stats=st)
    dt = time.perf_counter() - t0
    out[f"cfg2_bunny_tree_depth{depth}"] = {"boxes_per_s": st.get("n_evals", 2 ** (depth + 1) - 1) / dt, "s": dt}
print(json.dumps(out), flush=True)
t0 = time.perf_counter()
tri = otree.hierarchical_marching_cubes(fixed, mlps["bunny"], lo, hi, 5, n_subcell_depth=3)
dt = time.perf_counter() - t0
out["cfg2_bunny_hmc_depth5_sub3"] = {"leaves_per_s": 64 / dt, "triangles": int(tri.shape[0]), "s": dt}
print(json.dumps(out), flush=True)

# cfg 3: hammer x bunny, affine_truncate 64, seeded transforms (the first 4 of bench.py's list after its warm-up draw)
tr64 = net.AffineContext("affine_truncate", truncate_count=64)
rng = np.random.default_rng(0)
n_q, n_nodes, t_tot = 4, 0, 0.0
for i in range(n_q + 1):
    th = rng.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    tt = rng.uniform(-1.5, 1.5, 3).astype(np.float32)
    if i == 0:
        continue
    pB = net.prepend_op(mlps["bunny"], net.spatial_transformation(R, tt))
    st = {}
    t0 = time.perf_counter()
    otree.find_any_intersection((tr64, tr64), (mlps["hammer"], pB), lo, hi, 1e-3, stats=st)
    t_tot += time.perf_counter() - t0
    n_nodes += st["n_nodes"]
out["cfg3_intersection_truncate64"] = {"queries_per_s": n_q / t_tot, "nodes_per_s": n_nodes / t_tot, "s": t_tot, "queries": n_q}
print(json.dumps(out), flush=True)

# cfg 4: birdcage closest_point, 4 queries, window 2048
q = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)[:4]
st = {}
t0 = time.perf_counter()
otree.closest_point(fixed, mlps["birdcage_occ"], lo, hi, q, eps=1e-3, batch_process_size=2048, stats=st)
dt = time.perf_counter() - t0
out["cfg4_closest_point_window2048"] = {"queries_per_s": 4 / dt, "node_visits_per_s": st["n_visits"] / dt, "s": dt}
print(json.dumps(out), flush=True)
