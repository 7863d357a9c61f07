#!/usr/bin/env python
"""Development: run-to-run spread of the config-3 single-query latency (64 seeded transforms, 6 passes)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import _niq, implicit_mlp_utils, kd_tree, mlp  # noqa: E402,E401

ctx = _niq.default_context(0)
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
pA = bench.sample_mlp("hammer")
pB = mlp.prepend_op(bench.sample_mlp("bunny"), mlp.spatial_transformation())
kw = dict(affine_n_truncate=64, affine_truncate_policy="absolute")
fA = implicit_mlp_utils.generate_implicit_from_params(pA, "affine_truncate", **kw)
fB = implicit_mlp_utils.generate_implicit_from_params(pB, "affine_truncate", **kw)
R, t = bench.cfg3_transforms(64)
for rep in range(6):
    lat = []
    for i in range(64):
        pB["0000.spatial_transformation.R"], pB["0000.spatial_transformation.t"] = R[i], t[i]
        t0 = time.perf_counter()
        kd_tree.find_any_intersection((fA, fB), (pA, pB), lo, hi, 1e-3, ctx=ctx)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e3
    print(f"pass {rep}: mean {lat.mean():.2f} ms  median {np.median(lat):.2f}  max {lat.max():.2f} (query {int(lat.argmax())})  p90 {np.percentile(lat, 90):.2f}", flush=True)
