"""Frustum casting of the fox sample (1024x1024, default opts) a few times: the workload of the ncu capture of k_cast_frustum
(tools/gpu_frustum.sh) and a quick timing next to cast_rays of the same image."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT]
import implicit_mlp_utils  # noqa: E402
import queries  # noqa: E402
import render  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("fox/")}
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
eye = np.array((2., 1., 2.), np.float32)
look, up, left = render.look_at(eye)
cam = (eye, look, up, left, 30., 30., res, res)
for i in range(reps):
    t0 = time.perf_counter()
    it = []
    t, hit, cnt, n_evals = queries.cast_rays_frustum((f,), (p,), cam, queries.get_default_cast_opts(), iter_counts=it)
    dt = time.perf_counter() - t0
    print(f"frustum {res}x{res}: {dt * 1e3:.2f} ms, {res * res / dt / 1e6:.2f} Mpix/s, hits {int((hit > 0).sum())}, N_evals {n_evals}, "
          f"iterations {len(it)}, frusta finished {sum(a for a, _ in it)}, splits {sum(b for _, b in it)}", flush=True)
