"""cast_rays of the fox sample (BASELINE configs[0]: 512x512, affine_fixed, default opts) a few times: workload of the ncu
capture of k_cast_rays<32> and a quick timing."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), ROOT]
import implicit_mlp_utils  # noqa: E402
import queries  # noqa: E402
import render  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("fox/")}
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
eye = np.array((2., 1., 2.), np.float32)
look, up, left = render.look_at(eye)
roots, dirs = render.generate_camera_rays(eye, look, up, res=res, fov_deg=30.)
for i in range(reps):
    t0 = time.perf_counter()
    t, hit, cnt, n_evals = queries.cast_rays((f,), (p,), roots, dirs, queries.get_default_cast_opts())
    dt = time.perf_counter() - t0
    print(f"cast_rays {res}x{res}: {dt * 1e3:.2f} ms, {res * res / dt / 1e6:.2f} Mrays/s, {int(cnt.sum()) / dt / 1e6:.1f} M ray-steps/s, "
          f"hits {int((hit > 0).sum())}", flush=True)
