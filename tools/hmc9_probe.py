import os, sys, time, numpy as np
ROOT = "/root/repo"
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200")]
import implicit_mlp_utils as imu, kd_tree
with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("bunny/")}
f = imu.generate_implicit_from_params(p, "affine_fixed")
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 7, n_subcell_depth=3)
t0 = time.perf_counter(); tri = kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 9, n_subcell_depth=3); print("hmc9", time.perf_counter() - t0, tri.shape)
