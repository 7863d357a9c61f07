"""development: touch every kernel family once on tiny inputs (for compute-sanitizer memcheck)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "neural-implicit-queries_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import implicit_mlp_utils as imu, kd_tree, queries, render, mlp
from niq_oracle import net
with np.load(os.path.join(ROOT, "tests", "golden", "mlps.npz")) as d:
    P = {nm: {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")} for nm in ("fox", "bunny", "hammer")}
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
rng = np.random.default_rng(0)
c = rng.uniform(-1, 1, (70, 3)).astype(np.float32); h = (0.05 * rng.uniform(0.3, 1, (70, 3))).astype(np.float32)
for nm in ("fox", "bunny"):
    p = P[nm]
    for mode, kw in (("interval", {}), ("affine_fixed", {}), ("affine_truncate", dict(affine_n_truncate=8, affine_truncate_policy="absolute")),
                     ("affine_all", {}), ("affine_append", dict(affine_n_append=4)), ("sdf", dict(sdf_lipschitz=1.5)), ("slope_interval", {})):
        f = imu.generate_implicit_from_params(p, mode, **kw)
        f.bound_box(p, c - h, c + h)
        f.bound_general_box(p, c, h[:, None, :])
        print(nm, mode, "ok", flush=True)
p = P["fox"]; f = imu.generate_implicit_from_params(p, "affine_fixed")
eye = np.array((2., 1., 2.), np.float32); look, up, _ = render.look_at(eye)
roots, dirs = render.generate_camera_rays(eye, look, up, res=12, fov_deg=30.)
queries.cast_rays((f,), (p,), roots, dirs, queries.get_default_cast_opts()); print("rays ok", flush=True)
kd_tree.construct_uniform_unknown_levelset_tree(f, p, lo, hi, split_depth=9, with_interior_nodes=True, with_exterior_nodes=True); print("tree ok", flush=True)
kd_tree.hierarchical_marching_cubes(f, p, lo, hi, 4, n_subcell_depth=2); print("mc ok", flush=True)
pB = mlp.prepend_op(P["bunny"], mlp.spatial_transformation()); pB["0000.spatial_transformation.t"] = np.array((1.2, 0.1, 0.05), np.float32)
fa = imu.generate_implicit_from_params(P["hammer"], "affine_fixed"); fb = imu.generate_implicit_from_params(pB, "affine_fixed")
kd_tree.find_any_intersection((fa, fb), (P["hammer"], pB), lo, hi, 1e-2); print("isect ok", flush=True)
q = rng.uniform(-1, 1, (6, 3)).astype(np.float32)
kd_tree.closest_point(f, p, lo, hi, q, eps=0.05, batch_process_size=64); print("closest windowed ok", flush=True)
kd_tree.closest_point(f, p, lo, hi, q, eps=0.05, batch_process_size=2 ** 14); print("closest level ok", flush=True)
syn = net.random_mlp([3] + [128] * 3 + [1], "relu", seed=0)
fs = imu.generate_implicit_from_params(syn, "affine_fixed")
fs.bound_box(syn, c - 0.01 * h, c + 0.01 * h)
imu.generate_implicit_from_params(syn, "slope_interval").bound_box(syn, c - 0.01 * h, c + 0.01 * h)
o = queries.get_default_cast_opts(); o["n_max_step"] = 6
queries.cast_rays((fs,), (syn,), roots, dirs, o); print("streamed sparse ok", flush=True)
cam = (eye, look, up, render.look_at(eye)[2], 30., 30., 12, 12)
fo = queries.get_default_cast_opts(); fo["n_side_init"] = 3
queries.cast_rays_frustum((f,), (p,), cam, fo); print("frustum resident ok", flush=True)
fo2 = dict(fo); fo2["n_max_step"] = 8; fo2["n_substeps"] = 2
queries.cast_rays_frustum((fs,), (syn,), cam, fo2); print("frustum streamed ok", flush=True)
queries.cast_rays_frustum((imu.generate_implicit_from_params(p, "interval"),), (p,), cam, fo2); print("frustum interval ok", flush=True)
