# Round-2 closing run on one B200: full GPU suite, the driver's bench commands, ncu launch list + full captures, sanitizer
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2z_smi.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2z_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -c 600 gpurun_out/r2z_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2z_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:k_cast_rays -c 1 --launch-skip 3 -o gpurun_out/r2z_cast_rays256 python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --tiles-total 18 > gpurun_out/r2z_ncu_rays.log 2>&1
cat > /tmp/cp_small.py <<'PY'
import sys, numpy as np
sys.path[:0] = ["oracle", "neural-implicit-queries_b200", "."]
import implicit_mlp_utils, kd_tree, queries, render
with np.load("tests/golden/mlps.npz") as d:
    P = {nm: {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(nm + "/")} for nm in ("fox", "bunny", "birdcage_occ")}
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
what = sys.argv[1]
if what == "cp":
    p = P["birdcage_occ"]; f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
    q = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)[:32]
    print(kd_tree.closest_point(f, p, lo, hi, q, eps=1e-3)[0].sum())
else:
    p = P["fox"]; f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_truncate", affine_n_truncate=8, affine_truncate_policy="absolute")
    eye = np.array((2., 1., 2.), np.float32); look, up, left = render.look_at(eye)
    if what == "rays":
        r, d_ = render.generate_camera_rays(eye, look, up, res=48, fov_deg=30.)
        print(queries.cast_rays((f,), (p,), r, d_, queries.get_default_cast_opts())[2].sum())
    else:
        print(queries.cast_rays_frustum((f,), (p,), (eye, look, up, left, 30., 30., 48, 48), queries.get_default_cast_opts())[2].sum())
PY
timeout 600 $NCU -k regex:k_cp_persistent -c 1 -o gpurun_out/r2z_cp_persistent python /tmp/cp_small.py cp > gpurun_out/r2z_ncu_cp.log 2>&1
timeout 300 $NCU -k regex:k_cast_rays_grow -c 1 -o gpurun_out/r2z_cast_rays_grow python /tmp/cp_small.py rays > gpurun_out/r2z_ncu_raysgrow.log 2>&1
for w in cp rays frustum; do compute-sanitizer --tool memcheck python /tmp/cp_small.py $w > gpurun_out/r2z_memcheck_$w.txt 2>&1; tail -1 gpurun_out/r2z_memcheck_$w.txt; done
compute-sanitizer --tool synccheck python /tmp/cp_small.py frustum > gpurun_out/r2z_synccheck_frustum_grow.txt 2>&1; tail -1 gpurun_out/r2z_synccheck_frustum_grow.txt
ls -la gpurun_out/r2z_*
