set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "closest or cfg4 or ring or sample_surface" > gpurun_out/r2h_pytest.log 2>&1
tail -15 gpurun_out/r2h_pytest.log
python - <<'PY'
import os, sys, time, json
import numpy as np
sys.path[:0] = ["oracle", "neural-implicit-queries_b200", "."]
import implicit_mlp_utils, kd_tree
with np.load("tests/golden/mlps.npz") as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("birdcage_occ/")}
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
q = np.random.default_rng(0).uniform(-1, 1, (1000000, 3)).astype(np.float32)[:256]
out = {}
for legacy in ("0", "1"):
    os.environ["NIQ_CP_LEGACY"] = legacy
    kd_tree.closest_point(f, p, lo, hi, q[:32], eps=1e-3)
    st = {}
    t0 = time.perf_counter()
    d_, l_ = kd_tree.closest_point(f, p, lo, hi, q, eps=1e-3, stats=st)
    dt = time.perf_counter() - t0
    out["graph_loop" if legacy == "1" else "persistent"] = {"ms": dt * 1e3, "visits_per_s": st["n_visits"] / dt, **st, "sum": float(d_.sum())}
print(json.dumps(out, indent=1))
PY
