# Round-2 closing run on one B200 (after the dealt tree build / fingerprint / bench changes): full GPU suite, smoke, both
# bench arms as the driver runs them, ncu launch list, one full capture of the dealt tree kernel, sanitizer on the dealt build
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2e_smi.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2e_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
tail -c 600 gpurun_out/r2e_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2e_bench_under_ncu.log 2>&1
cat > /tmp/dealt_small.py <<'PY'
import sys, numpy as np
sys.path[:0] = ["oracle", "neural-implicit-queries_b200", "."]
import implicit_mlp_utils, kd_tree
with np.load("tests/golden/mlps.npz") as d:
    p = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith("bunny/")}
lo, hi = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
f = implicit_mlp_utils.generate_implicit_from_params(p, "affine_fixed")
depth = int(sys.argv[1])
n = 0
for r in range(2):
    t = kd_tree.build_tree_dealt(f, p, lo, hi, depth, min(12, depth - 2), r, 8)
    n += t.count(0); t.close()
print(n)
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:k_tree_persistent -c 1 -o gpurun_out/r2e_tree_dealt python /tmp/dealt_small.py 21 > gpurun_out/r2e_ncu_tree.log 2>&1
compute-sanitizer --tool memcheck python /tmp/dealt_small.py 15 > gpurun_out/r2e_memcheck_tree_dealt.txt 2>&1; tail -1 gpurun_out/r2e_memcheck_tree_dealt.txt
compute-sanitizer --tool synccheck python /tmp/dealt_small.py 15 > gpurun_out/r2e_synccheck_tree_dealt.txt 2>&1; tail -1 gpurun_out/r2e_synccheck_tree_dealt.txt
ls -la gpurun_out/r2e_*
