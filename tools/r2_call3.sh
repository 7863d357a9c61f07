set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -k "intersection or cfg or grow or classify or c_abi or truncate or append or rays_golden or render" > gpurun_out/r2c_pytest.log 2>&1
tail -40 gpurun_out/r2c_pytest.log
timeout 600 python tools/r2_probe.py time > gpurun_out/r2c_probe.json 2> gpurun_out/r2c_probe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_probe.json'))
for k,v in d.items():
    if 'isect' in k or 'cfg3' in k or 'cast_rays' in k or 'hmc' in k: print(k, v)
PY
tail -5 gpurun_out/r2c_probe.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_isect_persistent --launch-count 1 -o gpurun_out/r2c_isect_persistent python tools/r2_probe.py run isect_trunc64_disjoint > gpurun_out/r2c_ncu1.log 2>&1
NIQ_ISECT_LEGACY=1 timeout 300 $NCU -k regex:k_classify_grow --launch-skip 20 --launch-count 1 -o gpurun_out/r2c_classify_grow python tools/r2_probe.py run isect_trunc64_disjoint > gpurun_out/r2c_ncu2.log 2>&1
compute-sanitizer --tool memcheck python tools/r2_probe.py run isect_trunc64_disjoint > gpurun_out/r2c_memcheck_isect.txt 2>&1; tail -3 gpurun_out/r2c_memcheck_isect.txt
compute-sanitizer --tool memcheck python tools/r2_probe.py run tree_bunny_d12 > gpurun_out/r2c_memcheck_tree.txt 2>&1; tail -3 gpurun_out/r2c_memcheck_tree.txt
