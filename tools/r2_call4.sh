set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tanh.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -k "tanh or intersection or grow or truncate or append or classify_golden" > gpurun_out/r2d_pytest.log 2>&1
tail -40 gpurun_out/r2d_pytest.log
timeout 600 python tools/r2_probe.py time > gpurun_out/r2d_probe.json 2> gpurun_out/r2d_probe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_probe.json'))
for k,v in d.items():
    if 'isect' in k or 'cfg3' in k or 'cast_rays' in k: print(k, v)
PY
tail -5 gpurun_out/r2d_probe.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 1500 gpurun_out/r2d_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print(b['value'], b['roofline']['frac'])
for k,v in b['configs'].items():
    print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ('roofline','cpu_baseline','kernel','hbm_pass','parity','sample')}, round(v['roofline']['frac_algorithmic'],3))
PY
