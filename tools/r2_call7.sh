set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "tree or ring or cfg5" > gpurun_out/r2g_pytest.log 2>&1
tail -15 gpurun_out/r2g_pytest.log
timeout 300 python tools/r2_probe.py time > gpurun_out/r2g_probe.json 2> gpurun_out/r2g_probe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_probe.json'))
for k,v in d.items():
    if 'tree' in k: print(k, v)
PY
