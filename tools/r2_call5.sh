set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tanh.py tests/test_gpu_parity.py -m gpu -q -k "tanh or frustum or grow or render" > gpurun_out/r2e_pytest.log 2>&1
tail -40 gpurun_out/r2e_pytest.log
