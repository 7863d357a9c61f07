set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
tail -60 gpurun_out/r2b_pytest.log
timeout 600 python tools/r2_probe.py time > gpurun_out/r2b_probe.json 2> gpurun_out/r2b_probe.err
cat gpurun_out/r2b_probe.json; tail -5 gpurun_out/r2b_probe.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 3000 gpurun_out/r2b_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -2 gpurun_out/r2b_smoke.log
