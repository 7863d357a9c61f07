set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1f_smi.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r1f_pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --extra > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_ref.json 2> gpurun_out/r1f_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1f_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cast_rays -s 3 -c 1 -o gpurun_out/r1f_cast_rays256 python bench.py --steps 1 --warmup 3 --no-cpu --tiles 18 > gpurun_out/r1f_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1
tail -5 gpurun_out/r1f_pytest_gpu.log; cut -c1-900 gpurun_out/r1f_bench.json; tail -3 gpurun_out/r1f_bench.err; cut -c1-300 gpurun_out/r1f_bench_ref.json; cat gpurun_out/r1f_smoke.log
