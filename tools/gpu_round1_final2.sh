# Round-1 closing evidence (after cast_rays_frustum / render_image / barrier-free streamed exit): files r1z_*
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1z_smi.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r1z_pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1z_smoke.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --extra > gpurun_out/r1z_bench.json 2> gpurun_out/r1z_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1z_bench_ref.json 2> gpurun_out/r1z_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1z_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cast_rays -s 3 -c 1 -f -o gpurun_out/r1z_cast_rays256 python bench.py --steps 1 --warmup 3 --no-cpu --tiles 18 > gpurun_out/r1z_ncu_full.log 2>&1
timeout 200 compute-sanitizer --tool synccheck python tools/frustum_probe.py 48 1 > gpurun_out/r1z_synccheck_frustum.txt 2>&1
tail -5 gpurun_out/r1z_pytest_gpu.log; cut -c1-600 gpurun_out/r1z_bench.json; tail -3 gpurun_out/r1z_bench.err; cut -c1-300 gpurun_out/r1z_bench_ref.json; cat gpurun_out/r1z_smoke.log; tail -2 gpurun_out/r1z_synccheck_frustum.txt
