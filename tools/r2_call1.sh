set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -x -k "tree or cfg or strict or intersection or closest" > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/r2_probe.py time > gpurun_out/r2a_probe.json 2> gpurun_out/r2a_probe.err
cat gpurun_out/r2a_probe.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_tree_persistent -o gpurun_out/r2a_tree_persistent64 python tools/r2_probe.py run tree_bunny_d21 > gpurun_out/r2a_ncu1.log 2>&1
NIQ_TREE_LEGACY=1 timeout 300 $NCU -k regex:k_classify_fixed --launch-skip 19 --launch-count 1 -o gpurun_out/r2a_classify_fixed64 python tools/r2_probe.py run tree_bunny_d21 > gpurun_out/r2a_ncu2.log 2>&1
NIQ_TREE_LEGACY=1 timeout 300 $NCU -k regex:"k_tree_scatter|k_scan_apply" --launch-skip 40 --launch-count 4 -o gpurun_out/r2a_tree_hbm python tools/r2_probe.py run tree_bunny_d21 > gpurun_out/r2a_ncu3.log 2>&1
timeout 400 $NCU -k regex:"k_eval_points|k_mc_write|k_mc_count" --launch-count 3 -o gpurun_out/r2a_hmc python tools/r2_probe.py run hmc_bunny_d9 > gpurun_out/r2a_ncu4.log 2>&1
timeout 300 $NCU -k regex:k_classify_grow --launch-skip 20 --launch-count 1 -o gpurun_out/r2a_classify_grow python tools/r2_probe.py run isect_trunc64_disjoint > gpurun_out/r2a_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
