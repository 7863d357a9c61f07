#!/bin/bash
# GPU evidence for cast_rays_frustum: parity tests, timing, ncu launch list + one full capture of k_cast_frustum.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "frustum or render" > gpurun_out/r1g_frustum.log 2>&1; tail -3 gpurun_out/r1g_frustum.log
timeout 120 python tools/frustum_probe.py 1024 3 > gpurun_out/r1g_frustum_probe.txt 2>&1; cat gpurun_out/r1g_frustum_probe.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1g_frustum_launches.csv python tools/frustum_probe.py 1024 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cast_frustum -s 1 -c 1 -o gpurun_out/r1g_cast_frustum32 -f python tools/frustum_probe.py 1024 2 > gpurun_out/r1g_frustum_ncu.log 2>&1; tail -3 gpurun_out/r1g_frustum_ncu.log
