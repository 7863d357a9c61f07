set -x
mkdir -p gpurun_out
build/probes/ffma2_reuse_probe build/probes > gpurun_out/r1_ffma2_reuse_probe.txt 2>&1
cat gpurun_out/r1_ffma2_reuse_probe.txt
timeout 600 python tools/engine_probe.py build/libniq_v0_O3.so build/libniq_v8_O1.so build/libniq_v16_O1.so build/libniq_v16_O3.so build/libniq_v8_O3.so > gpurun_out/r1_engine_variants.txt 2>&1
cat gpurun_out/r1_engine_variants.txt
