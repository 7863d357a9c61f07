set -x
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/r2_nccl_check.py > gpurun_out/r2_nccl_check_n$N.json 2> gpurun_out/r2_nccl_check_n$N.err
cat gpurun_out/r2_nccl_check_n$N.json; tail -3 gpurun_out/r2_nccl_check_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 800 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('N', b['n_gpus'], 'value', b['value'], 'ms/step', b['ms_per_step'], 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'])
t=b['tree']['bunny_depth21']; print('d21', t['value'], t['ms'], 'e2e', t['e2e'])
PY
