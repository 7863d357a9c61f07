set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/r2_nccl_check.py > gpurun_out/r2f_nccl_check_n2.json 2> gpurun_out/r2f_nccl_check_n2.err
cat gpurun_out/r2f_nccl_check_n2.json; tail -5 gpurun_out/r2f_nccl_check_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -c 1500 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2f_bench_n2.json').read().strip().splitlines()[-1])
print('N', b['n_gpus'], 'value', b['value'], 'ms/step', b['ms_per_step'], 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'])
print(json.dumps(b['tree'], indent=1)[:1800])
PY
