set -x
timeout 600 python -m pytest tests -m gpu -q -k "min_distance or unsupported or slope" 2>&1 | tail -6
for i in 1 2; do timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_cfgvar_$i.json 2>gpurun_out/r2_cfgvar_$i.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_cfgvar_$i.json').read().strip().splitlines()[-1])
c=d['configs']
print('RUN $i', 'cfg3', {k:round(c['cfg3_hammer_x_bunny_intersection_truncate64'][k],2) for k in ('mean_ms','max_ms','median_ms')}, 'batched', round(c['cfg3_batched_64_transforms']['queries_per_s']), 'cfg4 2048', round(c['cfg4_birdcage_closest_point_window2048']['ms']), 'ge', round(c['cfg4_birdcage_closest_point_window_ge_stack']['ms']), 'd14', round(d['tree']['cfg5_depth14']['ms'],2), c['clocks'])
PY
done
