#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-q}
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}); print(j['stage_ms'], j.get('filter_stats')); print('frac',j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['pipelined']['value'])
PY
