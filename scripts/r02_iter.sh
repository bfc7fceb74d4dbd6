#!/bin/bash
# round 2 iteration check: parity tests of the scan paths, short bench, launch list of one step
mkdir -p gpurun_out
TAG=${1:-iter}
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_fullsize.py tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -6
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}); print(j['stage_ms'], j.get('filter_stats')); print('frac',j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['pipelined']['value'])
PY
KREGEX='regex:scan3_kernel|scan2_kernel|tile_desc|select|dist_matrix|group_|normalize|pair_distance|gemm_dist|bf16|band|threshold|overflow|filter_dense|qtable|probe_terms|pq_rescore|pack_records|count_below|finalize|slack|coarse'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 300 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv
lines=[l for l in open('gpurun_out/${TAG}_launches.csv') if not l.startswith('==')]
r=list(csv.DictReader(lines))
for x in r[:40]:
    print(x['ID'], x['Kernel Name'][:60], x['Metric Value'])
PY
