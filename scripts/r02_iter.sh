#!/bin/bash
# one iteration on the GPU box: coarse/tensor-core parity tests, C2 quick bench, C5-shard + C3 quick lines
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
bash scripts/r02_quick.sh it 2>&1 | grep -v Warn
bash scripts/r02_c5quick.sh 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        j = json.loads(l); print(j['config']['nlist'], 'ms', round(j['ms_per_batch'], 3), {k: round(v, 3) for k, v in j['stage_ms'].items()}, j.get('filter_stats'), j.get('oracle_check'))
"
