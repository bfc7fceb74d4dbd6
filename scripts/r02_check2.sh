#!/bin/bash
# round 2: GPU suite on the filter-scan build + short bench
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -5 gpurun_out/r02b_bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02b_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','stage_ms')}, j['roofline']['frac'], j.get('e2e'), j.get('gate'))
PY
