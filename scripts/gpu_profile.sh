#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench line, ncu launch list, one full ncu capture of
# the scan kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
# launch list (cold-cache, serialised: compare shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lgpu' -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 2 -f -o gpurun_out/scan_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_scan.log 2>&1
ls -la gpurun_out
