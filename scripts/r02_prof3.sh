#!/bin/bash
# full ncu captures of the steady-state (B=1024) launches of the two heaviest kernels
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan3_kernel -s 14 -c 1 \
    -o gpurun_out/r02_scan3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_scan3.log 2>&1
tail -2 gpurun_out/r02_ncu_scan3.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:cand_finalize -s 14 -c 1 \
    -o gpurun_out/r02_fin -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_fin.log 2>&1
tail -2 gpurun_out/r02_ncu_fin.log
ls -la gpurun_out/*.ncu-rep
