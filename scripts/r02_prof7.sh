#!/bin/bash
# launch list of the C5-shaped shard (library kernels only), default modes
mkdir -p gpurun_out
C5="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 3 --check 0"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:4lgpu -c 600 --csv \
    --log-file gpurun_out/r02_launches_c5.csv python scripts/bench_config.py $C5 > gpurun_out/r02_ncu_c5.log 2>&1
tail -c 300 gpurun_out/r02_ncu_c5.log
