#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -8
ARGS="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 3 --check 0"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:coarse_finish -s 3 -c 1 \
    -o gpurun_out/r02_cfin -f python scripts/bench_config.py $ARGS > gpurun_out/r02_ncu_cfin.log 2>&1
tail -2 gpurun_out/r02_ncu_cfin.log
