#!/bin/bash
# full capture of the FILTERED coarse GEMM at the C5 shard shape (second gemm_dist launch of a step)
mkdir -p gpurun_out
C5="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 3 --check 0"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dist -s 5 -c 1 \
    -o gpurun_out/r02_gemmf -f python scripts/bench_config.py $C5 > gpurun_out/r02_ncu_gemmf.log 2>&1
ls -la gpurun_out/r02_gemmf.ncu-rep
