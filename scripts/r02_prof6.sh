#!/bin/bash
# full captures: coarse_finish at the C5 shard shape, cand_rescore at C3
mkdir -p gpurun_out
C5="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 3 --check 0"
C3="--n 10000000 --nlist 4096 --nprobes 50 --k 100 --batch 4096 --metric cosine --steps 3 --check 0"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:coarse_finish -s 3 -c 1 \
    -o gpurun_out/r02_cfin -f python scripts/bench_config.py $C5 > gpurun_out/r02_ncu_cfin.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dist -s 3 -c 1 \
    -o gpurun_out/r02_gemm5 -f python scripts/bench_config.py $C5 > gpurun_out/r02_ncu_gemm5.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:cand_rescore -s 3 -c 1 \
    -o gpurun_out/r02_rsc3 -f python scripts/bench_config.py $C3 > gpurun_out/r02_ncu_rsc3.log 2>&1
ls -la gpurun_out/*.ncu-rep
