#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:scan3_kernel|scan2_kernel|tile_desc|select|dist_matrix|group_|normalize|pair_distance|gemm_dist|bf16|band|threshold|overflow|filter_dense|qtable|probe_terms|pq_rescore|pack_records|count_below|finalize|slack'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 180 -c 90 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02_ncu_bench.log
