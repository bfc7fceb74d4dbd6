#!/bin/bash
# C5-shaped shard (1/8 of 100M x 768, nlist 16384, B 8192) on one GPU: timing + launch list
mkdir -p gpurun_out
ARGS="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 5 --check 4"
timeout -s KILL 600 python scripts/bench_config.py $ARGS > gpurun_out/r02_c5shard.json 2> gpurun_out/r02_c5shard.err
tail -2 gpurun_out/r02_c5shard.err; cat gpurun_out/r02_c5shard.json
KREGEX='regex:scan3_kernel|scan2_kernel|tile_desc|select|dist_matrix|group_|normalize|pair_distance|gemm_dist|bf16|band|threshold|overflow|filter_dense|qtable|probe_terms|pq_rescore|pack_records|count_below|cand_|coarse'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 60 -c 40 --csv \
    --log-file gpurun_out/r02_c5_launches.csv python scripts/bench_config.py $ARGS --check 0 > gpurun_out/r02_c5_ncu.log 2>&1
python - <<PY
import csv
lines=[l for l in open('gpurun_out/r02_c5_launches.csv') if not l.startswith('==')]
r=list(csv.DictReader(lines))
for x in r[:40]:
    print(x['ID'], x['Kernel Name'][:60], x['Metric Value'])
PY
