#!/bin/bash
mkdir -p gpurun_out
ARGS="--n 100000000 --nlist 16384 --nprobes 20 --k 10 --batch 8192 --metric l2 --owned 0.125 --steps 5 --check 4"
timeout -s KILL 600 python scripts/bench_config.py $ARGS 2>&1 | tail -2
timeout -s KILL 600 python scripts/bench_config.py --n 10000000 --nlist 4096 --nprobes 50 --k 100 --batch 4096 --metric cosine --steps 4 --check 4 2>&1 | tail -1
LGPU_CAND_KMAX=128 timeout -s KILL 600 python scripts/bench_config.py --n 10000000 --nlist 4096 --nprobes 50 --k 100 --batch 4096 --metric cosine --steps 4 --check 4 2>&1 | tail -1
