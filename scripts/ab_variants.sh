#!/bin/bash
# A/B of kernel build variants on the GPU box: full GPU suite on the default library, the parity file and
# one bench line per variant library (built with csrc/Makefile OUT=... EXTRA=...).
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
for v in "$@"; do
    export LGPU_LIB_PATH=$PWD/lancedb_b200/_lib/$v.so
    timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
    timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
