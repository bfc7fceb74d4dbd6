#!/bin/bash
# round 2, first GPU pass: smoke, GPU suite, bench N=1 (with extras), reference arm
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
tail -3 gpurun_out/r02a_bench_ref.err
timeout -s KILL 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_bench.err
wc -c gpurun_out/r02a_bench.json gpurun_out/r02a_bench_ref.json
