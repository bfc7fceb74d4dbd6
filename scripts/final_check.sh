#!/bin/bash
# round-end style check on the GPU box: smoke, GPU suite, reference arm, bench line, launch list
mkdir -p gpurun_out
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout -s KILL 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
wc -l gpurun_out/bench.json gpurun_out/bench_ref.json
KREGEX='regex:scan2_kernel|scan_kernel|tile_desc|select|dist_matrix|group_|normalize|pair_distance|gemm|bf16|band|threshold|overflow|filter_dense'
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 96 -c 32 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
