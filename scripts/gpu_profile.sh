#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench line, ncu launch list, full ncu captures of the scan
# kernel (HBM-bound claim) and of the tcgen05 GEMM (tensor-pipe claim).  Outputs under gpurun_out/.
# Every python call is wrapped in `timeout` so a hung kernel cannot hold the box.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
KREGEX='regex:scan2_kernel|scan_kernel|tile_desc|select|dist_matrix|group_|normalize|pair_distance|gemm|bf16|band'
# launch list (cold-cache, serialised: compare shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 80 -c 32 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan2_kernel -s 4 -c 1 -f -o gpurun_out/scan_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_scan.log 2>&1
# flat path: throughput line + full capture of the GEMM
timeout 600 python scripts/bench_flat.py > gpurun_out/flat.json 2> gpurun_out/flat.err; cat gpurun_out/flat.json
timeout 900 ncu --set full --clock-control none -k regex:gemm_dist -s 1 -c 1 -f -o gpurun_out/gemm_full \
    python scripts/bench_flat.py --steps 1 --check 0 > gpurun_out/ncu_gemm.log 2>&1
fi
ls -la gpurun_out
