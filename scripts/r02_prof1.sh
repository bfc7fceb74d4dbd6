#!/bin/bash
# launch list of the bench (the library's own kernels only: mangled names carry the lgpu namespace) + a full capture
# of the filter scan kernel; scripts/launch_steps.py cuts one steady-state step out of the list
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:4lgpu -c 3000 --csv \
    --log-file gpurun_out/r02_launches_all.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02_ncu_bench.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan3_kernel -s 4 -c 1 \
    -o gpurun_out/r02_scan3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_scan3.log 2>&1
tail -2 gpurun_out/r02_ncu_scan3.log
ls -la gpurun_out/r02_scan3.ncu-rep
