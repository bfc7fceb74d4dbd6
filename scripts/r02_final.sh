#!/bin/bash
# end-of-round artefacts: full bench line, reference arm, launch list of one C2 step, full capture of the scan kernel
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c2_reference_arm.json 2> gpurun_out/r02_ref.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:4lgpu -c 3000 --csv \
    --log-file gpurun_out/r02_launches_all.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:scan3_kernelILb0E -s 6 -c 1 \
    -o gpurun_out/r02_scan3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_scan3.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:cand_rescore -s 6 -c 1 \
    -o gpurun_out/r02_fin -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_fin.log 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/design_numbers.py gpurun_out/r02_bench_c2.json
cat gpurun_out/r02_bench_c2_reference_arm.json | cut -c1-400
