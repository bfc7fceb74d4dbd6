#!/bin/bash
# multi-GPU check: sharded NCCL test + bench at N ranks (replicas value + sharded / c5 blocks)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -5 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
j=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:j[k] for k in ('value','ms_per_step','n_gpus')}); print('sharded',j.get('sharded')); print('c5',j.get('c5'))
PY
