#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-m8}
IGMC_BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 100 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_dbg_8gpu.json 2> gpurun_out/${T}_dbg_8gpu.err
echo "rc=$?"; grep "^\[rank" gpurun_out/${T}_dbg_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_dbg_8gpu.json')); print('value', round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['runs_ms_per_step'], 'warm', d['warm_l2']['ms_per_step'])"
