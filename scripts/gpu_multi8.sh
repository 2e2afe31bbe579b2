#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-m8}
run() {  # n workload
  IGMC_BENCH_DEBUG=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $1 --steps 200 --warmup 10 --skip-cpu-baseline --workload $2 > gpurun_out/${T}_bench_$2_$1gpu.json 2> gpurun_out/${T}_bench_$2_$1gpu.err
  echo "bench $2 N=$1 rc=$?"; grep "^\[rank 0\]" gpurun_out/${T}_bench_$2_$1gpu.err | cut -c1-260
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$2_$1gpu.json"))
    print("  value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["runs_ms_per_step"])
except Exception as e:
    print("  ERR", e)
PY
}
run 8 ml_1m
run 1 ml_1m
run 8 ml_1m_r02
