#!/bin/bash
# multi-GPU validation: DP identity test (peer exchange) + bench under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; T=${2:-m}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_train.py -q --tb=short -p no:cacheprovider -k "dp2 or fused_update" 2>&1 | tail -5
for n in $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_${n}gpu.json 2> gpurun_out/${T}_bench_${n}gpu.err
  echo "bench $n rc=$?"; tail -2 gpurun_out/${T}_bench_${n}gpu.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_${n}gpu.json"))
    print("N=$n value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["runs_ms_per_step"])
except Exception as e:
    print("ERR", e)
PY
done
