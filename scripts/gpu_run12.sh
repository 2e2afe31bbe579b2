#!/bin/bash
# full GPU suite + bench lines of the new workloads / model (bounded)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dgcnn.py tests/test_gpu_train.py -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -5
for cfg in "ml_1m igmc" "flixster igmc" "ml_1m dgcnn_rs" "ml_100k igmc"; do
  set -- $cfg
  timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload $1 --model $2 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "== $cfg rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$1_$2.json"))
    print(round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel_ms"], d["batch_stats"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_$1_$2.err").read()[-1500:])
PY
done
