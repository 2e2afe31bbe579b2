#!/bin/bash
# 1-GPU lines of every BASELINE workload + the headline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-w}
timeout 300 python -m pytest tests/test_gpu_train.py -q --tb=short -p no:cacheprovider 2>&1 | tail -2
for cfg in "ml_1m igmc" "ml_100k igmc" "flixster igmc" "ml_1m_r02 igmc" "ml_1m dgcnn_rs"; do
  set -- $cfg
  timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload $1 --model $2 > gpurun_out/${T}_bench_$1_$2.json 2> gpurun_out/${T}_bench_$1_$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$1_$2.json"))
    print("$1 $2: value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), {k: round(v * 1000, 1) for k, v in d["roofline"]["kernel_ms"].items()}, "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$1 $2 ERR", e)
PY
done
