#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-s}
line() {  # workload plan
  IGMC_PLAN=$2 timeout 300 python bench.py --steps 100 --warmup 10 --skip-cpu-baseline --workload $1 > gpurun_out/${T}_$1_p$2.json 2> gpurun_out/${T}_$1_p$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_$1_p$2.json"))
    print("$1 plan $2: value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4), {k: round(v * 1000, 1) for k, v in d["roofline"]["kernel_ms"].items()})
except Exception as e:
    print("$1 plan $2 ERR", e)
PY
}
line ml_100k 2
line ml_100k 4
line ml_100k 3
line flixster 1
line flixster 2
line flixster 4
