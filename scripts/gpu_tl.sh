#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t}
timeout 200 python scripts/step_timeline.py 2>&1 | grep -A4 "back-to-back replay [23]"
timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["runs_ms_per_step"])
PY
IGMC_PDL=0 timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu_nopdl.json 2> /dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu_nopdl.json"))
print("no-PDL value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
