#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_dgcnn.py tests/test_gpu_train.py -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -4
timeout 200 python scripts/phase_profile.py > gpurun_out/phase22.txt 2>&1; head -66 gpurun_out/phase22.txt | cut -c1-100
for w in ml_1m flixster; do
  timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload $w > gpurun_out/bench22_$w.json 2> gpurun_out/bench22_$w.err
  echo "== $w rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench22_$w.json"))
    print(round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel_ms"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench22_$w.err").read()[-1500:])
PY
done
