#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ver2}
timeout 600 python -m pytest tests -q -m gpu --tb=short --timeout=300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail -5
for i in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "warm", round(d["warm_l2"]["ms_per_step"], 4), {k: round(v * 1e3, 1) for k, v in d["roofline"]["kernel_ms"].items()}, d["roofline"]["traffic"], d["gpu_launches"])
PY
done
