#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-b}
timeout 600 python -m pytest tests/test_gpu_full_batch.py tests/test_gpu_main_script.py tests/test_gpu_train.py -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail
timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["runs_ms_per_step"], d["roofline"]["kernel_ms"])
PY
