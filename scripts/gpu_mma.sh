#!/bin/bash
# tile-loop changes: parity tests, in-kernel phase timeline, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-mma}
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_full_batch.py -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail
timeout 200 python scripts/phase_profile.py > gpurun_out/${T}_phase_timeline.txt 2>&1; grep -E "mma|wgrad|readout done|cluster synced" gpurun_out/${T}_phase_timeline.txt | head -40
timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "warm", round(d["warm_l2"]["ms_per_step"], 4), d["roofline"]["kernel_ms"])
PY
