#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_full_batch.py -q --tb=short -p no:cacheprovider 2>&1 | tail -3
for th in 1024 512; do
  IGMC_RS_THREADS=$th timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload flixster > gpurun_out/r02x_flixster_t$th.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/r02x_flixster_t$th.json"))
print("flixster threads $th: value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), {k: round(v * 1000, 1) for k, v in d["roofline"]["kernel_ms"].items()})
PY
done
IGMC_RS_THREADS=512 timeout 300 python -m pytest tests/test_gpu_full_batch.py -q --tb=short -p no:cacheprovider -k flixster 2>&1 | tail -2
