#!/bin/bash
# one-launch forward+backward against the two-launch step: equality test, then the bench both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-fb}
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail
for f in 0 1 0 1; do
  IGMC_FUSED_FB=$f timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_fb$f.json 2> gpurun_out/${T}_bench_fb$f.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_fb$f.json"))
print("fused=$f value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "warm", round(d["warm_l2"]["ms_per_step"], 4))
PY
done
