#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-mma2}
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_full_batch.py tests/test_gpu_extract.py -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail
timeout 200 python scripts/phase_profile.py > gpurun_out/${T}_phase_timeline.txt 2>&1; sed -n 1,62p gpurun_out/${T}_phase_timeline.txt | grep -E "L0 gather done|L1 w0 fold|L1 mma done|readout done|lists\+readout|L3 dgrad mma|L0 cluster"
run() {
  env "$@" timeout 400 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ($*)"
  python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print("   value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "warm", round(d["warm_l2"]["ms_per_step"], 4), {k: round(v * 1e3, 1) for k, v in d["roofline"]["kernel_ms"].items()})
PY
}
run IGMC_FUSED_FB=1
run IGMC_FUSED_FB=0
run IGMC_FUSED_FB=0 IGMC_PDL=2
run IGMC_FUSED_FB=1
run IGMC_FUSED_FB=0
