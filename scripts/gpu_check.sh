#!/bin/bash
# quick GPU iteration: full GPU suite, smoke, short headline bench, in-kernel phase timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02}
timeout 1200 python -m pytest tests -q -m gpu --tb=short --timeout=300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest.txt | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --steps 100 --warmup 10 --skip-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench_1gpu.err
timeout 200 python scripts/phase_profile.py > gpurun_out/${T}_phase_timeline.txt 2>&1; echo "phase rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["roofline"]["kernel_ms"])
except Exception as e:
    print("bench ERR", e)
PY
timeout 300 python scripts/step_breakdown.py > gpurun_out/${T}_step_breakdown.txt 2>&1; echo "breakdown rc=$?"; cat gpurun_out/${T}_step_breakdown.txt | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_extract_fast -s 6 -c 1 -f -o gpurun_out/${T}_k_extract_fast \
    python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_k_extract_fast.log 2>&1; echo "ncu extract rc=$?"
timeout 200 python scripts/step_timeline.py > gpurun_out/${T}_step_timeline.txt 2>&1; echo "timeline rc=$?"; tail -12 gpurun_out/${T}_step_timeline.txt
