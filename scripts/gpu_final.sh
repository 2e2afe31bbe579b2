#!/bin/bash
# round-end validation on one B200: full GPU suite, smoke, bench lines of every workload, reference arm, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short --timeout=240 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
for cfg in "ml_100k igmc" "ml_1m_r02 igmc" "flixster igmc" "ml_1m dgcnn_rs"; do
  set -- $cfg
  timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload $1 --model $2 > gpurun_out/final_bench_$1_$2.json 2> gpurun_out/final_bench_$1_$2.err
  echo "== $cfg rc=$?"
done
for k in k_forward_rs k_backward_rs; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r01_final_$k \
    python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_final_launches.csv \
  python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/launches.log 2>&1
echo "launch list rc=$?"
timeout 200 python scripts/phase_profile.py > gpurun_out/r01_final_phase_timeline.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("final_bench_")[1], round(d["value"], 1), round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), d.get("roofline", {}).get("kernel_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
