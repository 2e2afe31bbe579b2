#!/bin/bash
# round-end evidence after the tile-loop / one-launch changes (trimmed version of gpu_final.sh: the reference arm and the
# extraction / update-kernel captures of r02_final_* are unchanged by them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r02b}
timeout 900 python -m pytest tests -q -m gpu --tb=short --timeout=300 --timeout-method=thread -p no:cacheprovider > gpurun_out/${R}_final_pytest.txt 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${R}_final_pytest.txt | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${R}_final_bench_1gpu.json 2> gpurun_out/${R}_final_bench_1gpu.err; echo "bench rc=$?"
for cfg in "ml_100k igmc" "ml_1m_r02 igmc" "flixster igmc" "ml_1m dgcnn_rs"; do
  set -- $cfg
  timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline --workload $1 --model $2 > gpurun_out/${R}_final_bench_$1_$2.json 2> gpurun_out/${R}_final_bench_$1_$2.err
  echo "== $cfg rc=$?"
done
for k in k_train_rs k_reduce_allreduce_adam; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${R}_final_$k \
    python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_final_launches.csv \
  python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/launches.log 2>&1
echo "launch list rc=$?"
timeout 200 python scripts/phase_profile.py > gpurun_out/${R}_final_phase_timeline.txt 2>&1
timeout 200 python scripts/step_timeline.py > gpurun_out/${R}_final_step_timeline.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02b_final_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("final_bench_")[1], round(d["value"], 1), round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), (d.get("roofline") or {}).get("kernel_ms"), "frac", (d.get("roofline") or {}).get("frac"), "gpu_baseline", (d.get("gpu_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
