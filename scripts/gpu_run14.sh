#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in k_forward_rs k_backward_rs; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r01_v4_$k \
    python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"; tail -3 gpurun_out/ncu_$k.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_v4_launches.csv \
  python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | tail -8
