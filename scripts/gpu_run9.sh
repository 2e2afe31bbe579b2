#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -k full_size --tb=short --timeout=120 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_full.log 2>&1
echo "exit $?" >> gpurun_out/t_full.log
tail -30 gpurun_out/t_full.log | grep -v "Warning\|Consider\|^$\|Docs"
for w in ml_100k ml_1m_r02; do
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --skip-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err
  echo "$w rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
    print('$w value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value']); print(d['roofline']['kernel_ms'], d['batch_stats'])
except Exception as e: print('parse fail',e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
