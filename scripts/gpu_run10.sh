#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "exit $?" >> gpurun_out/t_all.log
tail -25 gpurun_out/t_all.log | grep -v "Warning\|Consider\|^$\|Docs\|assert abs"
timeout 300 python bench.py --steps 200 --warmup 10 --cpu-steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'warm',d['warm_l2']['value']); print(d['roofline']['kernel_ms'])
except Exception as e: print('bench parse fail',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
