#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_extract.py -q -m gpu --tb=short --timeout=120 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "exit $?" >> gpurun_out/t_all.log
timeout 200 python scripts/phase_profile.py > gpurun_out/phases.log 2>&1
timeout 500 python bench.py --steps 200 --warmup 10 --cpu-steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.err
tail -8 gpurun_out/t_all.log; head -62 gpurun_out/phases.log; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'warm',d['warm_l2']['value']); print(d['roofline']['kernel_ms'])
except Exception as e: print('bench parse fail',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
