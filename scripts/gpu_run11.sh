#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "exit $?" >> gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log | grep -v "Warning\|Consider\|^$\|Docs\|assert abs"
for nt in 1024 512; do
IGMC_RS_THREADS=$nt timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu-baseline > gpurun_out/bench_$nt.log 2> gpurun_out/bench_$nt.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$nt.log').read().strip().splitlines()[-1])
    print('threads $nt value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'warm',d['warm_l2']['value']); print(d['roofline']['kernel_ms'])
except Exception as e: print('bench parse fail',e); print(open('gpurun_out/bench_$nt.err').read()[-2000:])
PY
done
IGMC_RS_THREADS=512 timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -3
