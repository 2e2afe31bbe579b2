#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short --timeout=100 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_model.log 2>&1
echo "exit $?" >> gpurun_out/t_model.log
timeout 200 python scripts/gpu_smoke.py > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
timeout 500 python bench.py --steps 100 --warmup 5 --cpu-steps 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "exit $?" >> gpurun_out/ncu_bench.log
tail -40 gpurun_out/t_model.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
