#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_extract.py -q -m gpu --tb=short --timeout=120 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_extract.log 2>&1
echo "exit $?" >> gpurun_out/t_extract.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short --timeout=120 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_model.log 2>&1
echo "exit $?" >> gpurun_out/t_model.log
timeout 200 python scripts/gpu_smoke.py > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
timeout 500 python bench.py --steps 200 --warmup 10 --cpu-steps 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "exit $?" >> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_backward_rs -s 4 -c 1 -o gpurun_out/prof_bwd -f python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_forward_rs -s 4 -c 1 -o gpurun_out/prof_fwd -f python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_fwd.log 2>&1
tail -4 gpurun_out/t_extract.log; tail -30 gpurun_out/t_model.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
