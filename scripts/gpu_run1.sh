#!/bin/bash
# first GPU contact: sanitizer on the smoke path, then the two GPU test files in separate processes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python scripts/gpu_smoke.py > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
timeout 900 python -m pytest tests/test_gpu_extract.py -q -m gpu -x --tb=short > gpurun_out/t_extract.log 2>&1
echo "exit $?" >> gpurun_out/t_extract.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/t_model.log 2>&1
echo "exit $?" >> gpurun_out/t_model.log
tail -5 gpurun_out/sanitizer.log; tail -15 gpurun_out/t_extract.log; tail -30 gpurun_out/t_model.log
