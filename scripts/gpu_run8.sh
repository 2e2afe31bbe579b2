#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short --timeout=180 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "exit $?" >> gpurun_out/t_all.log
tail -25 gpurun_out/t_all.log
