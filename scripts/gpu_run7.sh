#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_forward_rs -s 4 -c 1 -o gpurun_out/prof_fwd -f python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_backward_rs -s 4 -c 1 -o gpurun_out/prof_bwd -f python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
