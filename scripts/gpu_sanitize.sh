#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/gpu_smoke.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
