#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-q}
for th in 1024 512 256; do
  echo "== IGMC_EX_THREADS=$th"
  IGMC_EX_THREADS=$th timeout 400 python scripts/step_breakdown.py 2>&1 | grep -E "eager|model branch|extraction branch|both branches  " 
done
