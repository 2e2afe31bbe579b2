#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in 0 1; do
  echo "=== IGMC_FUSED_FB=$f"
  IGMC_FUSED_FB=$f timeout 200 python scripts/step_timeline.py 2>&1 | grep -A4 "back-to-back replay [123]"
done
