#!/bin/bash
# A/B of the persistent kernel's switches through its phase trace (stand-alone decoder forward)
mkdir -p gpurun_out
for cfg in "B2C_RECUR_EARLY=1 B2C_RECUR_EPNC=0" "B2C_RECUR_EARLY=0 B2C_RECUR_EPNC=0" "B2C_RECUR_EARLY=0 B2C_RECUR_EPNC=1"; do
  echo "=== $cfg"
  env $cfg timeout 300 python tools/recur_trace.py 2>&1 | grep -E "grid|epi|mma|producer|Error|error"
done
exit 0
