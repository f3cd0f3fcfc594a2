#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r2s32.txt
for v in "EVR_X=0" "EVR_SG4_DETERMINISTIC=1" "EVR_SG4_FORCE_GENERIC=1" "EVR_SG4_MIXED=0"; do
  echo "## $v" >> gpurun_out/r2s32.txt
  env $v timeout 300 python -m pytest tests -m gpu -q -k "remainder or larger_than_shared or pipelined or complex_psi" 2>&1 | tail -3 | cut -c1-200 >> gpurun_out/r2s32.txt
done
cat gpurun_out/r2s32.txt
