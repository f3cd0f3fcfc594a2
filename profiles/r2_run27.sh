#!/bin/bash
# round 2, run 27: generic kernel at 64 (default) / 80 / 128 registers per thread (launch bounds 256 x 4 / 3 / 2)
cd /root/repo
mkdir -p gpurun_out
for v in default g3 g2; do
  echo "## generic kernel build: $v"
  if [ $v = default ]; then timeout 600 python profiles/shape_bench.py 2>&1 | head -4
  else EVR_SG4_LIB=/root/repo/variants/libevr_$v.so timeout 600 python profiles/shape_bench.py 2>&1 | head -4; fi
done > gpurun_out/r2s27_regs.log 2>&1
cut -c1-200 gpurun_out/r2s27_regs.log
