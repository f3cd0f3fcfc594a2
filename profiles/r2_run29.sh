#!/bin/bash
# round 2, run 29: size-class thresholds of the generic kernel (values per term -> CTA of 256 / 128 / 64 / 32 threads)
cd /root/repo
mkdir -p gpurun_out
for v in "768 192 48" "512 128 32" "256 128 32" "256 64 32" "1024 256 64" "384 96 32"; do
  set -- $v
  echo "## EVR_SG4_GTH0=$1 EVR_SG4_GTH1=$2 EVR_SG4_GTH2=$3"
  EVR_SG4_GTH0=$1 EVR_SG4_GTH1=$2 EVR_SG4_GTH2=$3 timeout 600 python profiles/shape_bench.py 2>&1 | head -4 | cut -c1-150
done > gpurun_out/r2s29_gth.log 2>&1
cat gpurun_out/r2s29_gth.log
