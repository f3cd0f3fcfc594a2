#!/bin/bash
# round 2, run 24: terms beyond shared memory (generic + nested kernels), HH 12-D L=8 size check, full GPU suite
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "larger_than_shared or beyond_smem" > gpurun_out/r2s24_big.log 2>&1; echo "big rc=$?" >> gpurun_out/r2s24_big.log
timeout 900 python profiles/hh_L8_check.py 8 > gpurun_out/r2s24_L8.log 2>&1; echo "L8 rc=$?" >> gpurun_out/r2s24_L8.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s24_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/r2s24_suite.log
tail -5 gpurun_out/r2s24_big.log gpurun_out/r2s24_L8.log gpurun_out/r2s24_suite.log
