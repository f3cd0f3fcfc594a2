#!/bin/bash
# round 2, run 28: all-reduce with in-kernel barriers, one device (4 "ranks" on streams)
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "in_kernel_barriers or peer" > gpurun_out/r2s28_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2s28_fused.log
tail -n 8 gpurun_out/r2s28_fused.log
