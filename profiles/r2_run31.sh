#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s31_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/r2s31_suite.log
tail -n 6 gpurun_out/r2s31_suite.log
