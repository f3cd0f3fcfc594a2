#!/bin/bash
# round 2, run 26: four-accumulator dot products in the generic / nested / type-10 kernels: shapes + generic tests
cd /root/repo
mkdir -p gpurun_out
timeout 900 python profiles/shape_bench.py > gpurun_out/r2s26_shapes.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s26_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/r2s26_suite.log
grep -v "^$" gpurun_out/r2s26_shapes.log | cut -c1-250; tail -n 4 gpurun_out/r2s26_suite.log
