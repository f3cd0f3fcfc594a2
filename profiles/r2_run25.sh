#!/bin/bash
# round 2, run 25: fast-path plans with a generic remainder (HH L=8), pipelined host blocks, full suite
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "remainder or pipelined or larger_than_shared or beyond_smem" > gpurun_out/r2s25_new.log 2>&1; echo "new rc=$?" >> gpurun_out/r2s25_new.log
timeout 900 python profiles/hh_L8_check.py 8 > gpurun_out/r2s25_L8.log 2>&1; echo "L8 rc=$?" >> gpurun_out/r2s25_L8.log
(timeout 600 python profiles/block_e2e.py 7 8; EVR_SG4_PIPELINE=0 timeout 600 python profiles/block_e2e.py 7 8) > gpurun_out/r2s25_block.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s25_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/r2s25_suite.log
timeout 600 python bench.py > gpurun_out/r2s25_bench.json 2> gpurun_out/r2s25_bench.err
tail -n 6 gpurun_out/r2s25_new.log; tail -n 3 gpurun_out/r2s25_L8.log; cat gpurun_out/r2s25_block.log | tail -n 4; tail -n 4 gpurun_out/r2s25_suite.log; cat gpurun_out/r2s25_bench.json | cut -c1-600
