#!/bin/bash
# round 2: DMMA vs DFMA micro-benchmark (with ncu pipe counters), GPU tests of the new C-ABI rows
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
./profiles/micro/dmma_vs_dfma > $O/r2s11_dmma.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none --csv ./profiles/micro/dmma_vs_dfma > $O/r2s11_dmma_ncu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/r2s11_dmma.txt
cat $O/r2s11_dmma.txt; grep -v "^==" $O/r2s11_dmma_ncu.csv | head -60
