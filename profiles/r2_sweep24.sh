#!/bin/bash
# DMMA mode products wired into the generic / nested kernels: parity tests, HCN and HNO3 shapes with and without
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2s24.txt
for d in 1 0; do echo "## EVR_SG4_DMMA=$d" >> $O/r2s24.txt; EVR_SG4_DMMA=$d timeout 600 python profiles/shape_bench.py 2>>$O/r2s24_err.log | head -4 >> $O/r2s24.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:sg4_term_kernel_generic -c 12 --csv python profiles/gen_case.py hcn 27 2>&1 | grep '^"' | tail -16 | cut -d, -f5,9,13,15 >> $O/r2s24.txt
cat $O/r2s24.txt; tail -3 $O/r2s24_err.log
