#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s7_sweep.txt
run() { echo "## $*" >> $O/r2s7_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s7_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s7_sweep.txt 2>&1; }
run EVR_SG4_V2_THREADS=512 EVR_SG4_SMEM_KB=200 EVR_SG4_BCAP=3300
run EVR_SG4_V2_THREADS=512 EVR_SG4_SMEM_KB=180 EVR_SG4_BCAP=3000
run EVR_SG4_V2_THREADS=512 EVR_SG4_SMEM_KB=160 EVR_SG4_BCAP=2600
run EVR_SG4_V2_THREADS=768 EVR_SG4_SMEM_KB=190 EVR_SG4_BCAP=3100
run EVR_SG4_V2_THREADS=768 EVR_SG4_SMEM_KB=160 EVR_SG4_BCAP=2600
run EVR_SG4_V2=0 EVR_SG4_SMEM_KB=190 EVR_SG4_BCAP=2350
run EVR_SG4_V2=0 EVR_SG4_BCAP=2350
cat $O/r2s7_sweep.txt
