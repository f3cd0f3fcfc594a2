#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s22.txt
run() { echo "## $*" >> $O/r2s22.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s22_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s22.txt 2>&1; }
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=2200
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=2800
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=2200 EVR_SG4_G1=128
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=2200 EVR_SG4_G1=64
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=2200 EVR_SG4_MAX35=1000
run EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=1800 EVR_SG4_G1=64
cat $O/r2s22.txt; tail -3 $O/r2s22_err.log
