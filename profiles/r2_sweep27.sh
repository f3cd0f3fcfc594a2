#!/bin/bash
# round 2, sweep 27: phase isolation on the FINAL kernel (results are wrong by construction: timing only)
#   4 = no transform passes, 8 = no RED, 16 = no psi gather, 32 = no V staging
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s27.txt
run() { echo "## $*" >> $O/r2s27.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s27_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s27.txt 2>&1; }
run EVR_X=0
run EVR_SG4_DEBUG=4
run EVR_SG4_DEBUG=12
run EVR_SG4_DEBUG=20
run EVR_SG4_DEBUG=28
run EVR_SG4_DEBUG=60
run EVR_SG4_DEBUG=8
run EVR_SG4_DEBUG=16
cat $O/r2s27.txt; tail -3 $O/r2s27_err.log
