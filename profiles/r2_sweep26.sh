#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2s26.txt
run() { echo "## $*" >> $O/r2s26.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s26_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s26.txt 2>&1; }
run EVR_X=0
run EVR_X=1
run EVR_SG4_G1=128
run EVR_SG4_G1=64
cat $O/r2s26.txt; tail -3 $O/r2s26_err.log
