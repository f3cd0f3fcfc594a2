#!/bin/bash
# round 2, sweep 31: zero-fill inside the permute-in kernel + read-out through the inverse permutation (EVR_SG4_PERMUTE=1)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s31.txt
EVR_SG4_PERMUTE=1 timeout 600 python -m pytest tests -m gpu -x -q -k "benchmarked or henon or remainder or device_entry or scaled" 2>&1 | tail -2 >> $O/r2s31.txt
run() { echo "## $*" >> $O/r2s31.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 20 --warmup 3 2>>$O/r2s31_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s31.txt 2>&1; }
run EVR_X=0
run EVR_SG4_PERMUTE=1
run EVR_X=0
run EVR_SG4_PERMUTE=1
cat $O/r2s31.txt; tail -2 $O/r2s31_err.log
