#!/bin/bash
# round 2, sweep 29: early staging of the next item's gather map (default) vs EVR_SG4_DEBUG=2048 (off)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s29.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O/r2s29.txt
run() { echo "## $*" >> $O/r2s29.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 20 --warmup 3 2>>$O/r2s29_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s29.txt 2>&1; }
run EVR_X=0
run EVR_SG4_DEBUG=2048
run EVR_X=0
run EVR_SG4_DEBUG=2048
run EVR_SG4_DEBUG=60
cat $O/r2s29.txt; tail -3 $O/r2s29_err.log
