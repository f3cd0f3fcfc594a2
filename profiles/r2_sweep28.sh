#!/bin/bash
# round 2, sweep 28: L2 prefetch of the next item's slices on the final kernel (0 = off, 1 = per-lane prefetch.global.L2,
# 2 = one cp.async.bulk.prefetch.L2 per slice)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s28.txt
run() { echo "## $*" >> $O/r2s28.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 20 --warmup 3 2>>$O/r2s28_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s28.txt 2>&1; }
run EVR_X=0
run EVR_SG4_PREFETCH=2
run EVR_SG4_PREFETCH=1
run EVR_X=0
run EVR_SG4_PREFETCH=2
run EVR_SG4_PREFETCH=2 EVR_SG4_DEBUG=60
cat $O/r2s28.txt; tail -3 $O/r2s28_err.log
