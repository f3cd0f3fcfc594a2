#!/bin/bash
# round 2, sweep 30: what the 0.134 ms "skeleton" of the term kernel (no passes, no RED, no psi gather, no V) consists of
#   60 = 4+8+16+32; +8192 no scatter phase (map staging + loop); +16384 no gather phase (map staging + loop); 4096 = item loop
#   and descriptor copies only
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s30.txt
run() { echo "## $*" >> $O/r2s30.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 20 --warmup 3 2>>$O/r2s30_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s30.txt 2>&1; }
run EVR_SG4_DEBUG=60
run EVR_SG4_DEBUG=8252
run EVR_SG4_DEBUG=16444
run EVR_SG4_DEBUG=24636
run EVR_SG4_DEBUG=4096
cat $O/r2s30.txt; tail -2 $O/r2s30_err.log
