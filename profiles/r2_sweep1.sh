#!/bin/bash
# round 2, first GPU session: batched work items ("super-terms") -- tests, then a sweep of the batch capacity
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s1_tests.log 2>&1; echo "tests rc=$?" > $O/r2s1_sweep.txt
run() { echo "## $*" >> $O/r2s1_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s1_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s1_sweep.txt 2>&1; }
run EVR_SG4_BATCH=0
run EVR_SG4_BCAP=4700
run EVR_SG4_BCAP=2350
run EVR_SG4_BCAP=7000 EVR_SG4_G0=384
run EVR_SG4_BCAP=4700 EVR_SG4_G0=128
run EVR_SG4_BCAP=3500 EVR_SG4_G0=192
run EVR_SG4_BCAP=2350 EVR_SG4_G1=64
run EVR_SG4_BCAP=1400 EVR_SG4_G2=32 EVR_SG4_TH2=1500
run EVR_SG4_BCAP=4700 EVR_SG4_DEBUG=4
run EVR_SG4_BCAP=4700 EVR_SG4_DEBUG=60
run EVR_SG4_BCAP=4700 EVR_SG4_BLOCK_ORDER=0
echo "## L=6" >> $O/r2s1_sweep.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --L 6 2>>$O/r2s1_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s1_sweep.txt 2>&1
echo "## npsi=8" >> $O/r2s1_sweep.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --npsi 8 2>>$O/r2s1_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s1_sweep.txt 2>&1
cat $O/r2s1_sweep.txt; tail -5 $O/r2s1_tests.log
