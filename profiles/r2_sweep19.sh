#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2s19.txt
run() { echo "## $*" >> $O/r2s19.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s19_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s19.txt 2>&1; }
run EVR_X=0
run EVR_SG4_DYNAMIC=0
run EVR_SG4_G1=64
run EVR_SG4_G1=128
run EVR_SG4_BCAP=3000
run EVR_SG4_BCAP=2500
run EVR_SG4_BCAP=2200 EVR_SG4_G1=64
run EVR_SG4_BCAP=1800 EVR_SG4_G1=64
run EVR_SG4_BCAP=3400 EVR_SG4_TH0=3600
echo "## L=6 / npsi=8" >> $O/r2s19.txt
for a in "--L 6" "--npsi 8 --steps 5"; do timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 $a 2>>$O/r2s19_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s19.txt 2>&1; done
cat $O/r2s19.txt; tail -3 $O/r2s19_err.log
