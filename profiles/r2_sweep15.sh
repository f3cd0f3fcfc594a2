#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s15.txt
run() { echo "## $*" >> $O/r2s15.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s15_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s15.txt 2>&1; }
L=$PWD/elvibrot-tnumtana_b200/libevr_sg4_t640.so
run EVR_X=0
run EVR_SG4_LIB=$L
run EVR_SG4_LIB=$L EVR_SG4_BCAP=2800
run EVR_SG4_LIB=$L EVR_SG4_BCAP=1900
run EVR_SG4_LIB=$L EVR_SG4_BCAP=2800 EVR_SG4_G1=160 
run EVR_SG4_LIB=$L EVR_SG4_BCAP=4400 EVR_SG4_G0=320
cat $O/r2s15.txt; tail -3 $O/r2s15_err.log
