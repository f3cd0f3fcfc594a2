#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2s16.txt
run() { echo "## $*" >> $O/r2s16.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s16_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s16.txt 2>&1; }
L5=$PWD/elvibrot-tnumtana_b200/libevr_sg4_t512.so
L6=$PWD/elvibrot-tnumtana_b200/libevr_sg4_t640.so
run EVR_SG4_LIB=$L5 EVR_SG4_BCAP=3500
run EVR_SG4_LIB=$L5 EVR_SG4_BCAP=2300
run EVR_SG4_LIB=$L5 EVR_SG4_BCAP=2800
run EVR_SG4_LIB=$L5 EVR_SG4_BCAP=3500 EVR_SG4_G0=256 EVR_SG4_TH0=3600
run EVR_SG4_LIB=$L6 EVR_SG4_BCAP=2800
run EVR_SG4_LIB=$L6 EVR_SG4_BCAP=2600
run EVR_SG4_LIB=$L6 EVR_SG4_BCAP=2800 EVR_SG4_ITEMS_PER_SM=64
echo "## t640 L=6 / npsi=8" >> $O/r2s16.txt
EVR_SG4_LIB=$L6 EVR_SG4_BCAP=2800 timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --L 6 2>>$O/r2s16_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s16.txt 2>&1
EVR_SG4_LIB=$L6 EVR_SG4_BCAP=2800 timeout 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --npsi 8 2>>$O/r2s16_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s16.txt 2>&1
cat $O/r2s16.txt; tail -3 $O/r2s16_err.log
