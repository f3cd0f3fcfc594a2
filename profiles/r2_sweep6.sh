#!/bin/bash
# round 2, sixth GPU session: v2 kernel, at most one 15-value tile per term; 768 x 80 registers vs 512 x 128 registers
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s6_tests.log 2>&1; echo "tests rc=$?" > $O/r2s6_sweep.txt
tail -15 $O/r2s6_tests.log >> $O/r2s6_sweep.txt
run() { echo "## $*" >> $O/r2s6_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s6_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s6_sweep.txt 2>&1; }
run EVR_SG4_V2=1
run EVR_SG4_MAX35=1000
run EVR_SG4_V2_THREADS=512
run EVR_SG4_V2_THREADS=512 EVR_SG4_MAX35=1000
run EVR_SG4_V2_THREADS=512 EVR_SG4_BCAP=5600
run EVR_SG4_V2_THREADS=512 EVR_SG4_G0=128 EVR_SG4_BCAP=2700
run EVR_SG4_V2_THREADS=512 EVR_SG4_G0=512 EVR_SG4_BCAP=11000
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_v2 -c 1 -o $O/r2s6_ncu -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s6_ncu.log 2>&1
EVR_SG4_V2_THREADS=512 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_v2 -c 1 -o $O/r2s6_ncu512 -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s6_ncu512.log 2>&1
cat $O/r2s6_sweep.txt
