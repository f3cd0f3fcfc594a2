#!/bin/bash
# round 2, fourth GPU session: second-generation kernel (sg4_fast2.cuh) -- tests, bench, phase isolation, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s4_tests.log 2>&1; echo "tests rc=$?" > $O/r2s4_sweep.txt
tail -15 $O/r2s4_tests.log >> $O/r2s4_sweep.txt
run() { echo "## $*" >> $O/r2s4_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s4_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s4_sweep.txt 2>&1; }
run EVR_SG4_V2=1
run EVR_SG4_V2=0
run EVR_SG4_BCAP=2350
run EVR_SG4_BCAP=7000 EVR_SG4_G0=384
run EVR_SG4_DEBUG=64
run EVR_SG4_DEBUG=8
run EVR_SG4_DEBUG=16
run EVR_SG4_DEBUG=32
run EVR_SG4_DEBUG=4
run EVR_SG4_DEBUG=60
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_v2 -c 1 -o $O/r2s4_ncu -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s4_ncu.log 2>&1
cat $O/r2s4_sweep.txt
