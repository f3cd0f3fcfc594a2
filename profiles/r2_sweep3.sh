#!/bin/bash
# round 2, third GPU session: TMA bulk staging of the map / V slices; phase isolation; one ncu --set full capture
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s3_tests.log 2>&1; echo "tests rc=$?" > $O/r2s3_sweep.txt
run() { echo "## $*" >> $O/r2s3_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s3_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s3_sweep.txt 2>&1; }
run EVR_SG4_DEBUG=0
run EVR_SG4_DEBUG=60
run EVR_SG4_DEBUG=56
run EVR_SG4_DEBUG=4
run EVR_SG4_DEBUG=128
run EVR_SG4_ISO=2
run EVR_SG4_ISO=2 EVR_SG4_BCAP=7000
run EVR_SG4_BCAP=2350
run EVR_SG4_BATCH=0
EVR_SG4_DEBUG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_fast -c 1 -o $O/r2s3_ncu_dbg0 -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s3_ncu_dbg0.log 2>&1
EVR_SG4_DEBUG=60 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_fast -c 1 -o $O/r2s3_ncu_dbg60 -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s3_ncu_dbg60.log 2>&1
cat $O/r2s3_sweep.txt; tail -3 $O/r2s3_tests.log; ls -la $O
