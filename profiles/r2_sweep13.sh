#!/bin/bash
# round 2: v1 kernel with the scatter fused into the last G->B pass and the hi-major lane mapping for small strides
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2s13.txt
run() { echo "## $*" >> $O/r2s13.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s13_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s13.txt 2>&1; }
run EVR_SG4_DEBUG=0
run EVR_SG4_DEBUG=256
run EVR_SG4_DEBUG=512
run EVR_SG4_DEBUG=768
run EVR_SG4_BCAP=4700
run EVR_SG4_BCAP=3100
run EVR_SG4_BCAP=1800
echo "## L=6" >> $O/r2s13.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --L 6 2>>$O/r2s13_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s13.txt 2>&1
echo "## npsi=8" >> $O/r2s13.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --npsi 8 2>>$O/r2s13_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" >> $O/r2s13.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_fast -c 1 -o $O/r2s13_ncu -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s13_ncu.log 2>&1
cat $O/r2s13.txt
