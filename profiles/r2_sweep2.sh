#!/bin/bash
# round 2, second GPU session: L1TEX cost micro-benchmark, phase isolation with the debug bits, ncu --set full captures
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
./profiles/micro/l1tex_costs > $O/r2s2_micro.txt 2>&1
run() { echo "## $*" >> $O/r2s2_sweep.txt; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>>$O/r2s2_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $O/r2s2_sweep.txt 2>&1; }
: > $O/r2s2_sweep.txt
run EVR_SG4_DEBUG=0
run EVR_SG4_DEBUG=12
run EVR_SG4_DEBUG=20
run EVR_SG4_DEBUG=36
run EVR_SG4_DEBUG=56
run EVR_SG4_DEBUG=8
run EVR_SG4_DEBUG=16
run EVR_SG4_DEBUG=24
for dbg in 0 4 56; do
EVR_SG4_DEBUG=$dbg timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_fast -c 3 -o $O/r2s2_ncu_dbg$dbg -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2s2_ncu_dbg$dbg.log 2>&1
done
cat $O/r2s2_micro.txt $O/r2s2_sweep.txt; ls -la $O
