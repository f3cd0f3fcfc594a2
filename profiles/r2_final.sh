#!/bin/bash
# round 2, final single-GPU evidence: smoke, bench (both arms), ncu launch list, DRAM traffic of one H|psi>, ncu --set full
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
python __graft_entry__.py smoke > $O/r2f_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2f_bench.json 2>$O/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/r2f_bench_ref.json 2>>$O/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r2f_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/r2f_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/r2f_traffic.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_fast -c 1 -o $O/r2f_ncu -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > $O/r2f_ncu.log 2>&1
cat $O/r2f_smoke.log | tail -2; cat $O/r2f_bench.json | cut -c1-1800; cat $O/r2f_bench_ref.json | cut -c1-600; tail -3 $O/r2f_bench.err
