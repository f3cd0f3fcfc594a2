#!/bin/bash
# round 2, final single-GPU confirmation of the last build: full GPU suite, smoke, bench (both arms), launch list
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2g_suite.log 2>&1; echo "suite rc=$?" >> $O/r2g_suite.log
python __graft_entry__.py smoke > $O/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2g_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2g_bench.json 2>$O/r2g_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/r2g_bench_ref.json 2>>$O/r2g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r2g_launches.log 2>&1
tail -n 3 $O/r2g_suite.log; tail -n 2 $O/r2g_smoke.log; cut -c1-2200 $O/r2g_bench.json; cut -c1-500 $O/r2g_bench_ref.json; tail -n 2 $O/r2g_bench.err
