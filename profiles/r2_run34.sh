#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s34_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/r2s34_suite.log
python __graft_entry__.py smoke > gpurun_out/r2s34_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2s34_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2s34_bench.json 2> gpurun_out/r2s34_bench.err
tail -n 3 gpurun_out/r2s34_suite.log; tail -n 2 gpurun_out/r2s34_smoke.log; python -c "
import json; d=json.load(open('gpurun_out/r2s34_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'], d['clocks'], d['gpu_launches'])"
