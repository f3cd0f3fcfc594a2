#!/bin/bash
# usage: sweep.sh "ENV1=.. ENV2=.." ... ; prints ms_per_step for each environment setting
for cfg in "$@"; do
  r=$(env $cfg python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "$cfg => $r"
done
