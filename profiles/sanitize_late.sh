#!/bin/bash
# compute-sanitizer on the kernels added late in round 2 (profiles/sanitize_cases.py late)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
S=/usr/local/cuda/bin/compute-sanitizer
: > $O/sanitize_r2_late.txt
for tool in memcheck racecheck synccheck; do
  echo "=== $tool / late" >> $O/sanitize_r2_late.txt
  timeout 1200 $S --tool $tool --print-limit 5 python profiles/sanitize_cases.py late 2>&1 | grep -v "^$" | tail -12 >> $O/sanitize_r2_late.txt
done
cat $O/sanitize_r2_late.txt
