#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck (shared-memory hazards between the thread groups and the
# bulk-copy / cp.async staging) and synccheck (named barriers) on small cases of every kernel family.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for fam in fast generic nested algebra; do
    echo "=== $tool / $fam" >> $O/sanitize_r2.txt
    timeout 900 $S --tool $tool --print-limit 5 python profiles/sanitize_cases.py $fam 2>&1 | grep -v "^$" | tail -12 >> $O/sanitize_r2.txt
  done
done
echo "=== memcheck / fast, second-generation kernel (EVR_SG4_V2=1)" >> $O/sanitize_r2.txt
EVR_SG4_V2=1 timeout 900 $S --tool memcheck --print-limit 5 python profiles/sanitize_cases.py fast 2>&1 | grep -v "^$" | tail -8 >> $O/sanitize_r2.txt
echo "=== racecheck / fast, second-generation kernel (EVR_SG4_V2=1)" >> $O/sanitize_r2.txt
EVR_SG4_V2=1 timeout 900 $S --tool racecheck --print-limit 5 python profiles/sanitize_cases.py fast 2>&1 | grep -v "^$" | tail -8 >> $O/sanitize_r2.txt
cat $O/sanitize_r2.txt
