#!/bin/bash
# round 2: CUDA graph of the per-call launch sequence -- tests, L=7 bench, small-shape bench with and without the graph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2s12.txt
for g in 1 0; do
  echo "## EVR_SG4_GRAPH=$g  L=7" >> $O/r2s12.txt
  EVR_SG4_GRAPH=$g timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 2>>$O/r2s12_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])" >> $O/r2s12.txt 2>&1
  echo "## EVR_SG4_GRAPH=$g  shapes" >> $O/r2s12.txt
  EVR_SG4_GRAPH=$g timeout 600 python profiles/shape_bench.py >> $O/r2s12.txt 2>>$O/r2s12_err.log
done
cat $O/r2s12.txt; tail -5 $O/r2s12_err.log
