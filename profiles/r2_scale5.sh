#!/bin/bash
# round 2: per-rank kernel times under the points-balanced and the cost-balanced contiguous term ranges
# usage: r2_scale5.sh N
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
mkdir -p $O
F=$O/r2_scale5_N$N.txt
: > $F
run() { n=$1; shift; echo "## N=$n $*" >> $F;
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-e2e "$@" 2>>$O/r2_scale5_err.log | grep '^{' >> $F; }
run $N --partition points
run $N --partition cost
run $N --partition count
python - <<PY
import json
for l in open("$F"):
    if l.startswith("##"): print(l.strip())
    elif l.startswith("{"):
        d = json.loads(l); print(d["n_gpus"], round(d["ms_per_step"],4), round(d["value"],1), d["roofline"].get("kernel_ms"), d.get("allreduce_ms"), d.get("kernel_ms_per_rank"))
PY
