#!/bin/bash
# round 2: all-reduce with in-kernel barriers (default) against the two separate barrier launches (EVR_SG4_ALLREDUCE=barriers)
# usage: r2_scale3.sh N
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
mkdir -p $O
F=$O/r2_scale3_N$N.txt
: > $F
run() { n=$1; shift; echo "## N=$n $*" >> $F;
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>>$O/r2_scale3_err.log | grep '^{' >> $F; }
run $N EVR_SG4_ALLREDUCE=fused
run $N EVR_X=0
run $N EVR_SG4_ALLREDUCE=fused
if [ "$N" = 2 ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "devices or peer or barriers or cpp_host" >> $F 2>&1; fi
python - <<PY
import json
for l in open("$F"):
    if l.startswith("##"): print(l.strip())
    elif l.startswith("{"):
        d = json.loads(l); print(d["n_gpus"], d["ms_per_step"], round(d["value"],1), d.get("kernel_ms"), d.get("allreduce_ms"), d.get("e2e",{}).get("value"), d.get("parity",{}).get("rel_l2"), d["config"].get("parallelism","")[-80:])
    else: print(l.strip()[:200])
PY
