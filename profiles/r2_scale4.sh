#!/bin/bash
# round 2: grid size of the all-reduce kernel (CTAs per SM) at N GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
mkdir -p $O
F=$O/r2_scale4_N$N.txt
: > $F
run() { n=$1; shift; echo "## N=$n $*" >> $F;
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-e2e 2>>$O/r2_scale4_err.log | grep '^{' >> $F; }
run $N EVR_SG4_ALLREDUCE=barriers
run $N EVR_SG4_ALLREDUCE=barriers EVR_SG4_AR_DEBUG=1
run $N EVR_SG4_ALLREDUCE=barriers EVR_SG4_AR_DEBUG=2
run $N EVR_SG4_ALLREDUCE=barriers EVR_SG4_AR_DEBUG=3
run $N EVR_SG4_ALLREDUCE=nccl
python - <<PY
import json
for l in open("$F"):
    if l.startswith("##"): print(l.strip())
    elif l.startswith("{"):
        d = json.loads(l); print(d["n_gpus"], d["ms_per_step"], round(d["value"],1), d["roofline"].get("kernel_ms"), d.get("allreduce_ms"))
PY
