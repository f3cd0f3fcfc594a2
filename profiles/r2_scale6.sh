#!/bin/bash
# round 2: default bench line at N GPUs under torchrun with the last build (both arms as the driver launches them)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
mkdir -p $O
F=$O/r2_scale6_N$N.txt
: > $F
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 2>>$O/r2_scale6_err.log | grep '^{' >> $F
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>>$O/r2_scale6_err.log | grep '^{' >> $F
python - <<PY
import json
for l in open("$F"):
    d = json.loads(l)
    if d.get("impl") == "reference": print("reference arm:", round(d["value"],2), d["cpu_baseline"].get("cores")); continue
    print(d["n_gpus"], round(d["ms_per_step"],4), round(d["value"],1), d.get("kernel_ms_per_rank"), d.get("allreduce_ms"), (d.get("e2e") or {}).get("value"), d.get("parity"), d.get("clocks"))
PY
