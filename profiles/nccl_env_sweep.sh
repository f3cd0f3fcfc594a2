#!/bin/bash
# usage: nccl_env_sweep.sh N "ENV=.. ENV=.." ... ; prints ms per H|psi>, slowest rank's kernels and the all-reduce alone at N GPUs
N=$1; shift
port=29600
for cfg in "$@"; do
  port=$((port+1))
  r=$(env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --no-cpu --no-e2e --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['allreduce_ms'])" 2>&1 | tail -1)
  echo "N=$N $cfg => $r"
done
