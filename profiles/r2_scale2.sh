#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2_scale2.txt
run() { n=$1; shift; echo "## N=$n $*" >> $O/r2_scale2.txt;
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-e2e 2>>$O/r2_scale2_err.log | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['allreduce_ms'])" >> $O/r2_scale2.txt; }
run 8 EVR_X=0
run 8 EVR_SG4_ITEMS_PER_SM=8
run 8 EVR_SG4_ITEMS_PER_SM=32
run 8 EVR_SG4_ITEMS_PER_SM=64
run 8 EVR_SG4_ITEMS_PER_SM=8 EVR_SG4_BCAP=4700
cat $O/r2_scale2.txt
