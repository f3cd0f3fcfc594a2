#!/bin/bash
# round 2: scaling on one 8 x B200 box -- bench.py at N = 1, 2, 4, 8 (torchrun, as the driver launches it), block order on/off at
# N = 8, and the single-process multi-GPU C-ABI path (evr_sg4_set_devices) end to end
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l > $O/r2_scale.txt
run() { n=$1; shift; echo "## N=$n $*" >> $O/r2_scale.txt;
  if [ $n -eq 1 ]; then env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 >> $O/r2_scale.txt 2>>$O/r2_scale_err.log;
  else env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 2>>$O/r2_scale_err.log | grep '^{' >> $O/r2_scale.txt; fi; }
run 1 EVR_X=0
run 2 EVR_X=0
run 4 EVR_X=0
run 8 EVR_X=0
run 8 EVR_SG4_BLOCK_ORDER=0
run 4 EVR_SG4_BLOCK_ORDER=0
echo "## set_devices e2e (one process, C-ABI)" >> $O/r2_scale.txt
timeout 900 python profiles/multi_e2e.py 1 2 4 8 >> $O/r2_scale.txt 2>>$O/r2_scale_err.log
python - <<'PY' >> gpurun_out/r2_scale.txt
import json
print("## summary: N, ms_per_step, Hpsi/s, kernel_ms, allreduce_ms, e2e Hpsi/s, parity rel_l2")
for ln in open('gpurun_out/r2_scale.txt'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(d['n_gpus'], round(d['ms_per_step'], 4), round(d['value'], 1), round(d['roofline']['kernel_ms'], 4), d.get('allreduce_ms'), round(d['e2e']['value'], 1), d['parity']['rel_l2'])
PY
tail -20 $O/r2_scale.txt; tail -5 $O/r2_scale_err.log
