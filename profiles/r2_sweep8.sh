#!/bin/bash
# round 2, 2-GPU session: multi-device C-ABI (evr_sg4_set_devices), slice-wise host I/O, lane-consecutive gather in v1
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2s8_sweep.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2s8_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s8_sweep.txt
tail -15 $O/r2s8_tests.log >> $O/r2s8_sweep.txt
echo "## N=1 default" >> $O/r2s8_sweep.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2s8_bench1.json 2>>$O/r2s8_err.log; cat $O/r2s8_bench1.json >> $O/r2s8_sweep.txt
echo "## N=2 torchrun" >> $O/r2s8_sweep.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2s8_bench2.json 2>>$O/r2s8_err.log; cat $O/r2s8_bench2.json >> $O/r2s8_sweep.txt
echo "## set_devices e2e (one process)" >> $O/r2s8_sweep.txt
timeout 600 python profiles/multi_e2e.py 1 2 >> $O/r2s8_sweep.txt 2>>$O/r2s8_err.log
cat $O/r2s8_sweep.txt; tail -20 $O/r2s8_err.log
