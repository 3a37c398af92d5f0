#!/bin/bash
# Training-step scaling point (BASELINE config 3) at N GPUs plus the op bench at the same N.  Usage: scale_train_n.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 -m weed_instance_segmentation_b200.train --batch 16 --steps 6 --warmup 2 --impl b200 > gpurun_out/train_n$N.log 2>&1
tail -1 gpurun_out/train_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.log 2>/dev/null
tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3))"
