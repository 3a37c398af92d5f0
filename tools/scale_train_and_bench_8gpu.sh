#!/bin/bash
# Scaling runs on one 8-GPU box: training step (BASELINE config 3) at N=1 and N=8, op bench at N=8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_smi.log
timeout 600 python -m weed_instance_segmentation_b200.train --batch 16 --steps 6 --warmup 2 --impl b200 > gpurun_out/train_n1.log 2>&1
tail -1 gpurun_out/train_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 -m weed_instance_segmentation_b200.train --batch 16 --steps 6 --warmup 2 --impl b200 > gpurun_out/train_n8.log 2>&1
tail -1 gpurun_out/train_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.log 2>&1
tail -c 400 gpurun_out/bench_n8.log
echo done
