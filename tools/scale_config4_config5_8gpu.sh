#!/bin/bash
# BASELINE configs 4 and 5 on one 8-GPU box (N=8); the N=1 legs run in a separate 1-GPU call.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 -m weed_instance_segmentation_b200.train --backbone swin_b --height 1024 --width 1024 --classes 5 --batch 8 --amp bf16 --steps 6 --warmup 2 > gpurun_out/c4_n8.log 2>&1
tail -1 gpurun_out/c4_n8.log
timeout 600 $TR --master-port 29522 -m weed_instance_segmentation_b200.train --infer --height 2048 --width 2048 --batch 4 --steps 6 --warmup 2 > gpurun_out/c5_n8.log 2>&1
tail -1 gpurun_out/c5_n8.log
echo done
