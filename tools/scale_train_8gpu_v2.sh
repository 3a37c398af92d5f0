#!/bin/bash
# Training step at N=8 (and N=1 on the same box) for BASELINE configs 3 and 4, full B200 path
# (MSDeformAttn modules + batched criterion + pinned uint8 masks, prefetched), 6 warm-up micro-batches.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
C3="--batch 16 --steps 8 --warmup 6 --impl b200"
C4="--backbone swin_b --height 1024 --width 1024 --classes 5 --batch 8 --amp bf16 --steps 10 --warmup 6 --impl b200"
timeout 600 python -m weed_instance_segmentation_b200.train $C3 2>/dev/null | tail -1 > gpurun_out/v2_c3_n1.json
timeout 600 python -m weed_instance_segmentation_b200.train $C4 2>/dev/null | tail -1 > gpurun_out/v2_c4_n1.json
timeout 900 $TR --master-port 29531 -m weed_instance_segmentation_b200.train $C3 2>/dev/null | tail -1 > gpurun_out/v2_c3_n8.json
timeout 900 $TR --master-port 29532 -m weed_instance_segmentation_b200.train $C4 2>/dev/null | tail -1 > gpurun_out/v2_c4_n8.json
for f in gpurun_out/v2_c*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], d["n_gpus"], round(d["value"],2), "img/s", round(d["ms_per_step"],1), "ms")
PY
done
