#!/bin/bash
# Training-step throughput at BASELINE configs 3 and 4 on one GPU, ablating the loss path and the input path:
#   reference            stock HF model
#   b200-stock-loss      B200 MSDeformAttn path, stock criterion, reference input layout
#   b200/reference-input + batched criterion
#   b200                 + pinned uint8 masks prefetched on a side stream
mkdir -p gpurun_out
run() { # tag, extra args...
  tag=$1; shift
  python -m weed_instance_segmentation_b200.train --batch 16 --steps 6 --warmup 2 "$@" 2>/dev/null | tail -1 > gpurun_out/abl_c3_$tag.json
  python -m weed_instance_segmentation_b200.train --backbone swin_b --height 1024 --width 1024 --classes 5 --batch 8 --amp bf16 --steps 8 --warmup 2 "$@" 2>/dev/null | tail -1 > gpurun_out/abl_c4_$tag.json
}
[ "$1" = "with-reference" ] && run reference --impl reference
run b200-stock-loss --impl b200-stock-loss
run b200-reference-input --impl b200 --input reference
run b200 --impl b200
for f in gpurun_out/abl_c*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(d["value"],2), "img/s", round(d["ms_per_step"],1), "ms", "loss", round(d["loss"],3))
PY
done
