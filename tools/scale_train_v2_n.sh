#!/bin/bash
# One scaling point of the full-path training step (BASELINE config 3) at N GPUs.  Usage: scale_train_v2_n.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  -m weed_instance_segmentation_b200.train --batch 16 --steps 8 --warmup 6 --impl b200 2>/dev/null | tail -1 > gpurun_out/v2_c3_n$N.json
python - gpurun_out/v2_c3_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); print(d["n_gpus"], round(d["value"],2), "img/s", round(d["ms_per_step"],1), "ms")
PY
