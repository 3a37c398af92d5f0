#!/usr/bin/env python
"""Same-process A/B of the training micro-batch: one model, the criterion and the input path toggled between
measurements, repeated so that box-to-box and warm-up noise cancels.  Usage: train_ab.py [c3|c4] [rounds]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weed_instance_segmentation_b200 import synth, train  # noqa: E402
from weed_instance_segmentation_b200.criterion import convert_criterion, restore_criterion  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bb, B, Hh, Ww, C, amp = ("swin_b", 8, 1024, 1024, 5, True) if cfg == "c4" else ("swin_t", 16, 966, 1296, 3, False)
    dev = torch.device("cuda", 0)
    model = train.build_model(bb, C)
    train.use_b200_path(model, "modules", criterion=False)
    tr = train.Trainer(model, dev, amp_dtype=torch.bfloat16 if amp else None)
    ref_batches = [synth.collate_batch(B, Hh, Ww, C, seed=i) for i in range(2)]
    nat_batches = [synth.collate_batch(B, Hh, Ww, C, seed=i, mask_dtype=torch.uint8, pin_memory=True) for i in range(2)]
    variants = {
        "stock-loss/reference-input": (False, ref_batches, False),
        "b200-loss/reference-input": (True, ref_batches, False),
        "b200-loss/b200-input": (True, nat_batches, True),
    }
    res = {k: [] for k in variants}
    steps = 6
    for r in range(rounds):
        for name, (conv, batches, pre) in variants.items():
            (convert_criterion if conv else restore_criterion)(model)
            secs = train.throughput(tr, batches, steps, 2, prefetch=pre)
            res[name].append(secs / steps * 1e3)
    print(json.dumps({"config": cfg, "ms_per_micro_batch": res,
                      "median": {k: sorted(v)[len(v) // 2] for k, v in res.items()}}))


if __name__ == "__main__":
    main()
