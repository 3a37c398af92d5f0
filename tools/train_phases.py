#!/usr/bin/env python
"""Wall-clock phases of one training micro-batch (synchronised between phases): host->device copy of the batch,
forward (+loss), backward, optimizer.  Usage: train_phases.py [c3|c4] [impl]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weed_instance_segmentation_b200 import synth, train  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
    impl = sys.argv[2] if len(sys.argv) > 2 else "b200"
    bb, B, Hh, Ww, C, amp = ("swin_b", 8, 1024, 1024, 5, True) if cfg == "c4" else ("swin_t", 16, 966, 1296, 3, False)
    dev = torch.device("cuda", 0)
    model = train.build_model(bb, C).to(dev).train()
    if impl != "reference":
        train.use_b200_path(model, "modules", criterion=impl == "b200")
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5)
    batches = [synth.collate_batch(B, Hh, Ww, C, seed=i) for i in range(2)]
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if amp else torch.autocast("cuda", enabled=False)
    acc = {"h2d": 0.0, "forward": 0.0, "backward": 0.0, "optimizer": 0.0}
    n = 0
    for it in range(6):
        b = batches[it % 2]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pv = b["pixel_values"].to(dev, non_blocking=True)
        ml = [m.to(dev, non_blocking=True) for m in b["mask_labels"]]
        cl = [c.to(dev, non_blocking=True) for c in b["class_labels"]]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with ctx:
            out = model(pixel_values=pv, mask_labels=ml, class_labels=cl)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        (out.loss / 2).backward()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it % 2 == 1:
            opt.step()
            opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if it >= 2:
            n += 1
            for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                acc[k] += v * 1e3
    mb = sum(m.numel() * m.element_size() for m in batches[0]["mask_labels"]) / 1e6
    print(json.dumps({"config": cfg, "impl": impl, "ms_per_micro_batch": {k: v / n for k, v in acc.items()},
                      "mask_labels_MB": mb, "pixel_values_MB": batches[0]["pixel_values"].numel() * 4 / 1e6}))


if __name__ == "__main__":
    main()
