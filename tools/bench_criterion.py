#!/usr/bin/env python
"""Loss / matcher path alone: reference Mask2FormerLoss vs the batched B200 criterion, forward + backward, at the
loss geometry of BASELINE config 4 (10 layers, batch 8, 100 queries, 256x256 logits, 1024x1024 targets)."""
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_criterion_host import make_problem  # noqa: E402
from weed_instance_segmentation_b200 import _cabi  # noqa: E402
from weed_instance_segmentation_b200.criterion import convert_criterion  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_tgt = tuple((3, 9, 14, 1, 20, 7, 5, 11, 2, 16, 8, 4, 12, 6, 10, 13)[:B])
    loss, masks, classes, mask_labels, class_labels = make_problem(
        1, B=B, Q=100, C=5, L=10, h=256, w=256, H=1024, W=1024, n_tgt=n_tgt, num_points=12544)
    loss = loss.cuda()
    masks, classes = [m.cuda() for m in masks], [c.cuda() for c in classes]
    mask_labels, class_labels = [m.cuda() for m in mask_labels], [c.cuda() for c in class_labels]
    res = {}
    for name, crit in (("reference", loss), ("b200", convert_criterion(copy.deepcopy(loss)))):
        def step():
            ms = [m.clone().requires_grad_(True) for m in masks]
            cs = [c.clone().requires_grad_(True) for c in classes]
            aux = [{"masks_queries_logits": m, "class_queries_logits": c} for m, c in zip(ms[:-1], cs[:-1])]
            out = crit(ms[-1], cs[-1], mask_labels, class_labels, aux)
            sum(out.values()).backward()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _cabi.load().msda_b200_launch_count(1)
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        res[name] = {"ms_fwd_bwd": (time.perf_counter() - t0) / n * 1e3,
                     "own_kernel_launches_per_step": _cabi.load().msda_b200_launch_count(0) / n}
    res["speedup"] = res["reference"]["ms_fwd_bwd"] / res["b200"]["ms_fwd_bwd"]
    print(json.dumps({"metric": "mask2former_criterion_fwd_bwd_ms", "batch": B, "layers": 10, "results": res}))


if __name__ == "__main__":
    main()
