#!/usr/bin/env python
"""torch profiler over the batched criterion alone (config-4 loss geometry): top device ops and host regions."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_criterion_host import make_problem  # noqa: E402
from weed_instance_segmentation_b200.criterion import convert_criterion  # noqa: E402


def main():
    loss, masks, classes, mask_labels, class_labels = make_problem(
        1, B=8, Q=100, C=5, L=10, h=256, w=256, H=1024, W=1024, n_tgt=(3, 9, 14, 1, 20, 7, 5, 11), num_points=12544)
    crit = convert_criterion(loss.cuda())
    masks, classes = [m.cuda() for m in masks], [c.cuda() for c in classes]
    mask_labels, class_labels = [m.to(torch.uint8).cuda() for m in mask_labels], [c.cuda() for c in class_labels]

    def step():
        ms = [m.clone().requires_grad_(True) for m in masks]
        cs = [c.clone().requires_grad_(True) for c in classes]
        aux = [{"masks_queries_logits": m, "class_queries_logits": c} for m, c in zip(ms[:-1], cs[:-1])]
        out = crit(ms[-1], cs[-1], mask_labels, class_labels, aux)
        sum(out.values()).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))


if __name__ == "__main__":
    main()
