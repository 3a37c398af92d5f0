#!/usr/bin/env python
"""Where a training step spends its time (BASELINE config 3 or 4 geometry, one GPU): torch profiler over a few
micro-batches, grouped into the model's top-level parts by forward hooks (CUDA events), plus the top kernels.

    python tools/profile_train.py [--backbone swin_t] [--batch 16] [--height 966] [--width 1296] [--amp]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weed_instance_segmentation_b200 import synth, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="swin_t")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=966)
    ap.add_argument("--width", type=int, default=1296)
    ap.add_argument("--labels", type=int, default=3)
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = train.build_model(args.backbone, args.labels).to(dev).train()
    if args.impl == "b200":
        train.use_b200_path(model, "modules")
    parts = {
        "backbone": model.model.pixel_level_module.encoder,
        "pixel_decoder": model.model.pixel_level_module.decoder,
        "pixel_decoder.encoder(6 MSDA layers)": model.model.pixel_level_module.decoder.encoder,
        "transformer_decoder": model.model.transformer_module,
        "criterion(loss+matcher)": model.criterion,
    }
    spans = {k: [] for k in parts}

    def hook_pair(name):
        def pre(mod, inp):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            spans[name].append([e, None])

        def post(mod, inp, out):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            spans[name][-1][1] = e
        return pre, post

    for name, mod in parts.items():
        pre, post = hook_pair(name)
        mod.register_forward_pre_hook(pre)
        mod.register_forward_hook(post)

    raw = [synth.collate_batch(args.batch, args.height, args.width, args.labels, seed=i) for i in range(2)]
    batches = [dict(pixel_values=b["pixel_values"].to(dev), mask_labels=[m.to(dev) for m in b["mask_labels"]],
                    class_labels=[c.to(dev) for c in b["class_labels"]]) for b in raw]
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if args.amp else torch.autocast("cuda", enabled=False)

    def micro(b):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        with ctx:
            out = model(**b)
        e1.record()
        out.loss.backward()
        e2.record()
        return e0, e1, e2

    for b in batches:  # warm-up
        micro(b)
    model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    for v in spans.values():
        v.clear()
    marks = [micro(b) for b in batches]
    torch.cuda.synchronize()
    fwd = sum(a.elapsed_time(b) for a, b, _ in marks) / len(marks)
    bwd = sum(b.elapsed_time(c) for _, b, c in marks) / len(marks)
    split = {k: sum(a.elapsed_time(b) for a, b in v) / len(marks) for k, v in spans.items()}
    print(json.dumps({"micro_batch": args.batch, "forward_ms": fwd, "backward_ms": bwd, "forward_split_ms": split}), flush=True)

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        micro(batches[0])
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        if ev.key.startswith("b200_loss::"):
            print(json.dumps({"region": ev.key, "cpu_ms": ev.cpu_time_total / 1e3, "cuda_ms": ev.device_time_total / 1e3,
                              "calls": ev.count}))
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=70))


if __name__ == "__main__":
    main()
