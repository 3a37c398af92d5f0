#!/usr/bin/env python
"""PCIe ceiling vs the host pipeline at BASELINE config 2: raw pinned H2D / D2H / duplex bandwidth, then
HostPipeline.step over chunk sizes and slot counts.  Prints JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import weed_instance_segmentation_b200 as wis  # noqa: E402
from weed_instance_segmentation_b200.synth import msda_inputs  # noqa: E402

SHAPES = [(32, 32), (64, 64), (128, 128)]


def timed(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    nbytes = 341311488
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def h2d():
        d_a.copy_(h_a, non_blocking=True)

    def d2h():
        h_b.copy_(d_b, non_blocking=True)

    def duplex():
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    for name, fn in (("h2d", h2d), ("d2h", d2h), ("duplex", duplex)):
        ms = timed(fn)
        print(json.dumps({"copy": name, "ms": ms, "GBps_per_direction": nbytes / ms / 1e6}), flush=True)

    B = 8
    x = msda_inputs(B, SHAPES, dist="init", seed=0, value_dtype=torch.bfloat16)
    host_in = [x[k].contiguous().pin_memory() for k in ("value", "sampling_locations", "attention_weights", "grad_out")]
    for chunk, slots in ((1, 2), (1, 3), (1, 4), (2, 3), (4, 3), (8, 2)):
        with wis.HostPipeline(B, SHAPES, 8, 32, 4, chunk_images=chunk, slots=slots) as pipe:
            res = pipe.empty_outputs()

            def step():
                pipe.step(*host_in, **res)

            # per-step join (no overlap across steps) and free-running (join once at the end)
            def step_join():
                pipe.step(*host_in, **res)
                pipe.join()

            ms_join = timed(step_join)
            for _ in range(2):
                step()
            pipe.join()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            pipe.join()
            e1.record()
            torch.cuda.synchronize()
            ms_free = e0.elapsed_time(e1) / 10
        print(json.dumps({"chunk_images": chunk, "slots": slots, "ms_per_step_joined": ms_join,
                          "ms_per_step_free_running": ms_free, "images_per_s": B / ms_free * 1e3}), flush=True)


if __name__ == "__main__":
    main()
