#!/usr/bin/env python
"""Encoder-layer benchmark: one pixel-decoder encoder layer, forward + backward,
at BASELINE config 2 (B=8, S=21504, d_model=256, 8 heads, 3 levels, 4 points) on one B200.

Compares, with identical weights and inputs:
  reference : stock HF layer (M2F:986-1072) with the stock op (M2F:798-837, grid_sample)
  function  : stock HF layer, op rebound to the B200 operator (hf_patch.install)
  modules   : modules.EncoderLayer + MSDeformAttn, un-fused prologue
  fused     : modules.EncoderLayer + MSDeformAttn with the fused softmax/location prologue
  fused+norm: the same plus the fused residual + LayerNorm kernels (layer_norm.py)
  fused+norm+linear: the same plus projections with the fused bias-gradient reduction (linear.py), fc1 with bias + ReLU in
              the GEMM epilogue, and the bfloat16 copy of the first LayerNorm's output feeding fc1 (no cast kernel)
  full      : the same plus reference points computed inside the kernels (no reference-point tensor)

    python bench_layer.py [--amp bf16|none] [--steps K] [--warmup W] [--batch B]
Prints one JSON line.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPES = [(32, 32), (64, 64), (128, 128)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--amp", choices=["bf16", "none"], default="bf16")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()

    import torch
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former import modeling_mask2former as m2f

    import weed_instance_segmentation_b200 as wis
    from weed_instance_segmentation_b200 import modules, synth

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cfg = Mask2FormerConfig()
    ref_layer = m2f.Mask2FormerPixelDecoderEncoderLayer(cfg).to(dev).train()
    # realistic offsets: the init pattern (M2F:2116-2128) moved off the grid lines
    with torch.no_grad():
        ref_layer.self_attn.sampling_offsets.weight.zero_()
        ref_layer.self_attn.sampling_offsets.bias.copy_(
            (synth.init_offsets(8, 3, 4) + 0.1 + 0.3 * torch.rand(8, 3, 4, 2)).reshape(-1))
    S = sum(h * w for h, w in SHAPES)
    B = args.batch
    x = torch.randn(B, S, 256, device=dev)
    pos = torch.randn(B, S, 256, device=dev)
    mask = torch.zeros(B, S, dtype=torch.bool, device=dev)
    ref_pts = synth.reference_points(SHAPES, device=dev)[None].expand(B, -1, -1, -1).contiguous()
    lsi = torch.tensor(synth.level_start_index(SHAPES), device=dev)
    go = torch.randn(B, S, 256, device=dev)

    def variant(name):
        layer = copy.deepcopy(ref_layer)
        if name in ("modules", "fused", "fused+norm", "fused+norm+linear", "full"):
            layer = modules.EncoderLayer.from_hf(layer)
            layer.self_attn.assume_no_padding = True
            layer.self_attn.fused_prologue = name != "modules"
            layer.fused_norm = name in ("fused+norm", "fused+norm+linear", "full")
            layer.fused_linear = layer.self_attn.fused_linear = name in ("fused+norm+linear", "full")
            layer.self_attn.implicit_reference_points = name == "full"
        return layer

    def run(layer, patched):
        xin = x.clone().requires_grad_(True)
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if args.amp == "bf16" else torch.autocast("cuda", enabled=False)

        def step():
            xin.grad = None
            for p in layer.parameters():
                p.grad = None
            with ctx:
                out = layer(xin, mask, position_embeddings=pos, reference_points=ref_pts, spatial_shapes_list=SHAPES,
                            level_start_index=lsi)[0]
            out.backward(go.to(out.dtype))
            return out

        def timed():
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            # three timed repeats of `steps`, the median reported: a variant's first repeat occasionally carries one-off
            # library work (cuBLASLt kernel selection for a new epilogue) that the warm-up did not trigger
            reps = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    out = step()
                e1.record()
                torch.cuda.synchronize()
                reps.append(e0.elapsed_time(e1) / args.steps)
            return sorted(reps)[1], out.detach().float(), xin.grad.detach().float()

        if patched:
            with wis.installed():
                return timed()
        return timed()

    results, outs = {}, {}
    for name, patched in (("reference", False), ("function", True), ("modules", True), ("fused", True), ("fused+norm", True),
                          ("fused+norm+linear", True), ("full", True)):
        ms, out, gx = run(variant(name), patched)
        results[name] = {"ms_per_layer_fwd_bwd": ms}
        outs[name] = (out, gx)
        torch.cuda.empty_cache()
    r_out, r_gx = outs["reference"]
    for name in ("function", "modules", "fused", "fused+norm", "fused+norm+linear", "full"):
        o, g = outs[name]
        results[name]["out_rel_err_vs_reference"] = ((o - r_out).abs().max() / r_out.abs().max()).item()
        results[name]["grad_input_rel_err_vs_reference"] = ((g - r_gx).abs().max() / r_gx.abs().max()).item()
        results[name]["speedup_vs_reference"] = results["reference"]["ms_per_layer_fwd_bwd"] / results[name]["ms_per_layer_fwd_bwd"]
    print(json.dumps({
        "metric": "pixel_decoder_encoder_layer_fwd_bwd_ms", "unit": "ms", "higher_is_better": False,
        "config": {"workload": f"one Mask2Former pixel-decoder encoder layer, B={B}, S={S}, d_model=256, "
                               f"{'bf16 autocast' if args.amp == 'bf16' else 'fp32'}", "spatial_shapes": SHAPES},
        "steps": args.steps, "warmup": args.warmup, "results": results,
    }), flush=True)


if __name__ == "__main__":
    main()
