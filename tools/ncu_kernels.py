#!/usr/bin/env python
"""One launch of every kernel in libmsda_b200.so at its production geometry, for an `ncu --set full` capture:

    python tools/ncu_kernels.py                       # must exit 0 on its own first
    ncu --set full --clock-control none -k regex:'msda_|point_sample|add_layernorm|colsum' \
        -o gpurun_out/r01_all_kernels python tools/ncu_kernels.py

MSDeformAttn forward / backward (bf16 sorted, fp32 v1, fused prologue) at BASELINE config 2, the encoder-layer epilogue
kernels at the same token count, and the loss-path point sampling at the config-4 loss geometry.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import weed_instance_segmentation_b200 as wis  # noqa: E402
from weed_instance_segmentation_b200 import layer_norm, linear, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    # 1-2. plain operator, both contracts (direct C-ABI calls as in bench.py)
    for dt in ("bf16", "fp32"):
        p = bench.Problem("init", dt, dev, seed=0)
        p.fwd()
        p.bwd()
        torch.cuda.synchronize()
        del p
        torch.cuda.empty_cache()
    # 3. fused prologue (softmax + locations inside the kernels)
    shapes = bench.SHAPES_C2
    B, H, D, L, P = bench.B_PER_GPU, bench.H, bench.D, bench.L, bench.P
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, H, D, device=dev, generator=g).bfloat16().requires_grad_(True)
    off = (synth.init_offsets(H, L, P).to(dev)[None, None] + 0.5 * torch.randn(B, S, H, L, P, 2, device=dev, generator=g))
    off = off.bfloat16().requires_grad_(True)
    logits = torch.randn(B, S, H, L * P, device=dev, generator=g).bfloat16().requires_grad_(True)
    ref = synth.reference_points(shapes, device=dev)[None].expand(B, -1, -1, -1).contiguous()
    out = wis.ms_deform_attn_fused(value, shapes, None, off, logits, ref)
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    del value, off, logits, out
    # 4-5. encoder-layer epilogue: residual + LayerNorm, bias-gradient column sum
    x = torch.randn(B * S, 256, device=dev, generator=g).bfloat16().requires_grad_(True)
    r = torch.randn(B * S, 256, device=dev, generator=g).requires_grad_(True)
    w = torch.ones(256, device=dev, requires_grad=True)
    b = torch.zeros(256, device=dev, requires_grad=True)
    y = layer_norm.add_layer_norm(x, r, w, b, 1e-5)
    y.backward(torch.randn_like(y))
    linear.column_sum(x.detach())
    # the layer's closing clamp inside the LayerNorm kernels, and the attention module's bf16 operands from one kernel
    yc = layer_norm.add_layer_norm(x, r, w, b, 1e-5, clamp=float(torch.finfo(torch.float32).max - 1000))
    yc.backward(torch.randn_like(yc))
    hq = torch.randn(B * S, 256, device=dev, generator=g).requires_grad_(True)
    pq = torch.randn(B * S, 256, device=dev, generator=g).requires_grad_(True)
    qb, vb = linear.query_value_cast(hq, pq)
    torch.autograd.backward([qb, vb], [torch.randn_like(qb), torch.randn_like(vb)])
    torch.cuda.synchronize()
    del yc, hq, pq, qb, vb
    del x, r, y
    torch.cuda.empty_cache()
    # 6. loss / matcher path at the config-4 loss geometry
    from test_criterion_host import make_problem
    loss, masks, classes, mask_labels, class_labels = make_problem(
        1, B=8, Q=100, C=5, L=10, h=256, w=256, H=1024, W=1024, n_tgt=(3, 9, 14, 1, 20, 7, 5, 11), num_points=12544)
    crit = wis.convert_criterion(loss.cuda())
    ms = [m.cuda().requires_grad_(True) for m in masks]
    cs = [c.cuda().requires_grad_(True) for c in classes]
    aux = [{"masks_queries_logits": m, "class_queries_logits": c} for m, c in zip(ms[:-1], cs[:-1])]
    outd = crit(ms[-1], cs[-1], [m.to(torch.uint8).cuda() for m in mask_labels], [c.cuda() for c in class_labels], aux)
    sum(outd.values()).backward()
    torch.cuda.synchronize()
    print("ncu_kernels: done, own kernel launches:", wis._cabi.launch_count())


if __name__ == "__main__":
    main()
