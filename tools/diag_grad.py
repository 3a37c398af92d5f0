"""Diagnostic: op-level fp32 gradients, B200 op vs the HF function on the same GPU vs HF in fp64."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import weed_instance_segmentation_b200 as wis
from weed_instance_segmentation_b200.synth import msda_inputs, init_offsets, reference_points
from oracle.hf_reference import hf_forward_torch

shapes = [(32, 32), (64, 64), (128, 128)]
B, H, L, P, D = 2, 8, 3, 4, 32
torch.manual_seed(0)
S = sum(h * w for h, w in shapes)
value = torch.randn(B, S, H, D, device="cuda")
ref = reference_points(shapes, device="cuda")
off = (init_offsets(H, L, P).cuda() + 0.1 + 0.3 * torch.rand(H, L, P, 2, device="cuda"))[None, None].expand(B, S, -1, -1, -1, -1)
wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device="cuda")
loc = (ref[None, :, None, :, None, :] + off / wh[None, None, None, :, None, :]).contiguous()
attn = torch.softmax(torch.randn(B, S, H, L * P, device="cuda"), -1).view(B, S, H, L, P)
go = torch.randn(B, S, H * D, device="cuda")

def run(fn, dt):
    v, lo, a = (t.detach().to(dt).clone().requires_grad_(True) for t in (value, loc, attn))
    out = fn(v, shapes, lo, a)
    out.backward(go.to(dt))
    return [t.detach().double() for t in (out, v.grad, lo.grad, a.grad)]

mine = run(lambda v, s, lo, a: wis.ms_deform_attn(v, s, None, lo, a), torch.float32)
hf32 = run(hf_forward_torch, torch.float32)
hf64 = run(hf_forward_torch, torch.float64)
rel = lambda x, y: ((x - y).abs().max() / y.abs().max()).item()
for i, name in enumerate(("out", "grad_value", "grad_loc", "grad_attn")):
    print(f"{name:11s} mine-vs-hf64 {rel(mine[i], hf64[i]):.3e}   hf32-vs-hf64 {rel(hf32[i], hf64[i]):.3e}   mine-vs-hf32 {rel(mine[i], hf32[i]):.3e}")
