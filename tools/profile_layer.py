"""Kernel-level time split of one B200 encoder layer fwd+bwd under bf16 autocast (torch profiler)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import Mask2FormerConfig
from transformers.models.mask2former import modeling_mask2former as m2f
import weed_instance_segmentation_b200 as wis
from weed_instance_segmentation_b200 import modules, synth

SHAPES = [(32, 32), (64, 64), (128, 128)]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
layer = modules.EncoderLayer.from_hf(m2f.Mask2FormerPixelDecoderEncoderLayer(Mask2FormerConfig()).to(dev).train())
layer.self_attn.assume_no_padding = True
layer.self_attn.fused_prologue = True
# the full B200 layer (bench_layer.py's "fused+norm+linear"): fused residual + LayerNorm, projections with the fused
# bias-gradient reduction, fc1 bias + ReLU epilogue -- PROFILE_LAYER_PLAIN=1 keeps only the fused prologue
if os.environ.get("PROFILE_LAYER_PLAIN", "0") != "1":
    layer.fused_norm = True
    layer.fused_linear = layer.self_attn.fused_linear = True
S = sum(h * w for h, w in SHAPES); B = 8
x = torch.randn(B, S, 256, device=dev, requires_grad=True); pos = torch.randn(B, S, 256, device=dev)
mask = torch.zeros(B, S, dtype=torch.bool, device=dev)
ref = synth.reference_points(SHAPES, device=dev)[None].expand(B, -1, -1, -1).contiguous()
lsi = torch.tensor(synth.level_start_index(SHAPES), device=dev); go = torch.randn(B, S, 256, device=dev)

AMP = os.environ.get("PROFILE_LAYER_AMP", "bf16") == "bf16"  # PROFILE_LAYER_AMP=none: the float32 layer


def step():
    x.grad = None
    for p_ in layer.parameters():
        p_.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=AMP):
        out = layer(x, mask, position_embeddings=pos, reference_points=ref, spatial_shapes_list=SHAPES, level_start_index=lsi)[0]
    out.backward(go.to(out.dtype))

for _ in range(3): step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=110))
