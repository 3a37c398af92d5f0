"""Tuning run on a GPU box: forward time at BASELINE config 2 against the shared-memory carve-out of the head-pair
kernel (MSDA_B200_FWD_CARVEOUT, percent; what is left of the SM's 256 KB is L1), and the fused-prologue kernels next
to the plain ones.

    python tools/dev_carveout.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weed_instance_segmentation_b200 import synth  # noqa: E402
from weed_instance_segmentation_b200.synth import msda_inputs  # noqa: E402

H, D, P = 8, 32, 4


def sweep(label, env, prob, which, values):
    """Kernel time (library events) for each carve-out value; `None` = the library's default for that kernel."""
    line = f"{label:34s}"
    for v in values:
        if v is None:
            os.environ.pop(env, None)
        else:
            os.environ[env] = str(v)
        ts = []
        for _ in range(9):
            ts.append(bench.profile_kernels(prob, 1)[which])
        ts.sort()
        line += f" | {'dflt' if v is None else v:>4}: {ts[4]:.3f}"
    os.environ.pop(env, None)
    print(line, flush=True)


def main():
    dev = torch.device("cuda", 0)
    c3 = [(31, 41), (61, 81), (121, 162)]
    vals = (None, 100, 86, 72, 58, 44, 29)  # -> 228 / 196 / 164 / 132 / 100 / 64 KB carve-outs
    for tag, shapes, batch in (("config2", None, None), ("config3", c3, 16)):
        for dist in ("init", "trained"):
            pr = bench.Problem(dist, "bf16", dev, seed=0, shapes=shapes, batch=batch)
            sweep(f"{tag}/{dist}/bf16 fwd (head pairs)", "MSDA_B200_FWD_CARVEOUT", pr, "fwd", vals)
            del pr
            torch.cuda.empty_cache()
    for dist in ("init", "trained"):
        pr = bench.Problem(dist, "fp32", dev, seed=0)
        sweep(f"config2/{dist}/fp32 fwd (one head)", "MSDA_B200_FWD1_CARVEOUT", pr, "fwd", vals)
        sweep(f"config2/{dist}/fp32 bwd v1", "MSDA_B200_BWD1_CARVEOUT", pr, "bwd_main", vals)
        del pr
        torch.cuda.empty_cache()
    # plain vs fused prologue through autograd (forward + backward incl. zero-fill and convert), CUDA events
    import weed_instance_segmentation_b200 as wis
    shapes = [(32, 32), (64, 64), (128, 128)]
    B, L = 8, 3
    x = msda_inputs(B, shapes, num_heads=H, head_dim=D, num_points=P, dist="init", seed=5, device="cuda",
                    value_dtype=torch.bfloat16)
    S = x["value"].shape[1]
    g = torch.Generator(device="cuda").manual_seed(1)
    off = (synth.init_offsets(H, L, P).to("cuda")[None, None]
           + 0.5 * torch.randn(B, S, H, L, P, 2, device="cuda", generator=g)).bfloat16().contiguous().requires_grad_(True)
    logits = torch.randn(B, S, H, L * P, device="cuda", generator=g).bfloat16().requires_grad_(True)
    value = x["value"].clone().requires_grad_(True)
    loc = x["sampling_locations"].clone().requires_grad_(True)
    attn = x["attention_weights"].clone().requires_grad_(True)
    go = x["grad_out"]
    lsi = x["level_start_index"]

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def plain():
        for t in (value, loc, attn):
            t.grad = None
        wis.ms_deform_attn(value, shapes, lsi, loc, attn).backward(go)

    def fused():
        for t in (value, off, logits):
            t.grad = None
        wis.ms_deform_attn_fused(value, shapes, None, off, logits, None).backward(go)

    print(f"config2/init autograd step: plain {timed(plain):.3f} ms | fused prologue (implicit reference points) {timed(fused):.3f} ms")


if __name__ == "__main__":
    main()
