"""Development check of the forward kernels on a GPU box: the head-pair kernel (default) against the one-head kernel
(MSDA_B200_FWD_NO_PAIR=1; bit-identical results expected), plain and fused prologue, then kernel times.

    python tools/dev_fwd.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weed_instance_segmentation_b200 import _cabi, functional, synth  # noqa: E402
from weed_instance_segmentation_b200.synth import msda_inputs  # noqa: E402

H, D, P = 8, 32, 4
lib = _cabi.load()
stream = torch.cuda.current_stream().cuda_stream
ptr = lambda t: t.data_ptr() if t is not None and t.numel() else None  # noqa: E731


def run_plain(x, shapes, adt, profile=False):
    value, loc, attn = x["value"], x["sampling_locations"], x["attention_weights"].to(adt)
    B, S = value.shape[:2]
    Q = loc.shape[1]
    order = functional.query_order_2d(shapes, functional._TILE, "cuda") if Q == S else None
    desc, keep = _cabi.make_desc(B, S, Q, value.shape[2], D, len(shapes), P, _cabi.BF16,
                                 _cabi.BF16 if adt == torch.bfloat16 else _cabi.F32, shapes, x["level_start_index"].tolist(),
                                 _cabi.FLAG_PROFILE if profile else 0)
    out = torch.empty(B, Q, value.shape[2] * D, dtype=value.dtype, device="cuda")
    _cabi.check(lib.msda_b200_forward(desc, ptr(value), ptr(loc), ptr(attn), ptr(out), ptr(order), stream))
    torch.cuda.synchronize()
    return out, (_cabi.profile_ms(_cabi.PROF_FWD) if profile else None)


def run_fused(x, shapes, adt, off, logits, ref, profile=False):
    value = x["value"]
    B, S = value.shape[:2]
    order = functional.query_order_2d(shapes, functional._TILE, "cuda")
    desc, keep = _cabi.make_desc(B, S, S, value.shape[2], D, len(shapes), P, _cabi.BF16,
                                 _cabi.BF16 if adt == torch.bfloat16 else _cabi.F32, shapes, x["level_start_index"].tolist(),
                                 _cabi.FLAG_PROFILE if profile else 0)
    out = torch.empty(B, S, value.shape[2] * D, dtype=value.dtype, device="cuda")
    _cabi.check(lib.msda_b200_forward_fused(desc, ptr(value), ptr(off), ptr(logits), ptr(ref), ptr(out), None, ptr(order),
                                            stream))
    torch.cuda.synchronize()
    return out, (_cabi.profile_ms(_cabi.PROF_FWD) if profile else None)


def main():
    ok = True
    cases = [("c2b1", 1, [(32, 32), (64, 64), (128, 128)], 8), ("odd", 2, [(8, 9), (10, 6), (5, 12)], 2),
             ("two_levels", 1, [(20, 20), (40, 40)], 8), ("c3b1", 1, [(31, 41), (61, 81), (121, 162)], 8)]
    for tag, B, shapes, heads in cases:
        for dist in ("init", "adversarial"):
            for adt in (torch.bfloat16, torch.float32):
                x = msda_inputs(B, shapes, num_heads=heads, head_dim=D, num_points=P, dist=dist, seed=5, device="cuda",
                                value_dtype=torch.bfloat16)
                S, L = x["value"].shape[1], len(shapes)
                g = torch.Generator(device="cuda").manual_seed(1)
                off = (synth.init_offsets(heads, L, P).to("cuda")[None, None]
                       + 0.5 * torch.randn(B, S, heads, L, P, 2, device="cuda", generator=g)).to(adt).contiguous()
                logits = torch.randn(B, S, heads, L * P, device="cuda", generator=g).to(adt)
                ref = synth.reference_points(shapes, device="cuda")[None].expand(B, -1, -1, -1).contiguous()
                os.environ.pop("MSDA_B200_FWD_NO_PAIR", None)
                a, _ = run_plain(x, shapes, adt)
                fa, _ = run_fused(x, shapes, adt, off, logits, ref)
                fi, _ = run_fused(x, shapes, adt, off, logits, None)
                os.environ["MSDA_B200_FWD_NO_PAIR"] = "1"
                b, _ = run_plain(x, shapes, adt)
                fb, _ = run_fused(x, shapes, adt, off, logits, ref)
                fib, _ = run_fused(x, shapes, adt, off, logits, None)
                os.environ.pop("MSDA_B200_FWD_NO_PAIR", None)
                same = (torch.equal(a, b), torch.equal(fa, fb), torch.equal(fi, fib))
                ok &= all(same)
                print(f"{tag:10s} {dist:12s} attn={'bf16' if adt == torch.bfloat16 else 'f32 '} bit-identical to the one-head "
                      f"kernel: plain {same[0]}, fused {same[1]}, fused with implicit reference points {same[2]}", flush=True)
    for tag, B, shapes in (("config2", 8, [(32, 32), (64, 64), (128, 128)]), ("config3", 16, [(31, 41), (61, 81), (121, 162)]),
                           ("config5", 4, [(64, 64), (128, 128), (256, 256)])):
        for dist in ("init", "trained"):
            x = msda_inputs(B, shapes, num_heads=H, head_dim=D, num_points=P, dist=dist, seed=5, device="cuda",
                            value_dtype=torch.bfloat16)
            S, L = x["value"].shape[1], len(shapes)
            g = torch.Generator(device="cuda").manual_seed(1)
            off = (synth.init_offsets(H, L, P).to("cuda")[None, None]
                   + 0.5 * torch.randn(B, S, H, L, P, 2, device="cuda", generator=g)).bfloat16().contiguous()
            logits = torch.randn(B, S, H, L * P, device="cuda", generator=g).bfloat16()
            line = f"{tag}/{dist:8s}"
            for name, env in (("pair", None), ("one-head", "1")):
                if env:
                    os.environ["MSDA_B200_FWD_NO_PAIR"] = env
                ts = sorted(run_plain(x, shapes, torch.bfloat16, profile=True)[1] for _ in range(8))
                tf = sorted(run_fused(x, shapes, torch.bfloat16, off, logits, None, profile=True)[1] for _ in range(8))
                os.environ.pop("MSDA_B200_FWD_NO_PAIR", None)
                line += f" | {name}: {ts[4]:.3f} / {tf[4]:.3f}"
            print(line, flush=True)
    print("OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
