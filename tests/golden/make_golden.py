"""Generate the golden vectors under tests/golden/ from the reference's own implementation.

Run once in the build container:  python tests/golden/make_golden.py

The reference (``/root/reference``) ships no tests or fixtures, and its MSDeformAttn is
the HuggingFace function ``multi_scale_deformable_attention`` (transformers 5.5.0,
``models/mask2former/modeling_mask2former.py:798-837``) that the reference reaches
through ``Mask2FormerForUniversalSegmentation`` (``/root/reference/models/mask2former/train.py:196``).
Each ``.npz`` stores seeded fp32 inputs and that function's forward output and autograd
gradients, computed once in fp32 (``*_f32``, what the reference runs) and once with the same
inputs promoted to fp64 (``*_f64``, the high-precision anchor).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.hf_reference import hf_forward_backward  # noqa: E402
from weed_instance_segmentation_b200.synth import msda_inputs  # noqa: E402


def _case(name, B, shapes, H, D, P, dist, seed, Q=None, edit=None):
    x = msda_inputs(B, shapes, num_heads=H, head_dim=D, num_points=P, dist=dist, seed=seed, num_queries=Q)
    value = x["value"].numpy()
    loc = x["sampling_locations"].numpy().copy()
    attn = x["attention_weights"].numpy()
    go = x["grad_out"].numpy()
    if edit is not None:
        edit(loc, shapes)
    rec = {
        "value": value, "shapes": np.asarray(shapes, dtype=np.int32), "loc": loc, "attn": attn, "grad_out": go,
    }
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        out, gv, gl, ga = hf_forward_backward(value, shapes, loc, attn, go, dtype=dt)
        rec[f"out_{tag}"], rec[f"grad_value_{tag}"] = out, gv
        rec[f"grad_loc_{tag}"], rec[f"grad_attn_{tag}"] = gl, ga
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def _edges(loc, shapes):
    """Overwrite the first queries with exact edge / centre / far-outside locations."""
    flat = loc.reshape(loc.shape[0], -1, 2)
    specials = [0.0, 1.0, 0.5, -0.0, 1e-7, 1.0 - 1e-7, -0.5, 1.5, -3.0, 4.0, 0.25, 0.75]
    k = 0
    for sx in specials:
        for sy in specials:
            if k >= flat.shape[1]:
                return
            flat[:, k, 0] = sx
            flat[:, k, 1] = sy
            k += 1


def main():
    # Mask2Former pixel-decoder geometry (H=8, D=32, L=3, P=4), Q == S, freshly-initialised offsets
    _case("m2f_init_small", 1, [(2, 2), (3, 4), (5, 6)], 8, 32, 4, "init", 1)
    # same geometry, trained-like offsets, batch 2
    _case("m2f_trained_small", 2, [(1, 2), (3, 3), (4, 5)], 8, 32, 4, "trained", 2)
    # odd everything, Q != S, locations spilling outside [0,1]
    _case("odd_adversarial", 2, [(3, 5), (6, 7)], 2, 8, 3, "adversarial", 3, Q=11)
    # degenerate 1-pixel-wide levels plus exact edge coordinates
    _case("degenerate_edges", 1, [(1, 1), (1, 7), (5, 1), (3, 4)], 4, 16, 2, "adversarial", 4, Q=40, edit=_edges)
    # a single level / single point / single head
    _case("single", 1, [(5, 6)], 1, 32, 1, "adversarial", 5, Q=17)


if __name__ == "__main__":
    main()
