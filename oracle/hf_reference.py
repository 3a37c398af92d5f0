"""The reference's own implementation of the hot path, run as-is.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference repository (``/root/reference``) holds no MSDeformAttn code; its
training/eval scripts call HuggingFace ``transformers``
(``/root/reference/models/mask2former/train.py:7,167-173,196``) where the op is
``multi_scale_deformable_attention`` (transformers 5.5.0,
``models/mask2former/modeling_mask2former.py:798-837``). That function is
imported here unmodified; ``transformers`` is part of the image, so this also
runs on the GPU box (nothing under ``/root/reference`` is read).
"""
from __future__ import annotations

import numpy as np
import torch


def _fn():
    from transformers.models.mask2former.modeling_mask2former import multi_scale_deformable_attention
    return multi_scale_deformable_attention


def _shapes_list(shapes):
    return [(int(h), int(w)) for h, w in np.asarray(shapes).reshape(-1, 2)]


def hf_forward_torch(value, shapes, loc, attn):
    """Torch tensors in, torch tensor out; any device/dtype the reference supports."""
    return _fn()(value, _shapes_list(shapes), loc, attn)


def hf_forward(value, shapes, loc, attn, dtype=torch.float32):
    """Numpy in/out convenience wrapper on CPU."""
    v = torch.as_tensor(np.asarray(value)).to(dtype)
    lo = torch.as_tensor(np.asarray(loc)).to(dtype)
    a = torch.as_tensor(np.asarray(attn)).to(dtype)
    with torch.no_grad():
        out = hf_forward_torch(v, shapes, lo, a)
    return out.numpy()


def hf_forward_backward(value, shapes, loc, attn, grad_out, dtype=torch.float32):
    """Forward + autograd backward through the reference function (numpy in/out, CPU)."""
    v = torch.as_tensor(np.asarray(value)).to(dtype).requires_grad_(True)
    lo = torch.as_tensor(np.asarray(loc)).to(dtype).requires_grad_(True)
    a = torch.as_tensor(np.asarray(attn)).to(dtype).requires_grad_(True)
    go = torch.as_tensor(np.asarray(grad_out)).to(dtype)
    out = hf_forward_torch(v, shapes, lo, a)
    out.backward(go.reshape(out.shape))
    return out.detach().numpy(), v.grad.numpy(), lo.grad.numpy(), a.grad.numpy()
