"""Vectorised numpy restatement of multi-scale deformable attention.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``transformers`` 5.5.0 ``models/mask2former/modeling_mask2former.py:798-837``
(the function the reference calls through ``Mask2FormerForUniversalSegmentation``,
``/root/reference/models/mask2former/train.py:196``):

* M2F:807 + M2F:822-824 -> pixel coordinates ``x = loc_x*W - 0.5``, ``y = loc_y*H - 0.5``
  (``2*loc-1`` followed by ``grid_sample(align_corners=False)``), four-corner bilinear,
  each corner zeroed independently when out of range (``padding_mode="zeros"``).
* M2F:806,815 -> level ``l`` is rows ``[start_l, start_l + H_l*W_l)`` of ``S``, row-major.
* M2F:832-837 -> ``out[b, q, h*D + d] = sum_{l,p} attn * sample``.
"""
from __future__ import annotations

import numpy as np


def _level_start(shapes):
    n = shapes[:, 0].astype(np.int64) * shapes[:, 1].astype(np.int64)
    return np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)


def _corners(loc_l, Hl, Wl):
    """loc_l (B,Q,H,P,2) -> corner indices, bilinear weights, derivative weights, validity."""
    dt = loc_l.dtype.type
    gx = dt(2) * loc_l[..., 0] - dt(1)
    gy = dt(2) * loc_l[..., 1] - dt(1)
    px = ((gx + dt(1)) * dt(Wl) - dt(1)) / dt(2)
    py = ((gy + dt(1)) * dt(Hl) - dt(1)) / dt(2)
    with np.errstate(invalid="ignore"):
        fx = np.floor(px)
        fy = np.floor(py)
    lx = px - fx
    ly = py - fy
    finite = np.isfinite(px) & np.isfinite(py)
    x0 = np.where(finite, np.clip(fx, -2, Wl + 1), -2).astype(np.int64)
    y0 = np.where(finite, np.clip(fy, -2, Hl + 1), -2).astype(np.int64)
    out = []
    one = dt(1)
    for c in range(4):
        cx, cy = c & 1, c >> 1
        xx = x0 + cx
        yy = y0 + cy
        ok = (xx >= 0) & (xx < Wl) & (yy >= 0) & (yy < Hl)
        wx = lx if cx else one - lx
        wy = ly if cy else one - ly
        dwx = (one if cx else -one) * wy
        dwy = (one if cy else -one) * wx
        idx = np.where(ok, yy * Wl + xx, 0)
        out.append((idx, wx * wy, dwx, dwy, ok))
    return out


def np_forward(value, shapes, loc, attn, level_start=None):
    """``value (B,S,H,D)``, ``loc (B,Q,H,L,P,2)``, ``attn (B,Q,H,L,P)`` -> ``(B,Q,H*D)``."""
    value = np.asarray(value)
    loc = np.asarray(loc, dtype=value.dtype)
    attn = np.asarray(attn, dtype=value.dtype)
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    ls = _level_start(shapes) if level_start is None else np.asarray(level_start, dtype=np.int64)
    out = np.zeros((B, Q, H, D), dtype=value.dtype)
    bi = np.arange(B)[:, None, None, None]
    hi = np.arange(H)[None, None, :, None]
    for l in range(L):
        Hl, Wl = int(shapes[l, 0]), int(shapes[l, 1])
        vl = value[:, ls[l]:ls[l] + Hl * Wl]  # (B, HW, H, D)
        for idx, w, _, _, ok in _corners(loc[:, :, :, l], Hl, Wl):
            g = vl[bi, idx, hi]  # (B,Q,H,P,D)
            coef = np.where(ok, w * attn[:, :, :, l], 0)
            out += (g * coef[..., None]).sum(axis=3)
    return out.reshape(B, Q, H * D)


def np_backward(value, shapes, loc, attn, grad_out, level_start=None):
    """Analytic gradient of :func:`np_forward`. Returns ``(grad_value, grad_loc, grad_attn)``."""
    value = np.asarray(value)
    dt = value.dtype
    loc = np.asarray(loc, dtype=dt)
    attn = np.asarray(attn, dtype=dt)
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    go = np.asarray(grad_out, dtype=dt).reshape(B, Q, H, D)
    ls = _level_start(shapes) if level_start is None else np.asarray(level_start, dtype=np.int64)
    gv = np.zeros_like(value)
    gl = np.zeros_like(loc)
    ga = np.zeros_like(attn)
    bi = np.arange(B)[:, None, None, None]
    hi = np.arange(H)[None, None, :, None]
    bb = np.broadcast_to(bi, (B, Q, H, P))
    hh = np.broadcast_to(hi, (B, Q, H, P))
    for l in range(L):
        Hl, Wl = int(shapes[l, 0]), int(shapes[l, 1])
        vl = value[:, ls[l]:ls[l] + Hl * Wl]
        gvl = gv[:, ls[l]:ls[l] + Hl * Wl]
        a = attn[:, :, :, l]
        for idx, w, dwx, dwy, ok in _corners(loc[:, :, :, l], Hl, Wl):
            g = vl[bi, idx, hi]  # (B,Q,H,P,D)
            dot = np.where(ok, (g * go[:, :, :, None, :]).sum(-1), 0)
            ga[:, :, :, l] += w * dot
            gl[:, :, :, l, :, 0] += dt.type(Wl) * a * dwx * dot
            gl[:, :, :, l, :, 1] += dt.type(Hl) * a * dwy * dot
            coef = np.where(ok, a * w, 0)
            np.add.at(gvl, (bb, idx, hh), coef[..., None] * go[:, :, :, None, :])
    return gv, gl, ga
