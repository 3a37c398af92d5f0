"""CPU oracle for multi-scale deformable attention -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker. The
product package ``weed_instance_segmentation_b200`` never imports it and has no
CPU fallback.

Three implementations of the same function live here:

* ``c_forward`` / ``c_backward``  -- plain C restatement (``msda_oracle.c``), fp32 or fp64.
* ``np_forward`` / ``np_backward`` -- vectorised numpy restatement (``msda_numpy.py``).
* ``hf_forward`` / ``hf_forward_backward`` -- the reference's own implementation,
  i.e. ``transformers.models.mask2former.modeling_mask2former.multi_scale_deformable_attention``
  (M2F:798-837) and autograd through it (``hf_reference.py``). It is used to pin
  the two restatements and as the ``--impl reference`` arm of ``bench.py``.

Parity status: pinned against the reference's own implementation run in the
build container (``tests/golden/*.npz`` made by ``tests/golden/make_golden.py``).
The reference repository itself ships no tests or golden vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .msda_numpy import np_backward, np_forward  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``msda_oracle.c`` with gcc (a few hundred ms). Returns the .so path."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-fno-fast-math", "-ffp-contract=off",
             "-o", _SO, src, "-lm"]
        )
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        for suffix in ("f32", "f64"):
            getattr(lib, f"msda_oracle_forward_{suffix}").restype = ctypes.c_int
            getattr(lib, f"msda_oracle_backward_{suffix}").restype = ctypes.c_int
        _lib = lib
    return _lib


def _prep(value, shapes, loc, attn, dtype):
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    attn = np.ascontiguousarray(attn, dtype=dtype)
    shapes = np.ascontiguousarray(np.asarray(shapes, dtype=np.int32).reshape(-1, 2))
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    assert shapes.shape[0] == L and attn.shape == (B, Q, H, L, P)
    return value, shapes, loc, attn, (B, S, Q, H, D, L, P)


def level_start_index(shapes) -> np.ndarray:
    """M2F:1321: ``cat(zeros(1), (H_l*W_l).cumsum(0)[:-1])``."""
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    n = shapes[:, 0] * shapes[:, 1]
    return np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def c_forward(value, shapes, loc, attn, level_start=None, dtype=np.float64):
    """C oracle forward. Returns ``out (B, Q, H*D)`` in ``dtype``."""
    lib = _load()
    value, shapes, loc, attn, dims = _prep(value, shapes, loc, attn, dtype)
    B, S, Q, H, D, L, P = dims
    ls = level_start_index(shapes) if level_start is None else np.ascontiguousarray(level_start, dtype=np.int64)
    out = np.empty((B, Q, H * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    rc = fn(_ptr(value), _ptr(shapes), _ptr(ls), _ptr(loc), _ptr(attn), _ptr(out),
            *[ctypes.c_int(int(x)) for x in dims])
    if rc != 0:
        raise ValueError(f"msda_oracle_forward: bad arguments (code {rc})")
    return out


def c_backward(value, shapes, loc, attn, grad_out, level_start=None, dtype=np.float64):
    """C oracle backward. Returns ``(grad_value, grad_loc, grad_attn)`` in ``dtype``."""
    lib = _load()
    value, shapes, loc, attn, dims = _prep(value, shapes, loc, attn, dtype)
    B, S, Q, H, D, L, P = dims
    go = np.ascontiguousarray(grad_out, dtype=dtype).reshape(B, Q, H * D)
    ls = level_start_index(shapes) if level_start is None else np.ascontiguousarray(level_start, dtype=np.int64)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    rc = fn(_ptr(value), _ptr(shapes), _ptr(ls), _ptr(loc), _ptr(attn), _ptr(go),
            _ptr(gv), _ptr(gl), _ptr(ga), *[ctypes.c_int(int(x)) for x in dims])
    if rc != 0:
        raise ValueError(f"msda_oracle_backward: bad arguments (code {rc})")
    return gv, gl, ga


def hf_forward(*args, **kwargs):
    from .hf_reference import hf_forward as f
    return f(*args, **kwargs)


def hf_forward_backward(*args, **kwargs):
    from .hf_reference import hf_forward_backward as f
    return f(*args, **kwargs)
