"""Pin the CPU oracle (C and numpy restatements) against the reference's own outputs.

Golden vectors come from the reference implementation (HF M2F:798-837 + autograd), see
tests/golden/make_golden.py. Bars: the fp64 restatements must match the reference run in
fp64 to 1e-10, and the reference's fp32 run to 1e-5 (the north-star fp32 bar).
"""
import numpy as np
import pytest

import oracle
from conftest import rel_err


def _shapes(g):
    return [tuple(int(v) for v in r) for r in g["shapes"]]


def test_golden_present():
    from conftest import golden_names
    assert len(golden_names()) >= 5


def test_c_oracle_matches_reference_f64(golden):
    name, g = golden
    out = oracle.c_forward(g["value"], _shapes(g), g["loc"], g["attn"], dtype=np.float64)
    gv, gl, ga = oracle.c_backward(g["value"], _shapes(g), g["loc"], g["attn"], g["grad_out"], dtype=np.float64)
    assert rel_err(out, g["out_f64"]) < 1e-10, name
    assert rel_err(gv, g["grad_value_f64"]) < 1e-10, name
    assert rel_err(gl, g["grad_loc_f64"]) < 1e-10, name
    assert rel_err(ga, g["grad_attn_f64"]) < 1e-10, name


def test_c_oracle_f32_matches_reference_f32(golden):
    name, g = golden
    out = oracle.c_forward(g["value"], _shapes(g), g["loc"], g["attn"], dtype=np.float32)
    gv, gl, ga = oracle.c_backward(g["value"], _shapes(g), g["loc"], g["attn"], g["grad_out"], dtype=np.float32)
    assert rel_err(out, g["out_f32"]) < 1e-5, name
    assert rel_err(gv, g["grad_value_f32"]) < 1e-5, name
    assert rel_err(gl, g["grad_loc_f32"]) < 1e-5, name
    assert rel_err(ga, g["grad_attn_f32"]) < 1e-5, name


def test_numpy_oracle_matches_reference(golden):
    name, g = golden
    v64 = g["value"].astype(np.float64)
    out = oracle.np_forward(v64, _shapes(g), g["loc"], g["attn"])
    gv, gl, ga = oracle.np_backward(v64, _shapes(g), g["loc"], g["attn"], g["grad_out"])
    assert rel_err(out, g["out_f64"]) < 1e-10, name
    assert rel_err(gv, g["grad_value_f64"]) < 1e-10, name
    assert rel_err(gl, g["grad_loc_f64"]) < 1e-10, name
    assert rel_err(ga, g["grad_attn_f64"]) < 1e-10, name
    # and the reference's own fp32 run sits within the fp32 bar of the fp64 anchor
    assert rel_err(g["out_f32"], g["out_f64"]) < 1e-5, name


def test_oracle_matches_live_reference():
    """Same check against the reference function imported live (transformers is in the image)."""
    pytest.importorskip("transformers")
    from weed_instance_segmentation_b200.synth import msda_inputs
    shapes = [(3, 4), (6, 8), (12, 16)]
    x = msda_inputs(2, shapes, dist="trained", seed=11)
    v, lo, a, go = (x[k].numpy() for k in ("value", "sampling_locations", "attention_weights", "grad_out"))
    out, gv, gl, ga = oracle.hf_forward_backward(v, shapes, lo, a, go)
    c_out = oracle.c_forward(v, shapes, lo, a)
    c_gv, c_gl, c_ga = oracle.c_backward(v, shapes, lo, a, go)
    assert rel_err(out, c_out) < 1e-5
    assert rel_err(gv, c_gv) < 1e-5
    assert rel_err(gl, c_gl) < 1e-5
    assert rel_err(ga, c_ga) < 1e-5


def test_oracle_rejects_bad_levels():
    v = np.zeros((1, 4, 1, 8))
    loc = np.zeros((1, 2, 1, 1, 1, 2))
    attn = np.zeros((1, 2, 1, 1, 1))
    with pytest.raises(ValueError):
        oracle.c_forward(v, [(3, 3)], loc, attn)  # 9 pixels do not fit in S=4


def test_empty_batch_and_queries():
    v = np.zeros((0, 4, 2, 8))
    loc = np.zeros((0, 3, 2, 1, 2, 2))
    attn = np.zeros((0, 3, 2, 1, 2))
    assert oracle.c_forward(v, [(2, 2)], loc, attn).shape == (0, 3, 16)
    v = np.ones((2, 4, 2, 8))
    loc = np.zeros((2, 0, 2, 1, 2, 2))
    attn = np.zeros((2, 0, 2, 1, 2))
    assert oracle.c_forward(v, [(2, 2)], loc, attn).shape == (2, 0, 16)
    gv, gl, ga = oracle.c_backward(v, [(2, 2)], loc, attn, np.zeros((2, 0, 16)))
    assert gv.shape == v.shape and not gv.any()
