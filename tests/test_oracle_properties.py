"""Property tests of the CPU oracle (hypothesis): the checker itself obeys the identities the GPU tests rely on."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

shape_st = st.lists(st.tuples(st.integers(1, 6), st.integers(1, 6)), min_size=1, max_size=3)


def _inputs(shapes, B, Q, H, D, P, seed, spread):
    rng = np.random.default_rng(seed)
    L, S = len(shapes), sum(h * w for h, w in shapes)
    value = rng.standard_normal((B, S, H, D))
    loc = rng.uniform(-spread, 1 + spread, (B, Q, H, L, P, 2))
    attn = rng.random((B, Q, H, L, P))
    go = rng.standard_normal((B, Q, H * D))
    return value, loc, attn, go


@settings(max_examples=25, deadline=None)
@given(shapes=shape_st, B=st.integers(1, 2), Q=st.integers(1, 5), H=st.integers(1, 2), P=st.integers(1, 3),
       seed=st.integers(0, 10_000), spread=st.floats(0.0, 0.6))
def test_backward_is_the_adjoint_of_the_forward(shapes, B, Q, H, P, seed, spread):
    D = 4
    value, loc, attn, go = _inputs(shapes, B, Q, H, D, P, seed, spread)
    out = oracle.c_forward(value, shapes, loc, attn)
    gv, gl, ga = oracle.c_backward(value, shapes, loc, attn, go)
    lhs = float((go * out).sum())
    assert abs(float((gv * value).sum()) - lhs) <= 1e-9 * max(1.0, abs(lhs))   # linear in value
    assert abs(float((ga * attn).sum()) - lhs) <= 1e-9 * max(1.0, abs(lhs))    # linear in attn
    # numpy restatement agrees with the C one
    assert np.allclose(oracle.np_forward(value, shapes, loc, attn), out, atol=1e-12)
    gv2, gl2, ga2 = oracle.np_backward(value, shapes, loc, attn, go)
    assert np.allclose(gv2, gv, atol=1e-12) and np.allclose(gl2, gl, atol=1e-12) and np.allclose(ga2, ga, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(shapes=shape_st, seed=st.integers(0, 10_000))
def test_sampling_a_pixel_centre_returns_the_pixel(shapes, seed):
    """loc = ((x+0.5)/W, (y+0.5)/H) lands exactly on pixel (x, y): px = x, py = y (M2F:807 + align_corners=False)."""
    rng = np.random.default_rng(seed)
    L, S = len(shapes), sum(h * w for h, w in shapes)
    value = rng.standard_normal((1, S, 1, 3))
    lvl = int(rng.integers(0, L))
    Hl, Wl = shapes[lvl]
    y, x = int(rng.integers(0, Hl)), int(rng.integers(0, Wl))
    loc = np.zeros((1, 1, 1, L, 1, 2))
    loc[..., 0], loc[..., 1] = -5.0, -5.0            # every other level samples far outside: contributes 0
    loc[0, 0, 0, lvl, 0] = ((x + 0.5) / Wl, (y + 0.5) / Hl)
    attn = np.ones((1, 1, 1, L, 1))
    out = oracle.c_forward(value, shapes, loc, attn)
    start = int(oracle.level_start_index(shapes)[lvl])
    assert np.allclose(out[0, 0], value[0, start + y * Wl + x, 0], atol=1e-12)


@settings(max_examples=20, deadline=None)
@given(shapes=shape_st, seed=st.integers(0, 10_000))
def test_location_gradient_matches_finite_differences_away_from_grid_lines(shapes, seed):
    rng = np.random.default_rng(seed)
    value, loc, attn, go = _inputs(shapes, 1, 2, 1, 3, 2, seed, 0.2)
    # keep every sample at least 0.05 px from a grid line so the central difference stays on one bilinear patch
    wh = np.asarray([[w, h] for h, w in shapes], dtype=np.float64)[None, None, None, :, None, :]
    pix = loc * wh - 0.5
    frac = pix - np.floor(pix)
    pix = np.floor(pix) + np.clip(frac, 0.05, 0.95)
    loc = (pix + 0.5) / wh
    _, gl, _ = oracle.c_backward(value, shapes, loc, attn, go)
    idx = tuple(int(rng.integers(0, n)) for n in loc.shape)
    eps = 1e-6
    lp, lm = loc.copy(), loc.copy()
    lp[idx] += eps
    lm[idx] -= eps
    fd = float((go * (oracle.c_forward(value, shapes, lp, attn) - oracle.c_forward(value, shapes, lm, attn))).sum()) / (2 * eps)
    assert abs(fd - gl[idx]) <= 1e-5 * max(1.0, abs(fd))
