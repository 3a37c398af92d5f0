"""Host-buffer pipeline (msda_b200_host_pipeline_*, weed_instance_segmentation_b200/host.py) against the oracle.

Same bars as the device-buffer operator: fp32 <= 1e-5 against the fp32 C oracle, bf16 <= 2e-2 against the oracle fed
with the bf16-rounded inputs. The forward result is also bit-identical to the device-buffer call (same kernel, the
chunking only changes which launch an image belongs to).
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [(8, 8), (16, 16), (32, 32)]


@pytest.fixture(scope="module")
def wis():
    import weed_instance_segmentation_b200 as w
    from weed_instance_segmentation_b200 import _cabi, build
    build.build()
    _cabi.load()
    return w


def _inputs(B, shapes, dtype, seed, Q=None, dist="init"):
    from weed_instance_segmentation_b200.synth import msda_inputs
    return msda_inputs(B, shapes, dist=dist, seed=seed, value_dtype=dtype, num_queries=Q)


def _oracle(x, shapes, dtype=np.float32):
    import oracle
    v, lo, a, go = (x[k].float().numpy() for k in ("value", "sampling_locations", "attention_weights", "grad_out"))
    out = oracle.c_forward(v, shapes, lo, a, dtype=dtype)
    return (out,) + tuple(oracle.c_backward(v, shapes, lo, a, go, dtype=dtype))


def _pinned(t):
    return t.detach().cpu().contiguous().pin_memory()


def _run(pipe, x):
    res = pipe.empty_outputs()
    pipe.step(_pinned(x["value"]), _pinned(x["sampling_locations"]), _pinned(x["attention_weights"]),
              _pinned(x["grad_out"]) if pipe.backward else None, **res)
    pipe.join()
    torch.cuda.synchronize()
    return res


def _kink_safe(x, shapes, eps=1e-3):
    loc = x["sampling_locations"].double()
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float64)[None, None, None, :, None, :]
    pix = loc * wh - 0.5
    return ((pix - pix.round()).abs() > eps).all(-1, keepdim=True).numpy()


@pytest.mark.parametrize("chunk,slots", [(1, 2), (2, 3), (8, 2)])
def test_pipeline_fp32_matches_oracle(wis, chunk, slots):
    B = 3
    x = _inputs(B, SHAPES, torch.float32, seed=11)
    with wis.HostPipeline(B, SHAPES, 8, 32, 4, value_dtype=torch.float32, chunk_images=chunk, slots=slots) as pipe:
        got = _run(pipe, x)
    want = _oracle(x, SHAPES)
    safe = _kink_safe(x, SHAPES)
    for name, w in zip(("output", "grad_value", "grad_sampling_locations", "grad_attention_weights"), want):
        g = got[name].numpy().reshape(w.shape)
        if name == "grad_sampling_locations":
            g, w = g * safe, w * safe
        assert rel_err(g, w) <= 1e-5, f"{name}: {rel_err(g, w):.3e}"


def test_pipeline_bf16_matches_oracle_and_device_call(wis):
    B = 4
    x = _inputs(B, SHAPES, torch.bfloat16, seed=12)
    with wis.HostPipeline(B, SHAPES, 8, 32, 4, chunk_images=1, slots=3) as pipe:
        got = _run(pipe, x)
    want = _oracle(x, SHAPES, dtype=np.float64)
    safe = _kink_safe(x, SHAPES)
    for name, w in zip(("output", "grad_value", "grad_sampling_locations", "grad_attention_weights"), want):
        g = got[name].float().numpy().reshape(w.shape)
        if name == "grad_sampling_locations":
            g, w = g * safe, w * safe
        assert rel_err(g, w) <= 2e-2, f"{name}: {rel_err(g, w):.3e}"
    dev = wis.ms_deform_attn(x["value"].cuda(), SHAPES, None, x["sampling_locations"].cuda(), x["attention_weights"].cuda())
    assert torch.equal(dev.cpu(), got["output"].view(dev.shape))


def test_pipeline_back_to_back_steps_overlap_safely(wis):
    """Steps enqueued without waiting in between reuse the staging ring; every step's results stay its own."""
    B = 3
    xs = [_inputs(B, SHAPES, torch.float32, seed=20 + i, dist="trained") for i in range(4)]
    with wis.HostPipeline(B, SHAPES, 8, 32, 4, value_dtype=torch.float32, chunk_images=1, slots=2) as pipe:
        host = [[_pinned(x[k]) for k in ("value", "sampling_locations", "attention_weights", "grad_out")] for x in xs]
        results = [pipe.empty_outputs() for _ in xs]
        for h, r in zip(host, results):
            pipe.step(*h, **r)
        pipe.synchronize()
    for x, r in zip(xs, results):
        want = _oracle(x, SHAPES)
        assert rel_err(r["output"].numpy().reshape(want[0].shape), want[0]) <= 1e-5
        assert rel_err(r["grad_value"].numpy().reshape(want[1].shape), want[1]) <= 1e-5
        assert rel_err(r["grad_attention_weights"].numpy().reshape(want[3].shape), want[3]) <= 1e-5


def test_pipeline_forward_only_ragged_queries(wis):
    import oracle
    B, Q = 2, 77
    shapes = [(5, 7), (9, 4)]
    g = torch.Generator().manual_seed(5)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(B, S, 4, 16, generator=g)
    loc = torch.rand(B, Q, 4, 2, 3, 2, generator=g) * 1.2 - 0.1
    attn = torch.softmax(torch.randn(B, Q, 4, 6, generator=g), -1).view(B, Q, 4, 2, 3)
    with wis.HostPipeline(B, shapes, 4, 16, 3, value_dtype=torch.float32, num_queries=Q, backward=False) as pipe:
        res = pipe.empty_outputs()
        assert list(res) == ["output"]
        pipe.step(_pinned(value), _pinned(loc), _pinned(attn), output=res["output"])
        pipe.synchronize()
        assert pipe.h2d_bytes_per_step == 4 * (value.numel() + loc.numel() + attn.numel())
        assert pipe.d2h_bytes_per_step == 4 * B * Q * 64
    want = oracle.c_forward(value.numpy(), shapes, loc.numpy(), attn.numpy(), dtype=np.float32)
    assert rel_err(res["output"].numpy().reshape(want.shape), want) <= 1e-5


def test_pipeline_pageable_memory_still_correct(wis):
    B = 2
    x = _inputs(B, SHAPES, torch.float32, seed=31)
    with wis.HostPipeline(B, SHAPES, 8, 32, 4, value_dtype=torch.float32) as pipe:
        res = pipe.empty_outputs(pin=False)
        pipe.step(x["value"].contiguous(), x["sampling_locations"].contiguous(), x["attention_weights"].contiguous(),
                  x["grad_out"].contiguous(), **res)
        pipe.synchronize()
    want = _oracle(x, SHAPES)
    assert rel_err(res["output"].numpy().reshape(want[0].shape), want[0]) <= 1e-5
    assert rel_err(res["grad_value"].numpy().reshape(want[1].shape), want[1]) <= 1e-5


def test_pipeline_errors(wis):
    from weed_instance_segmentation_b200 import MSDAError
    with pytest.raises(MSDAError):
        wis.HostPipeline(2, SHAPES, 8, 32, 4, slots=1)
    with pytest.raises(MSDAError):
        wis.HostPipeline(2, SHAPES, 8, 24, 4)  # head dim without a kernel
    with pytest.raises(TypeError):
        wis.HostPipeline(2, SHAPES, 8, 32, 4, value_dtype=torch.float16)
    x = _inputs(2, SHAPES, torch.float32, seed=1)
    with wis.HostPipeline(2, SHAPES, 8, 32, 4, value_dtype=torch.float32) as pipe:
        res = pipe.empty_outputs()
        args = [_pinned(x[k]) for k in ("value", "sampling_locations", "attention_weights", "grad_out")]
        with pytest.raises(RuntimeError):
            pipe.step(args[0].cuda(), *args[1:], **res)
        with pytest.raises(ValueError):
            pipe.step(args[0][:1], *args[1:], **res)
        with pytest.raises(TypeError):
            pipe.step(args[0].bfloat16(), *args[1:], **res)
        with pytest.raises(RuntimeError):
            pipe.step(*args[:3], None, **res)
    with pytest.raises(RuntimeError):
        pipe.step(*args, **res)  # closed
